from .cfg import PlenopticamConfig, PARAMS_KEYS, PARAMS_VALS, CALIBS_KEYS  # noqa: F401
