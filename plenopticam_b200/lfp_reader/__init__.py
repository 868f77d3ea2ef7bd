from .bayer import comp_bayer, comp_bayer_device  # noqa: F401
