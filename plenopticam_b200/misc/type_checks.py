def rint(val):
    """misc/type_checks.py:57-58 — Python round-half-even on a float64, as an int."""
    return int(round(val))
