"""Error type of the reference API (VPint/MRP.py:489-490)."""


class VPintError(Exception):
    pass
