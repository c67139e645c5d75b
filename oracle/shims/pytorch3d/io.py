"""Import shim: reference utils/misc.py:5 imports pytorch3d.io; nothing on the hot path uses it."""
