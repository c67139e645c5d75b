"""Import shim (test infrastructure only): matplotlib is imported by reference utils/visualize.py
and utils/sgpa_utils.py at module scope; nothing on the hot path draws."""


def use(*_a, **_k):
    pass


def rc(*_a, **_k):
    pass
