"""Import shim, see matplotlib/__init__.py."""


def __getattr__(name):  # pragma: no cover
    def _missing(*_a, **_k):
        raise RuntimeError(f"matplotlib.pyplot.{name} is not available in the oracle harness")
    return _missing
