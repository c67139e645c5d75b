"""Restatement shim for pytorch3d v0.7.2 (test infrastructure only).

pytorch3d is a third-party dependency of the reference (README.md:68-74) that is not
vendored under /root/reference and not installable here (no wheel, no network).  The three
functions on the hot path are restated in `transforms.py` from their published definitions;
see SURVEY.md A8.  `io` only needs to be importable (utils/misc.py:5)."""
from . import transforms  # noqa: F401
from . import io  # noqa: F401
