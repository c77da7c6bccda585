import os
from .. import REFERENCE_ROOT

_ref = os.path.join(REFERENCE_ROOT, "speechbrain", "lobes")
if os.path.isdir(_ref):
    __path__.append(_ref)
from . import models  # noqa: E402,F401
