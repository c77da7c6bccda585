import os
from .... import REFERENCE_ROOT

_ref = os.path.join(REFERENCE_ROOT, "speechbrain", "lobes", "models", "transformer")
if os.path.isdir(_ref):
    __path__.append(_ref)
