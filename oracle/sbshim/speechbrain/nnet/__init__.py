import os
from .. import REFERENCE_ROOT

_ref = os.path.join(REFERENCE_ROOT, "speechbrain", "nnet")
if os.path.isdir(_ref):
    __path__.append(_ref)

from . import containers, linear, normalization, activations, attention  # noqa: E402,F401
from . import hypermixing, CNN, embedding  # noqa: E402,F401
