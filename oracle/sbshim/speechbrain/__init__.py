"""Stand-in for the un-vendored SpeechBrain v1.0 package (TEST INFRASTRUCTURE ONLY).

The reference (SamsungLabs/SummaryMixing) is a file overlay onto SpeechBrain, which is
neither vendored under /root/reference nor installable here (no network).  This package
restates, from SpeechBrain's documented behaviour, the 16 symbols the reference imports
 and splices the *unmodified* reference modules into the
same namespace by extending ``__path__`` with the matching /root/reference directory.
Nothing is copied from /root/reference; it is imported from where it lies.

Only ``tests/`` (CPU, in the build container) and ``oracle/gen_golden.py`` import this.
It cannot travel to the GPU box (``/root/reference`` does not exist there).
"""
import os

REFERENCE_ROOT = os.environ.get("SMX_REFERENCE_ROOT", "/root/reference")
_ref = os.path.join(REFERENCE_ROOT, "speechbrain")
if os.path.isdir(_ref):
    __path__.append(_ref)

from . import nnet, lobes, utils, dataio  # noqa: E402,F401
