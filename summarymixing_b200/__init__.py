"""summarymixing_b200 — B200-native (sm_100a) SummaryMixing encoder path.

Mirrors the module surface of SamsungLabs/SummaryMixing (same import sub-paths, constructor signatures,
forward contracts and state_dict keys) over hand-written CUDA kernels reached through the C ABI in
include/smx.h.  CUDA only: there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .nnet.summary_mixing import SummaryMixing  # noqa: F401
from .nnet.activations import Swish  # noqa: F401
from .lobes.models.VanillaNN import ParallelLinear, VanillaNN  # noqa: F401
from .lobes.models.transformer.Conformer import (  # noqa: F401
    ConformerEncoder,
    ConformerEncoderLayer,
    ConvolutionModule,
)
from .lobes.models.transformer.Branchformer import (  # noqa: F401
    BranchformerEncoder,
    BranchformerEncoderLayer,
    ConvolutionBranch,
)

from .graphs import GraphedForward, HostPipeline  # noqa: F401
from ._host import invalidate_weights  # noqa: F401
from . import frontend  # noqa: F401

__version__ = "0.1.0"
