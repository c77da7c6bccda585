"""Mirror of speechbrain.nnet.activations for the one symbol the hot path uses."""
from .._host import Swish  # noqa: F401
