"""Import-only stub."""
import torch.nn as nn


class HyperMixing(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("HyperMixing is outside the SummaryMixing hot path")
