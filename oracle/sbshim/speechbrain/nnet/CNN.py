"""Import-only stub."""
import torch.nn as nn


class Conv1d(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("speechbrain.nnet.CNN.Conv1d is outside the SummaryMixing hot path")
