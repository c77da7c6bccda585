"""speechbrain.nnet.normalization stand-in: ``LayerNorm`` wraps ``torch.nn.LayerNorm`` as ``.norm``."""
import torch


class LayerNorm(torch.nn.Module):
    def __init__(self, input_size=None, input_shape=None, eps=1e-05, elementwise_affine=True):
        super().__init__()
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        if input_shape is not None:
            input_size = input_shape[2:]
        self.norm = torch.nn.LayerNorm(input_size, eps=self.eps, elementwise_affine=self.elementwise_affine)

    def forward(self, x):
        return self.norm(x)
