"""speechbrain.nnet.activations stand-in: ``Swish`` = x * sigmoid(beta * x)."""
import torch


class Swish(torch.nn.Module):
    def __init__(self, beta: float = 1.0):
        super().__init__()
        self.beta = beta
        self.sigmoid = torch.nn.Sigmoid()

    def forward(self, x):
        return x * self.sigmoid(self.beta * x)
