"""speechbrain.nnet.embedding stand-in: wraps nn.Embedding as ``self.Embedding``."""
import torch.nn as nn


class Embedding(nn.Module):
    def __init__(self, num_embeddings, embedding_dim=128, consider_as_one_hot=False, blank_id=0):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.Embedding = nn.Embedding(num_embeddings, embedding_dim)

    def forward(self, x):
        return self.Embedding(x.long())
