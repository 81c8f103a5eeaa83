"""Parameter containers of the bilinear model (reference: spotlight/layers.py:23-56).

Both are plain `nn.Embedding` subclasses so state-dict keys, `padding_idx` and `sparse` keep their
torch meaning; only the initialisation differs.  Storage stays torch-owned -- the CUDA kernels
borrow the raw pointers (see recommendation_gans_b200/engine.py).  BloomEmbedding and
ScaledEmbeddingBag (layers.py:59-244) are never instantiated on the MF path and are not provided.
"""
import torch.nn as nn


def _zero_padding_row(module):
    if module.padding_idx is not None:
        module.weight.data[module.padding_idx].fill_(0)


class ScaledEmbedding(nn.Embedding):
    """Embedding initialised from N(0, (1/embedding_dim)^2) with torch's global RNG (layers.py:30-37)."""

    def reset_parameters(self):
        std = 1.0 / self.embedding_dim
        self.weight.data.normal_(0, std)
        _zero_padding_row(self)


class ZeroEmbedding(nn.Embedding):
    """Embedding initialised to zeros; used for the bias columns (layers.py:49-56)."""

    def reset_parameters(self):
        self.weight.data.zero_()
        _zero_padding_row(self)
