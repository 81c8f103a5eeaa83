"""Bilinear (matrix factorisation) representation (reference:
spotlight/factorization/representations.py:12-91).

Same constructor, sub-module names and state-dict keys as the reference.  `forward` is the fused
CUDA gather + dot + bias + sigmoid kernel (mfb_predict_pairs); note that, as in this fork (F1), it
returns sigmoid(score), not the raw score.  Training does not differentiate through `forward`:
ImplicitFactorizationModel drives the fused step kernels instead (implicit.py)."""
import torch
import torch.nn as nn

from spotlight.layers import ScaledEmbedding, ZeroEmbedding


class BilinearNet(nn.Module):

    def __init__(self, num_users, num_items, embedding_dim=32,
                 user_embedding_layer=None, item_embedding_layer=None, sparse=False):
        super(BilinearNet, self).__init__()
        self.embedding_dim = embedding_dim
        self.user_embeddings = (user_embedding_layer if user_embedding_layer is not None
                                else ScaledEmbedding(num_users, embedding_dim, sparse=sparse))
        self.item_embeddings = (item_embedding_layer if item_embedding_layer is not None
                                else ScaledEmbedding(num_items, embedding_dim, sparse=sparse))
        self.user_biases = ZeroEmbedding(num_users, 1, sparse=sparse)
        self.item_biases = ZeroEmbedding(num_items, 1, sparse=sparse)
        self.__dict__['_mfb_engine'] = None     # not a sub-module, not part of the state dict

    # -- native engine plumbing ---------------------------------------------------------------
    def _engine(self):
        from recommendation_gans_b200.engine import MFEngine
        eng = self.__dict__.get('_mfb_engine')
        if eng is None or eng._params[0].data_ptr() != self.user_embeddings.weight.data_ptr():
            eng = MFEngine(self)                 # forward-only handle on the current storage
            self.__dict__['_mfb_engine'] = eng
        return eng

    def _attach_engine(self, engine):
        self.__dict__['_mfb_engine'] = engine

    def _apply(self, fn, *args, **kwargs):
        # .cuda()/.to() re-allocate parameter storage: drop the handle bound to the old pointers
        self.__dict__['_mfb_engine'] = None
        return super(BilinearNet, self)._apply(fn, *args, **kwargs)

    def forward(self, user_ids, item_ids):
        """sigmoid(<U[u],V[i]> + bu[u] + bi[i]) for id tensors of equal length."""
        users = torch.as_tensor(user_ids).reshape(-1)
        items = torch.as_tensor(item_ids).reshape(-1)
        if users.numel() == 1 and items.numel() == 1:
            # the reference squeezes the [1, D] embeddings to 1-D and then fails in .sum(1)
            raise IndexError('Dimension out of range (expected to be in range of [-1, 0], but got 1)')
        return self._engine().predict_pairs(users, items)
