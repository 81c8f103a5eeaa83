"""Bilinear (matrix factorisation) representation (reference:
spotlight/factorization/representations.py:12-91).

Same constructor, sub-module names and state-dict keys as the reference.  `forward` is the fused
CUDA gather + dot + bias + sigmoid kernel (mfb_predict_pairs); note that, as in this fork (F1), it
returns sigmoid(score), not the raw score.  Training does not differentiate through `forward`:
ImplicitFactorizationModel drives the fused step kernels instead (implicit.py)."""
import torch
import torch.nn as nn

from spotlight.layers import ScaledEmbedding, ZeroEmbedding


class _BilinearFunction(torch.autograd.Function):
    """sigmoid(<U[u],V[i]> + bu[u] + bi[i]) with the fused CUDA forward and a backward that produces the same dense
    table gradients as autograd through the reference's nn.Embedding expressions (embedding_dense_backward).  Used
    when a caller differentiates through `forward` itself -- a hand-written loop `loss(net(u, i), net(u', i'))
    .backward(); optimizer.step()` -- instead of ImplicitFactorizationModel's fused step.  The backward is plain
    torch: this is the compatibility path, not the hot path."""

    @staticmethod
    def forward(ctx, net, users, items, ue, ie, ub, ib):
        out = net._engine().predict_pairs(users, items)
        ctx.save_for_backward(users, items, ue, ie, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        users, items, ue, ie, out = ctx.saved_tensors
        dz = (grad_out * out * (1.0 - out)).unsqueeze(1)
        u_rows, i_rows = ue.detach()[users], ie.detach()[items]
        g_ue = torch.zeros_like(ue).index_add_(0, users, dz * i_rows)
        g_ie = torch.zeros_like(ie).index_add_(0, items, dz * u_rows)
        g_ub = torch.zeros(ue.shape[0], 1, dtype=ue.dtype, device=ue.device).index_add_(0, users, dz)
        g_ib = torch.zeros(ie.shape[0], 1, dtype=ie.dtype, device=ie.device).index_add_(0, items, dz)
        return None, None, None, g_ue, g_ie, g_ub, g_ib


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
        params = (self.user_embeddings.weight, self.item_embeddings.weight, self.user_biases.weight,
                  self.item_biases.weight)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            dev = params[0].device
            return _BilinearFunction.apply(self, users.to(dev).long(), items.to(dev).long(), *params)
        return self._engine().predict_pairs(users, items)
