"""Ranking losses for implicit feedback (reference: spotlight/losses.py:20-172).

Same names and call signatures as the reference.  Each loss is one CUDA kernel
(mfb_loss_forward_backward) wrapped in a `torch.autograd.Function`, so it stays differentiable
w.r.t. its prediction tensors exactly like the reference's torch expressions:

  pointwise_loss       BCE(pos, 1).mean() + BCE(neg, 0).mean()          losses.py:42-50
  bpr_loss             mean(1 - sigmoid(pos - neg))                     losses.py:88-96
  hinge_loss           mean(clamp(neg - pos + 1, min=0))                losses.py:121-130
  adaptive_hinge_loss  hinge against max(neg, dim 0)                    losses.py:170-172

All of them operate on *probabilities* because this fork's BilinearNet.forward ends in a sigmoid
(SURVEY F1).  Predictions are 1-D tensors, as produced by ImplicitFactorizationModel; with 1-D
negatives `adaptive_hinge_loss` reduces to a hinge against the single largest negative of the batch
(SURVEY F2/3.3); with [n, b] negatives it is the upstream per-positive maximum (losses.py:170), and hinge / bpr
broadcast the positives over the n rows like the reference's tensor expressions.  `mask` ([b]) is
supported as in the reference (loss*mask summed, divided by mask.sum()).  The explicit-feedback losses
(losses.py:175-250) are outside the accelerated path.
"""
import torch

from recommendation_gans_b200.engine import loss_forward_backward


class _NativeLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, kind, pos, neg, mask=None):
        need_grad = pos.requires_grad or (neg is not None and neg.requires_grad)
        loss, dpos, dneg = loss_forward_backward(kind, pos, neg, need_grad, mask)
        ctx.has_neg = neg is not None
        ctx.save_for_backward(*[t for t in (dpos, dneg) if t is not None])
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        dpos = saved[0] * grad_out
        dneg = saved[1] * grad_out if ctx.has_neg else None
        return None, dpos, dneg, None


def _prepare(name, positive_predictions, negative_predictions, mask):
    pos = positive_predictions
    neg = negative_predictions
    if not pos.is_cuda or (neg is not None and not neg.is_cuda):
        raise RuntimeError('%s: predictions must be CUDA tensors (no CPU path exists)' % name)
    if pos.dim() != 1 or (neg is not None and neg.dim() > (1 if name == 'pointwise_loss' else 2)):
        raise NotImplementedError('%s: 1-D prediction tensors are supported (as produced by '
                                  'ImplicitFactorizationModel), plus [n, b] negatives for the pairwise losses' % name)
    if neg is not None and neg.dim() == 0:
        neg = neg.reshape(1)
    return pos, neg


def pointwise_loss(positive_predictions, negative_predictions=None, mask=None):
    pos, neg = _prepare('pointwise_loss', positive_predictions, negative_predictions, mask)
    return _NativeLoss.apply('pointwise', pos, neg, mask)


def bpr_loss(positive_predictions, negative_predictions, mask=None, ratio=1):
    pos, neg = _prepare('bpr_loss', positive_predictions, negative_predictions, mask)
    return _NativeLoss.apply('bpr', pos, neg, mask)


def hinge_loss(positive_predictions, negative_predictions, mask=None, ratio=1):
    pos, neg = _prepare('hinge_loss', positive_predictions, negative_predictions, mask)
    return _NativeLoss.apply('hinge', pos, neg, mask)


def adaptive_hinge_loss(positive_predictions, negative_predictions, mask=None, ratio=1):
    pos, neg = _prepare('adaptive_hinge_loss', positive_predictions, negative_predictions, mask)
    return _NativeLoss.apply('adaptive_hinge', pos, neg, mask)
