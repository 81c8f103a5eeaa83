"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's implicit-MF fit / predict / evaluate hot path.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / CPU baseline.

Parity status: PINNED by execution of the reference.  The reference ships no tests or
golden vectors, so `oracle/make_golden.py` imports the real reference
from /root/reference in the build container, runs it on seeded inputs and freezes the
outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this restatement
against those vectors (bit-exact for integer streams, <=1e-6 for fp32).

The floating-point arithmetic of the reference lives in PyTorch (present in this image,
torch 2.11.0; the reference pins no version): `F.embedding`, autograd, `nn.BCELoss`,
`torch.optim.Adam/SGD`, called at implicit.py:183-192,348,356,361,363 and
spotlight/losses.py:42-50.  This restatement calls the same torch CPU ops on plain leaf
tensors (no nn.Module, no spotlight code), and additionally restates the optimiser
update rules elementwise in numpy (`adam_dense_step_numpy`, `sgd_dense_step_numpy`) so
the lazy row catch-up used by the CUDA path can be checked against the published rule.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from . import mt19937_ref

FLOAT_MAX = np.finfo(np.float32).max  # spotlight/evaluation.py:9


# ----------------------------------------------------------------------------------------
# network: spotlight/factorization/representations.py:62-91, spotlight/layers.py:23-56
# ----------------------------------------------------------------------------------------
def init_tables(num_users, num_items, dim, torch_seed=None):
    """ScaledEmbedding: normal_(0, 1/dim) (layers.py:35); ZeroEmbedding: zeros (layers.py:54).
    Draw order follows BilinearNet.__init__ (representations.py:47-60): user emb, item emb."""
    if torch_seed is not None:
        torch.manual_seed(torch_seed)
    # nn.Embedding.__init__ first draws its own default init (normal_(0,1)) and the subclass
    # reset_parameters() is what actually runs (it overrides), so one normal_ per table.
    user_emb = torch.empty(num_users, dim).normal_(0, 1.0 / dim)
    item_emb = torch.empty(num_items, dim).normal_(0, 1.0 / dim)
    user_bias = torch.zeros(num_users, 1)
    item_bias = torch.zeros(num_items, 1)
    return user_emb, item_emb, user_bias, item_bias


def bilinear_forward(user_emb, item_emb, user_bias, item_bias, user_ids, item_ids):
    """sigmoid(<U[u],V[i]> + bu[u] + bi[i])   (representations.py:80-91; note the sigmoid, F1)."""
    u = F.embedding(user_ids, user_emb)
    v = F.embedding(item_ids, item_emb)
    bu = F.embedding(user_ids, user_bias).squeeze(-1)
    bi = F.embedding(item_ids, item_bias).squeeze(-1)
    return torch.sigmoid((u * v).sum(1) + bu + bi)


# ----------------------------------------------------------------------------------------
# losses: spotlight/losses.py:20-172 (all operate on probabilities)
# ----------------------------------------------------------------------------------------
def _masked_mean(loss, mask):
    """losses.py:51-55 (and :91-95, :124-128): `mask` zeroes entries; the mean runs over mask.sum()."""
    if mask is not None:
        mask = mask.float()
        return (loss * mask).sum() / mask.sum()
    return loss.mean()


def pointwise_loss(pos, neg=None, mask=None):
    loss = F.binary_cross_entropy(pos, torch.ones_like(pos))          # losses.py:42-44
    if neg is not None:
        loss = loss + F.binary_cross_entropy(neg, torch.zeros_like(neg))  # losses.py:48-50
    if mask is not None:                                              # losses.py:51-55: scalar loss times mask
        mask = mask.float()
        return (loss * mask).sum() / mask.sum()
    return loss


def bpr_loss(pos, neg, mask=None):
    return _masked_mean(1.0 - torch.sigmoid(pos - neg), mask)         # losses.py:88-96


def hinge_loss(pos, neg, mask=None):
    return _masked_mean(torch.clamp(neg - pos + 1.0, 0.0), mask)      # losses.py:121-130


def adaptive_hinge_loss(pos, neg, mask=None):
    highest, _ = torch.max(neg, 0)                                    # losses.py:170 (neg [n, b] -> per-positive max)
    return hinge_loss(pos, highest.squeeze(), mask)                   # losses.py:172


def loss_for_model(name):
    """implicit.py:194-199: only 'pointwise' and 'hinge' are wired by name; everything else
    (including 'bpr') falls through to adaptive_hinge_loss."""
    if name == 'pointwise':
        return pointwise_loss
    if name == 'hinge':
        return hinge_loss
    return adaptive_hinge_loss


LOSS_FUNCTIONS = {'pointwise': pointwise_loss, 'bpr': bpr_loss,
                  'hinge': hinge_loss, 'adaptive_hinge': adaptive_hinge_loss}


# ----------------------------------------------------------------------------------------
# optimisers: spotlight/optimizers.py:4-22 -> torch.optim (dense)
# ----------------------------------------------------------------------------------------
def make_optimizer(params, kind, lr, weight_decay, betas=(0.5, 0.999), eps=1e-8):
    if kind == 'adam':
        return torch.optim.Adam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    if kind == 'sgd':
        return torch.optim.SGD(params, lr=lr, weight_decay=weight_decay)
    if kind == 'rms':                                                 # spotlight/optimizers.py:18-22
        return torch.optim.RMSprop(params, lr=lr, weight_decay=weight_decay)
    raise ValueError(kind)


def adam_scalars(step, lr, beta1, beta2):
    """Python-double scalars of torch/optim/adam.py `_single_tensor_adam` (non-capturable)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    return lr / bc1, bc2 ** 0.5


def adam_dense_step_numpy(p, m, v, g, step, lr, beta1, beta2, eps, wd):
    """Elementwise fp32 restatement of one dense torch Adam step (adam.py: weight decay
    `grad.add(param, alpha=wd)`, `exp_avg.lerp_(grad, 1-beta1)`, `exp_avg_sq.mul_(beta2)
    .addcmul_(grad, grad, value=1-beta2)`, `denom = sqrt(v)/bc2_sqrt + eps`,
    `param.addcdiv_(exp_avg, denom, value=-step_size)`).  In-place on p, m, v (float32)."""
    f = np.float32
    step_size, bc2_sqrt = adam_scalars(step, lr, beta1, beta2)
    g = g.astype(np.float32)
    if wd != 0:
        g = g + f(wd) * p
    w = f(1 - beta1)
    diff = g - m
    if abs(float(w)) < 0.5:                                            # ATen lerp
        m[...] = m + w * diff
    else:
        m[...] = g - diff * (f(1) - w)
    v[...] = v * f(beta2) + (f(1 - beta2) * g) * g
    denom = np.sqrt(v) / f(bc2_sqrt) + f(eps)
    p[...] = p + (f(-step_size) * m) / denom


def rms_dense_step_numpy(p, v, g, lr, alpha, eps, wd):
    """Elementwise fp32 restatement of one dense torch RMSprop step (rmsprop.py _single_tensor_rmsprop, momentum 0, not
    centered): `grad.add(param, alpha=wd)`, `square_avg.mul_(alpha).addcmul_(grad, grad, value=1-alpha)`,
    `avg = square_avg.sqrt().add_(eps)`, `param.addcdiv_(grad, avg, value=-lr)`.  In-place on p, v (float32)."""
    f = np.float32
    g = g.astype(np.float32)
    if wd != 0:
        g = g + f(wd) * p
    v[...] = v * f(alpha) + (f(1 - alpha) * g) * g
    avg = np.sqrt(v) + f(eps)
    p[...] = p + (f(-lr) * g) / avg


def sgd_dense_step_numpy(p, g, lr, wd):
    """torch/optim/sgd.py momentum=0: p <- p - lr*(g + wd*p)."""
    f = np.float32
    g = g.astype(np.float32)
    if wd != 0:
        g = g + f(wd) * p
    p[...] = p + f(-lr) * g


# ----------------------------------------------------------------------------------------
# model driver: implicit.py:163-212 (init), 347-379 (iterations), 381-415 (predict)
# ----------------------------------------------------------------------------------------
class OracleMF:
    def __init__(self, user_emb, item_emb, user_bias, item_bias, loss='pointwise',
                 optimizer='adam', lr=1e-2, l2=0.0, betas=(0.5, 0.999),
                 batch_size=256, num_negative_samples=3, neg_pairs=None, loss_fn=None):
        self.tables = [t.detach().clone().requires_grad_(True)
                       for t in (user_emb, item_emb, user_bias, item_bias)]
        self.num_users, self.num_items = user_emb.shape[0], item_emb.shape[0]
        self.loss_fn = loss_fn if loss_fn is not None else loss_for_model(loss)
        self.opt = make_optimizer(self.tables, optimizer, lr, l2, betas)
        self.batch_size = batch_size
        self.n_neg = num_negative_samples
        self.neg_pairs = None if neg_pairs is None else np.asarray(neg_pairs, dtype=np.int64)

    def forward(self, users, items):
        return bilinear_forward(*self.tables, users, items)

    def draw_negatives(self, gen):
        """implicit.py:352-354: k = num_neg * batch_size pairs (k uses B even on a partial batch)."""
        idx = mt19937_ref.choices_indices(gen, len(self.neg_pairs), self.n_neg * self.batch_size)
        pairs = self.neg_pairs[idx]
        return torch.from_numpy(pairs[:, 0].copy()), torch.from_numpy(pairs[:, 1].copy())

    def train_step(self, users, items, neg_users=None, neg_items=None):
        """implicit.py:347-364."""
        pos = self.forward(users, items)
        self.opt.zero_grad()
        if neg_users is not None:
            neg = self.forward(neg_users, neg_items)
            loss = self.loss_fn(pos, neg)
        else:
            loss = self.loss_fn(pos)
        loss.backward()
        self.opt.step()
        return loss.detach()

    @torch.no_grad()
    def val_step(self, users, items, neg_users=None, neg_items=None):
        """implicit.py:366-379."""
        pos = self.forward(users, items)
        if neg_users is not None:
            return self.loss_fn(pos, self.forward(neg_users, neg_items))
        return self.loss_fn(pos)

    @torch.no_grad()
    def predict(self, user_ids, item_ids=None):
        """implicit.py:381-415 + _components.py:8-25."""
        if item_ids is None:
            item_ids = np.arange(self.num_items, dtype=np.int64)
        item_ids = torch.from_numpy(np.asarray(item_ids, dtype=np.int64).reshape(-1))
        if np.isscalar(user_ids):
            users = torch.full_like(item_ids, int(user_ids))
        else:
            users = torch.from_numpy(np.asarray(user_ids, dtype=np.int64).reshape(-1))
        return self.forward(users, item_ids).numpy().flatten()

    @torch.no_grad()
    def logits(self, user_id):
        """Pre-sigmoid scores of one user for all items (ranking key of the CUDA path)."""
        ue, ie, ub, ib = self.tables
        return ((ie * ue[user_id]).sum(1) + ub[user_id, 0] + ib[:, 0]).numpy()

    def numpy_tables(self):
        return [t.detach().numpy().copy() for t in self.tables]


def fit_epochs(model, users, items, val_users, val_items, n_epochs, gen):
    """Loop structure of implicit.py:279-334 (one pre-shuffled id order for all epochs; the
    validation pass also draws negatives from the same stream).  `users/items` are the already
    shuffled arrays (torch_utils.shuffle is applied by the caller, implicit.py:259-262).
    Returns per-epoch mean train/val losses and the per-step losses."""
    B = model.batch_size
    out = {'train': [], 'val': [], 'train_steps': [], 'val_steps': []}
    tu, ti = torch.from_numpy(users).long(), torch.from_numpy(items).long()
    vu, vi = torch.from_numpy(val_users).long(), torch.from_numpy(val_items).long()
    for _ in range(n_epochs):
        tl = []
        for s in range(0, len(tu), B):
            nu, ni = model.draw_negatives(gen) if model.neg_pairs is not None else (None, None)
            tl.append(model.train_step(tu[s:s + B], ti[s:s + B], nu, ni).item())
        vl = []
        for s in range(0, len(vu), B):
            nu, ni = model.draw_negatives(gen) if model.neg_pairs is not None else (None, None)
            vl.append(model.val_step(vu[s:s + B], vi[s:s + B], nu, ni).item())
        out['train_steps'].append(tl)
        out['val_steps'].append(vl)
        out['train'].append(sum(tl) / len(tl))        # implicit.py:294,300 (python float sum)
        out['val'].append(sum(vl) / len(vl))          # implicit.py:314,320
    return out


# ----------------------------------------------------------------------------------------
# offline negative pairs: spotlight/sampling.py:37-70
# ----------------------------------------------------------------------------------------
def get_negative_samples(users, items, num_users, num_items, num_samples, rs):
    """sampling.py:46-70 on the train pairs (users, items), drawing from the numpy RandomState `rs` in the order the
    reference draws from numpy's global generator: choice(num_users, n), choice(num_items, n), then for every sample
    whose pair is a known interaction (CSR value == 1, interactions.py:159-160) one randint(0, free items, 1) mapped to
    the raw-th item outside the user's row (sampling.py:37-44).  Returns an int64 array [n, 2]."""
    import scipy.sparse as sp
    csr = sp.coo_matrix((np.ones(len(users)), (users, items)), shape=(num_users, num_items)).tocsr()
    csr.sort_indices()
    us = rs.choice(num_users, num_samples)
    its = rs.choice(num_items, num_samples)
    out = np.stack([us, its], 1).astype(np.int64)
    for k in range(num_samples):
        u, i = us[k], its[k]
        if csr[u, i] == 1:
            row = csr.indices[csr.indptr[u]:csr.indptr[u + 1]]
            raw = rs.randint(0, num_items - len(row), size=1)
            out[k, 1] = (raw + np.searchsorted(row - np.arange(len(row)), raw, side='right'))[0]
    return out


# ----------------------------------------------------------------------------------------
# evaluation: spotlight/evaluation.py:108-185
# ----------------------------------------------------------------------------------------
def csr_from_pairs(users, items, num_users, num_items):
    import scipy.sparse as sp
    data = np.ones(len(users))
    return sp.coo_matrix((data, (users, items)), shape=(num_users, num_items)).tocsr()


def topk_stable(logits, rated, kmax):
    """Ranking oracle: descending pre-sigmoid score, ties -> lower item id,
    known train items pushed to the end (evaluation.py:162-169 sets them to FLOAT_MAX of the
    negated score)."""
    key = -logits.astype(np.float32)
    if rated is not None and len(rated):
        key[rated] = FLOAT_MAX
    return np.argsort(key, kind='stable')[:kmax]


def precision_recall_from_topk(topk, targets, ks):
    """evaluation.py:108-113 for each k."""
    tset = set(int(t) for t in targets)
    prec, rec = [], []
    for k in ks:
        hits = len(set(int(x) for x in topk[:k]).intersection(tset))
        prec.append(hits / k)
        rec.append(hits / len(targets))
    return prec, rec


def precision_recall_score(model, test_csr, train_csr=None, k=10, ranking='stable_logit'):
    """evaluation.py:115-185.  ranking='reference' reproduces `(-predict).argsort()` literally;
    'stable_logit' is the tie-defined oracle the CUDA ids are compared with."""
    ks = np.array([k]) if np.isscalar(k) else np.asarray(k)
    precision, recall, cold = [], [], 0
    per_user_topk = {}
    for user_id in range(test_csr.shape[0]):
        row = test_csr[user_id]
        if not len(row.indices):
            continue
        rated = train_csr[user_id].indices if train_csr is not None else None
        if train_csr is not None and not len(rated):
            cold += 1
        if ranking == 'reference':
            pred = -model.predict(user_id)
            if rated is not None:
                pred[rated] = FLOAT_MAX
            order = pred.argsort(axis=0)
        else:
            order = topk_stable(model.logits(user_id), rated, int(ks.max()))
        per_user_topk[user_id] = order[:int(ks.max())]
        p, r = precision_recall_from_topk(order, row.indices, ks)
        precision.append(p)
        recall.append(r)
    precision = np.array(precision).squeeze()
    recall = np.array(recall).squeeze()
    return float(np.mean(precision)), float(np.mean(recall)), cold, per_user_topk


def apk(actual, predicted, k=10):
    """evaluation.py:278-310: average precision at k; a target array without any non-zero id scores 0.0."""
    predicted = predicted[:k] if len(predicted) > k else predicted
    score, num_hits = 0.0, 0.0
    actual_set = set(int(a) for a in actual)
    seen = set()
    for i, p in enumerate(predicted):
        p = int(p)
        if p in actual_set and p not in seen:
            num_hits += 1.0
            score += num_hits / (i + 1.0)
        seen.add(p)
    if not np.asarray(actual).any():
        return 0.0
    return score / min(len(actual), k)


def map_at_k(model, test_csr, k=5, ranking='stable_logit'):
    """evaluation.py:334-353: mean over users with test items of apk(targets, argsort(-predict(user)), k); no train
    mask.  ranking='reference' sorts the sigmoid outputs like the reference, 'stable_logit' is the tie-defined order."""
    vals = []
    for user_id in range(test_csr.shape[0]):
        row = test_csr[user_id]
        if not len(row.indices):
            continue
        if ranking == 'reference':
            order = (-model.predict(user_id)).argsort()
        else:
            order = topk_stable(model.logits(user_id), None, k)
        vals.append(apk(row.indices, order, k=k))
    return float(np.mean(np.array(vals).squeeze()))


def rmse_score(model, user_ids, item_ids):
    """evaluation.py:187-190: sum over the batch of (1 - prediction)^2 (model.test logs sqrt(sum / n) as "BCE")."""
    pred = model.predict(np.asarray(user_ids), np.asarray(item_ids))
    return float(np.sum((1 - pred) ** 2))
