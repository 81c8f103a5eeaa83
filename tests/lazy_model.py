"""Executable specification (numpy, fp32) of the algorithm the CUDA training step implements:
row-sparse traffic with *dense-optimiser semantics* via lazy exact catch-up.

Used by the CPU tests to prove -- before any GPU is involved -- that touching only the rows of
a batch and replaying the missed dense updates (`g = wd * p`) per row reproduces the
reference's dense torch Adam/SGD over whole tables.  Test helper, not product code.
"""
import numpy as np

from oracle.mf_oracle import adam_dense_step_numpy, rms_dense_step_numpy, sgd_dense_step_numpy

f32 = np.float32


def sigmoid32(z):
    return (f32(1) / (f32(1) + np.exp(-z.astype(np.float32)))).astype(np.float32)


def loss_and_dpred(kind, pos, neg):
    """Loss value and dLoss/dpred for 1-D pos [b] and neg [m] probabilities.
    Formulas follow torch's backward of spotlight/losses.py."""
    b, m = len(pos), len(neg)
    dpos = np.zeros(b, f32)
    dneg = np.zeros(m, f32)
    if kind == 'pointwise':
        with np.errstate(divide='ignore'):
            lp = -np.maximum(np.log(pos), f32(-100))
        with np.errstate(divide='ignore'):
            ln = -np.maximum(np.log(f32(1) - neg), f32(-100))     # torch: (1 - x).log(), clamped at -100
        loss = lp.mean(dtype=np.float64) + ln.mean(dtype=np.float64)
        dpos = ((pos - f32(1)) / np.maximum((f32(1) - pos) * pos, f32(1e-12)) / f32(b)).astype(f32)
        dneg = (neg / np.maximum((f32(1) - neg) * neg, f32(1e-12)) / f32(m)).astype(f32)
    elif kind == 'hinge':
        assert b == m
        d = neg - pos + f32(1)
        loss = np.maximum(d, 0).mean(dtype=np.float64)
        act = (d >= 0).astype(f32)
        dneg = act / f32(b)
        dpos = -act / f32(b)
    elif kind == 'bpr':
        assert b == m
        s = sigmoid32(pos - neg)
        loss = (f32(1) - s).mean(dtype=np.float64)
        gs = (s * (f32(1) - s)) / f32(b)
        dpos = -gs
        dneg = gs
    elif kind == 'adaptive_hinge':
        j = int(np.argmax(neg))            # first maximal index, as torch.max(neg, 0)
        d = neg[j] - pos + f32(1)
        loss = np.maximum(d, 0).mean(dtype=np.float64)
        act = (d >= 0).astype(f32)
        dpos = -act / f32(b)
        dneg[j] = act.sum(dtype=np.float64) / b
    else:
        raise ValueError(kind)
    return float(loss), dpos.astype(f32), dneg.astype(f32)


class LazyMF:
    def __init__(self, tables, optimizer, lr, wd, betas=(0.5, 0.999), eps=1e-8):
        self.ue, self.ie, self.ub, self.ib = [t.astype(np.float32).copy() for t in tables]
        self.opt, self.lr, self.wd, self.betas, self.eps = optimizer, lr, wd, betas, eps
        self.t = 0
        self.state = {}
        for name in ('ue', 'ie', 'ub', 'ib'):
            p = getattr(self, name)
            self.state[name] = (np.zeros_like(p), np.zeros_like(p))
        self.last_u = np.zeros(self.ue.shape[0], np.int64)   # step each row is current for
        self.last_i = np.zeros(self.ie.shape[0], np.int64)

    # one dense-equivalent update of a single row of one table with gradient g at step s
    def _row_step(self, name, r, g, s):
        p = getattr(self, name)
        m, v = self.state[name]
        if self.opt == 'adam':
            adam_dense_step_numpy(p[r], m[r], v[r], g, s, self.lr, self.betas[0], self.betas[1], self.eps, self.wd)
        elif self.opt == 'rms':                      # torch.optim.RMSprop defaults: alpha 0.99, eps 1e-8
            rms_dense_step_numpy(p[r], v[r], g, self.lr, 0.99, 1e-8, self.wd)
        else:
            sgd_dense_step_numpy(p[r], g, self.lr, self.wd)

    def _catch_up(self, table, r, upto):
        last = self.last_u if table == 'u' else self.last_i
        names = ('ue', 'ub') if table == 'u' else ('ie', 'ib')
        for s in range(int(last[r]) + 1, upto + 1):
            for n in names:
                self._row_step(n, r, np.zeros_like(getattr(self, n)[r]), s)
        last[r] = max(last[r], upto)

    def forward(self, users, items):
        dot = np.einsum('ij,ij->i', self.ue[users], self.ie[items]).astype(f32)
        return sigmoid32(dot + self.ub[users, 0] + self.ib[items, 0])

    def train_step(self, kind, pu, pi, nu, ni):
        self.t += 1
        t = self.t
        users = np.concatenate([pu, nu])
        items = np.concatenate([pi, ni])
        for r in np.unique(users):
            self._catch_up('u', r, t - 1)
        for r in np.unique(items):
            self._catch_up('i', r, t - 1)
        pred = self.forward(users, items)
        b = len(pu)
        loss, dpos, dneg = loss_and_dpred(kind, pred[:b], pred[b:])
        dpred = np.concatenate([dpos, dneg])
        dz = ((dpred * (f32(1) - pred)) * pred).astype(f32)      # sigmoid backward
        gu = {}
        gi = {}
        old_ue, old_ie = self.ue.copy(), self.ie.copy()
        for j in range(len(users)):                               # slot order accumulation
            u, i = int(users[j]), int(items[j])
            a = gu.setdefault(u, [np.zeros(self.ue.shape[1], f32), f32(0)])
            a[0] = a[0] + dz[j] * old_ie[i]
            a[1] = f32(a[1] + dz[j])
            c = gi.setdefault(i, [np.zeros(self.ie.shape[1], f32), f32(0)])
            c[0] = c[0] + dz[j] * old_ue[u]
            c[1] = f32(c[1] + dz[j])
        for u, (g, gb) in gu.items():
            self._row_step('ue', u, g, t)
            self._row_step('ub', u, np.array([gb], f32), t)
            self.last_u[u] = t
        for i, (g, gb) in gi.items():
            self._row_step('ie', i, g, t)
            self._row_step('ib', i, np.array([gb], f32), t)
            self.last_i[i] = t
        return loss

    def flush(self):
        for r in range(self.ue.shape[0]):
            self._catch_up('u', r, self.t)
        for r in range(self.ie.shape[0]):
            self._catch_up('i', r, self.t)

    def tables(self):
        return [self.ue, self.ie, self.ub, self.ib]
