"""Executable numpy specification of the row-sharded step's kernel phases (the API of
recommendation_gans_b200.sharded.CudaShardBackend), for the CPU tests only: it lets the real orchestration
(`ShardedMF.train_steps`) and the real transport (`DistComm` over gloo, world_size 2) run without a GPU.

It is written independently of the CUDA planner: the exchange layout comes from a direct stable grouping of the
step's (table, slot) entries, and the optimiser is the reference's DENSE Adam/SGD over the whole local tables every
step (oracle.mf_oracle.adam_dense_step_numpy) -- so agreement with the single-process oracle shows that the
exchange protocol is right.  Test helper, not product code."""
import numpy as np
import torch

from oracle.mf_oracle import adam_dense_step_numpy, rms_dense_step_numpy, sgd_dense_step_numpy
from recommendation_gans_b200.sharding import shard_range

f32 = np.float32


def _sigmoid(z):
    return (f32(1) / (f32(1) + np.exp(-z.astype(f32)))).astype(f32)


class SpecShardBackend(object):
    def __init__(self, rank, world, num_users, num_items, dim, local_tables, optimizer='adam', lr=1e-3, l2=0.0,
                 betas=(0.5, 0.999), eps=1e-8):
        self.rank, self.world, self.dim = rank, world, dim
        self.num_users, self.num_items = num_users, num_items
        self.tables = [np.array(t, dtype=f32).reshape(len(t), -1).copy() for t in local_tables]   # ue, ie, ub, ib
        self.m = [np.zeros_like(t) for t in self.tables]
        self.v = [np.zeros_like(t) for t in self.tables]
        self.opt, self.lr, self.l2, self.betas, self.eps = optimizer, lr, l2, betas, eps
        self.t = 0
        self.stride = (dim + 3) // 4 * 4 + 4
        self.Dp = self.stride - 4

    # -- buffers (CPU torch tensors so gloo can move them) ------------------------------------------
    def zeros(self, n, dtype):
        return torch.zeros(int(n), dtype=dtype)

    def empty(self, n, dtype):
        return torch.zeros(int(n), dtype=dtype)

    def ids(self, x):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(x), dtype=np.int64))

    # -- plan -----------------------------------------------------------------------------------------
    def plan(self, pos_users, pos_items, batch, n_neg, neg_users, neg_items, step0, nsteps):
        G, r = self.world, self.rank
        pu, pi = pos_users.numpy(), pos_items.numpy()
        nu, ni = neg_users.numpy(), neg_items.numpy()
        m = n_neg * batch
        counts = np.zeros((nsteps, G, G), dtype=np.int64)
        self.steps = []
        for s in range(nsteps):
            first = (step0 + s) * batch
            b = min(batch, len(pu) - first)
            users = np.concatenate([pu[first:first + b], nu[s * m:(s + 1) * m]])
            items = np.concatenate([pi[first:first + b], ni[s * m:(s + 1) * m]])
            if users.max() >= self.num_users or items.max() >= self.num_items or min(users.min(), items.min()) < 0:
                raise ValueError('id out of range')
            comp = np.empty(b + m, dtype=np.int64)            # computing rank of every slot
            for c in range(G):
                lo, hi = shard_range(b, c, G)
                comp[lo:hi] = c
                lo, hi = shard_range(m, c, G)
                comp[b + lo:b + hi] = c
            # entries in (table, slot) order; a block X[o->c] keeps that order
            ent_table = np.concatenate([np.zeros(b + m, np.int64), np.ones(b + m, np.int64)])
            ent_slot = np.concatenate([np.arange(b + m), np.arange(b + m)])
            ent_id = np.concatenate([users, items])
            ent_owner, ent_comp = ent_id % G, comp[ent_slot]
            for o in range(G):
                for c in range(G):
                    counts[s, o, c] = np.sum((ent_owner == o) & (ent_comp == c))
            serve = [k for c in range(G) for k in np.flatnonzero((ent_owner == r) & (ent_comp == c))]
            recv = [k for o in range(G) for k in np.flatnonzero((ent_owner == o) & (ent_comp == r))]
            b_lo, b_hi = shard_range(b, r, G)
            m_lo, m_hi = shard_range(m, r, G)
            my_slots = list(range(b_lo, b_hi)) + list(range(b + m_lo, b + m_hi))
            where = {(int(ent_table[k]), int(ent_slot[k])): p for p, k in enumerate(recv)}
            self.steps.append(dict(
                b=b, m=m, b_loc=b_hi - b_lo, m_lo=m_lo,
                serve_table=ent_table[serve], serve_row=ent_id[serve] // G,
                rpos_u=np.array([where[(0, j)] for j in my_slots], dtype=np.int64),
                rpos_i=np.array([where[(1, j)] for j in my_slots], dtype=np.int64)))
        return counts

    # -- phases ---------------------------------------------------------------------------------------
    def gather(self, s, send):
        st = self.steps[s]
        out = send.numpy()[:len(st['serve_row']) * self.stride].reshape(-1, self.stride)
        for p, (t, row) in enumerate(zip(st['serve_table'], st['serve_row'])):
            out[p, :self.dim] = self.tables[t][row]
            out[p, self.Dp] = self.tables[2 + t][row, 0]

    def _rows(self, s, recv):
        st = self.steps[s]
        rows = recv.numpy()[:2 * len(st['rpos_u']) * self.stride].reshape(-1, self.stride)
        return st, rows[st['rpos_u']], rows[st['rpos_i']]

    def forward(self, loss, s, recv, cell):
        st, ru, ri = self._rows(s, recv)
        dot = np.einsum('ij,ij->i', ru[:, :self.dim], ri[:, :self.dim]).astype(f32)
        st['pred'] = _sigmoid((dot + ru[:, self.Dp]) + ri[:, self.Dp])
        if loss == 'adaptive_hinge':
            neg = st['pred'][st['b_loc']:]
            packed = 0
            if len(neg):
                j = int(np.argmax(neg))
                packed = (int(neg[j:j + 1].view(np.uint32)[0]) << 32) | (0xFFFFFFFF - (st['m_lo'] + j))
            cell[0] = packed

    def backward(self, loss, s, recv, cell, gsend, partial):
        st, ru, ri = self._rows(s, recv)
        pred, b_loc, b, m = st['pred'], st['b_loc'], st['b'], st['m']
        pos, neg = pred[:b_loc], pred[b_loc:]
        dpos, dneg = np.zeros(len(pos), f32), np.zeros(len(neg), f32)
        p0 = p1 = 0.0
        if loss == 'pointwise':
            with np.errstate(divide='ignore'):
                p0 = float((-np.maximum(np.log(pos), f32(-100))).sum(dtype=np.float64))
                p1 = float((-np.maximum(np.log(f32(1) - neg), f32(-100))).sum(dtype=np.float64))
            dpos = (pos - f32(1)) / np.maximum((f32(1) - pos) * pos, f32(1e-12)) / f32(b)
            dneg = neg / np.maximum((f32(1) - neg) * neg, f32(1e-12)) / f32(m)
        elif loss == 'hinge':
            d = neg - pos + f32(1)
            p0 = float(np.maximum(d, 0).sum(dtype=np.float64))
            act = (d >= 0).astype(f32)
            dpos, dneg = -act / f32(b), act / f32(b)
        elif loss == 'bpr':
            sg = _sigmoid(pos - neg)
            p0 = float((f32(1) - sg).sum(dtype=np.float64))
            gs = (sg * (f32(1) - sg)) / f32(b)
            dpos, dneg = -gs, gs
        else:
            packed = int(cell[0])
            gmax = np.array([packed >> 32], dtype=np.uint32).view(f32)[0]
            jstar = 0xFFFFFFFF - (packed & 0xFFFFFFFF)
            d = gmax - pos + f32(1)
            p0 = float(np.maximum(d, 0).sum(dtype=np.float64))
            dpos = -(d >= 0).astype(f32) / f32(b)
            jl = jstar - st['m_lo']
            if 0 <= jl < len(neg):
                dneg[jl] = f32(1)                      # b active positives times 1/b (always active on [0,1])
        dpred = np.concatenate([dpos, dneg]).astype(f32)
        dz = ((dpred * (f32(1) - pred)) * pred).astype(f32)
        out = gsend.numpy()[:2 * len(pred) * self.stride].reshape(-1, self.stride)
        out[st['rpos_u'], :self.dim] = dz[:, None] * ri[:, :self.dim]
        out[st['rpos_u'], self.Dp] = dz
        out[st['rpos_i'], :self.dim] = dz[:, None] * ru[:, :self.dim]
        out[st['rpos_i'], self.Dp] = dz
        partial[0], partial[1] = p0, p1

    def update(self, s, grecv):
        st = self.steps[s]
        rows = grecv.numpy()[:len(st['serve_row']) * self.stride].reshape(-1, self.stride)
        grads = [np.zeros_like(t) for t in self.tables]
        for p, (t, row) in enumerate(zip(st['serve_table'], st['serve_row'])):     # arrival order
            grads[t][row] += rows[p, :self.dim]
            grads[2 + t][row, 0] += rows[p, self.Dp]
        self.t += 1
        for k in range(4):                                                        # dense: every local row steps
            if self.opt == 'adam':
                adam_dense_step_numpy(self.tables[k], self.m[k], self.v[k], grads[k], self.t, self.lr,
                                      self.betas[0], self.betas[1], self.eps, self.l2)
            elif self.opt == 'rms':                   # torch.optim.RMSprop defaults: alpha 0.99, eps 1e-8
                rms_dense_step_numpy(self.tables[k], self.v[k], grads[k], self.lr, 0.99, 1e-8, self.l2)
            else:
                sgd_dense_step_numpy(self.tables[k], grads[k], self.lr, self.l2)

    def flush(self):
        pass

    def local_tables(self):
        return [t.copy() for t in self.tables]

    def close(self):
        pass
