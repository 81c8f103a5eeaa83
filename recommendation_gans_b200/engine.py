"""Host-side owner of one native model handle (include/mfb200.h).

`MFEngine` borrows the four BilinearNet tables (torch `nn.Parameter` storage) and the torch
optimiser's hyper-parameters/state tensors, and exposes the fused CUDA path: negative draws,
training steps, validation losses, predict, top-k evaluation.  No torch ops run on this path;
torch supplies device memory and the current stream only.
"""
import ctypes
import os
import random as _py_random

import numpy as np
import torch

from . import _native as N


def _as_i64_cuda(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.int64)).to(device)


# ---------------------------------------------------------------------------------------------
# MT19937 state plumbing: the device sampler continues the HOST generator objects' streams, so the
# global `random` module / a numpy RandomState end up exactly where the reference leaves them.
# ---------------------------------------------------------------------------------------------
def _get_py_state(rng):
    version, internal, gauss = rng.getstate()
    return np.array(internal, dtype=np.uint32), (version, gauss)


def _set_py_state(rng, arr, extra):
    rng.setstate((extra[0], tuple(int(v) for v in arr), extra[1]))


def _get_np_state(rs):
    name, key, pos, has_gauss, cached = rs.get_state()
    if name != 'MT19937':
        raise NotImplementedError('only MT19937 RandomState streams are supported')
    arr = np.empty(625, dtype=np.uint32)
    arr[:624] = key
    arr[624] = pos
    return arr, (has_gauss, cached)


def _set_np_state(rs, arr, extra):
    rs.set_state(('MT19937', arr[:624].copy(), int(arr[624]), extra[0], extra[1]))


def sample_items_device(num_items, count, random_state, device=None):
    """spotlight/sampling.py:33 on the GPU: continues `random_state`'s MT19937 stream."""
    N.require_cuda()
    lib = N.load_library()
    device = device or torch.device('cuda', torch.cuda.current_device())
    out = torch.empty(int(count), dtype=torch.int64, device=device)
    state, extra = _get_np_state(random_state)
    N.check(lib.mfb_mt_sample_items(N.hptr(state), int(num_items), int(count), N.dptr(out), N.stream_ptr()),
            'sample_items')
    _set_np_state(random_state, state, extra)
    return out


def negative_pairs_device(key_csr, row_csr, num_users, num_items, num_samples, random_state, device=None):
    """spotlight/sampling.py:46-70 on the GPU, continuing `random_state`'s MT19937 stream (numpy legacy RandomState, e.g.
    np.random.mtrand._rand for the global generator).  key_csr: entries for which has_key is true; row_csr: all stored
    entries (scipy CSR, sorted indices).  Returns (users, items) int64 CUDA tensors and the number of re-drawn pairs."""
    N.require_cuda()
    lib = N.load_library()
    device = device or torch.device('cuda', torch.cuda.current_device())
    out_u = torch.empty(int(num_samples), dtype=torch.int64, device=device)
    out_i = torch.empty(int(num_samples), dtype=torch.int64, device=device)

    def dev(csr):
        return (torch.from_numpy(csr.indptr.astype(np.int64)).to(device),
                torch.from_numpy(csr.indices.astype(np.int32)).to(device))
    k_ptr, k_idx = dev(key_csr)
    r_ptr, r_idx = dev(row_csr)
    state, extra = _get_np_state(random_state)
    redrawn = ctypes.c_int64(0)
    with torch.cuda.device(device):
        N.check(lib.mfb_negative_pairs(N.hptr(state), int(num_users), int(num_items), int(num_samples), N.dptr(k_ptr),
                                       N.dptr(k_idx), N.dptr(r_ptr), N.dptr(r_idx), N.dptr(out_u), N.dptr(out_i),
                                       ctypes.byref(redrawn), N.stream_ptr()), 'get_negative_samples')
    _set_np_state(random_state, state, extra)
    return out_u, out_i, int(redrawn.value)


def choices_indices_device(pop_len, k, rng=None, device=None):
    """Index stream of random.choices(range(pop_len), k=k), continuing `rng` (default: global random)."""
    N.require_cuda()
    lib = N.load_library()
    rng = rng or _py_random
    device = device or torch.device('cuda', torch.cuda.current_device())
    out = torch.empty(int(k), dtype=torch.int64, device=device)
    state, extra = _get_py_state(rng)
    N.check(lib.mfb_mt_choices_indices(N.hptr(state), int(pop_len), int(k), N.dptr(out), N.stream_ptr()), 'choices')
    _set_py_state(rng, state, extra)
    return out


def draw_negative_pairs_device(pop_users, pop_items, k, rng=None):
    """random.choices(neg_examples, k=k) (implicit.py:352) on the device, continuing `rng` (default: global random);
    pop_users / pop_items: the pair list as two int64 CUDA tensors.  Model-less: also serves `representation=` modules
    that train on the torch-autograd step."""
    N.require_cuda()
    lib = N.load_library()
    rng = rng or _py_random
    out_u = torch.empty(int(k), dtype=torch.int64, device=pop_users.device)
    out_i = torch.empty(int(k), dtype=torch.int64, device=pop_users.device)
    if k == 0:
        return out_u, out_i
    state, extra = _get_py_state(rng)
    with torch.cuda.device(pop_users.device):
        N.check(lib.mfb_mt_choices_pairs(N.hptr(state), N.dptr(pop_users), N.dptr(pop_items), pop_users.numel(), int(k),
                                         N.dptr(out_u), N.dptr(out_i), N.stream_ptr()), 'draw_negative_pairs')
    _set_py_state(rng, state, extra)
    return out_u, out_i


def topk_scores_device(scores, k, user_ids=None, train_indptr=None, train_indices=None, with_scores=False):
    """Top-k item ids of every row of a dense CUDA score matrix [n_rows, n_items] (descending score, ties -> lower id,
    train items of the row's user last) -- mfb_topk_scores; the ranking kernel for models whose scores come from a
    torch module."""
    N.require_cuda()
    lib = N.load_library()
    scores = scores.detach().to(torch.float32).contiguous()
    n_rows, n_items = scores.shape
    ids = torch.empty((n_rows, int(k)), dtype=torch.int32, device=scores.device)
    out = torch.empty((n_rows, int(k)), dtype=torch.float32, device=scores.device) if with_scores else None
    scratch = torch.empty(2 * n_rows, dtype=torch.float32, device=scores.device) if k > 256 else None
    if user_ids is not None:
        user_ids = _as_i64_cuda(user_ids, scores.device)
    with torch.cuda.device(scores.device):
        N.check(lib.mfb_topk_scores(N.dptr(scores), n_rows, n_items, N.dptr(user_ids), N.dptr(train_indptr),
                                    N.dptr(train_indices), int(k), N.dptr(ids), N.dptr(out), N.dptr(scratch),
                                    N.stream_ptr()), 'topk_scores')
    return (ids, out) if with_scores else ids


def topk_hits_device(topk_ids, user_ids, test_indptr, test_indices, ks):
    """_get_precision_recall (evaluation.py:108-113) for every row of topk_ids: (hits [n, len(ks)], ntargets [n])."""
    N.require_cuda()
    lib = N.load_library()
    user_ids = _as_i64_cuda(user_ids, topk_ids.device)
    n, k = topk_ids.shape
    ks_arr = np.ascontiguousarray(ks, dtype=np.int32)
    hits = torch.empty((n, len(ks_arr)), dtype=torch.int32, device=topk_ids.device)
    ntargets = torch.empty(n, dtype=torch.int32, device=topk_ids.device)
    with torch.cuda.device(topk_ids.device):
        N.check(lib.mfb_topk_hits(N.dptr(topk_ids), N.dptr(user_ids), n, int(k), N.dptr(test_indptr),
                                  N.dptr(test_indices), N.hptr(ks_arr), len(ks_arr), N.dptr(hits), N.dptr(ntargets),
                                  N.stream_ptr()), 'topk_hits')
    return hits, ntargets


def mt_words_device(state625, nwords, device=None):
    N.require_cuda()
    lib = N.load_library()
    device = device or torch.device('cuda', torch.cuda.current_device())
    out = torch.empty(int(nwords), dtype=torch.int32, device=device)
    N.check(lib.mfb_mt_words(N.hptr(state625), int(nwords), N.dptr(out), N.stream_ptr()), 'mt_words')
    return out


def mt_words_device_parallel(state625, nwords, words_per_step, device=None):
    """mt_words_device through the multi-CTA generator (GF(2) jump-ahead, csrc/mfb_mt_jump.cu): same words, same state."""
    N.require_cuda()
    lib = N.load_library()
    device = device or torch.device('cuda', torch.cuda.current_device())
    out = torch.empty(int(nwords), dtype=torch.int32, device=device)
    N.check(lib.mfb_mt_words_parallel(N.hptr(state625), int(nwords), int(words_per_step), N.dptr(out), N.stream_ptr()),
            'mt_words_parallel')
    return out


def loss_forward_backward(kind, pos, neg, need_grad, mask=None):
    """spotlight/losses.py on probability tensors -> (loss, dpos, dneg).  pos is 1-D [b]; neg is 1-D, or [n, b] for
    adaptive_hinge (per-positive maximum over dim 0); mask is None or [b] (loss*mask summed over mask.sum())."""
    N.require_cuda()
    lib = N.load_library()
    pos = pos.detach().contiguous().float()
    neg_c = None if neg is None else neg.detach().contiguous().float()
    loss = torch.empty((), dtype=torch.float32, device=pos.device)
    dpos = torch.empty_like(pos) if need_grad else None
    dneg = (torch.empty_like(neg_c) if need_grad else None) if neg_c is not None else None
    with torch.cuda.device(pos.device):
        if mask is None and (neg_c is None or neg_c.dim() <= 1):
            N.check(lib.mfb_loss_forward_backward(N.LOSS[kind], N.dptr(pos), pos.numel(), N.dptr(neg_c),
                                                  0 if neg_c is None else neg_c.numel(), N.dptr(loss), N.dptr(dpos),
                                                  N.dptr(dneg), N.stream_ptr()), kind + '_loss')
        else:
            mask_c = None if mask is None else mask.detach().to(device=pos.device, dtype=torch.float32).contiguous()
            if mask_c is not None and mask_c.shape != pos.shape:
                raise RuntimeError('%s_loss: mask of shape %s does not match the %d positive predictions'
                                   % (kind, tuple(mask_c.shape), pos.numel()))
            rows, cols = (0, 0) if neg_c is None else ((0, neg_c.numel()) if neg_c.dim() <= 1 else neg_c.shape)
            N.check(lib.mfb_loss_forward_backward_ex(N.LOSS[kind], N.dptr(pos), pos.numel(), N.dptr(neg_c), int(rows),
                                                     int(cols), N.dptr(mask_c), N.dptr(loss), N.dptr(dpos),
                                                     N.dptr(dneg), N.stream_ptr()), kind + '_loss')
    return loss, dpos, dneg


class MFEngine(object):
    """Native handle bound to a BilinearNet and (optionally) a torch optimiser.

    optimizer=None gives a forward-only engine (predict / evaluate)."""

    def __init__(self, net, optimizer=None, fast_math=None):
        """fast_math: MUFU sqrt/rcp (flush-to-zero) in the Adam arithmetic instead of IEEE sqrt/div.  Both modes
        pass the 1e-5 parity tests; the default (None) follows $MFB_FAST_MATH, which defaults to 1."""
        N.require_cuda()
        if fast_math is None:
            fast_math = os.environ.get('MFB_FAST_MATH', '1') != '0'
        self.fast_math = bool(fast_math)
        self._lib = N.load_library()
        self._handle = ctypes.c_void_p(0)
        ue, ie = net.user_embeddings.weight, net.item_embeddings.weight
        ub, ib = net.user_biases.weight, net.item_biases.weight
        for name, p in (('user_embeddings', ue), ('item_embeddings', ie), ('user_biases', ub), ('item_biases', ib)):
            if not p.is_cuda:
                raise RuntimeError('mfb200: %s must live on a CUDA device (no CPU path exists)' % name)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError('mfb200: %s must be contiguous float32' % name)
        if ue.shape[1] != ie.shape[1] or ub.shape != (ue.shape[0], 1) or ib.shape != (ie.shape[0], 1):
            raise ValueError('mfb200: BilinearNet tables have inconsistent shapes')
        self.device = ue.device
        self.num_users, self.num_items, self.dim = ue.shape[0], ie.shape[0], ue.shape[1]
        self._params = (ue, ie, ub, ib)      # keep storage alive while the handle borrows it
        self._optimizer = optimizer
        self._state_tensors = []
        desc = N.ModelDesc()
        desc.num_users, desc.num_items, desc.dim = self.num_users, self.num_items, self.dim
        desc.d_user_emb, desc.d_item_emb = ue.data_ptr(), ie.data_ptr()
        desc.d_user_bias, desc.d_item_bias = ub.data_ptr(), ib.data_ptr()
        desc.fast_math = 1 if fast_math else 0
        kind, hyper = self._parse_optimizer(optimizer)
        desc.optimizer = kind
        desc.lr, desc.beta1, desc.beta2 = hyper['lr'], hyper['beta1'], hyper['beta2']
        desc.eps, desc.weight_decay = hyper['eps'], hyper['weight_decay']
        self.optimizer_kind = kind
        if kind == N.OPT_RMSPROP:
            sq = []
            for p in (ue, ie, ub, ib):
                st = optimizer.state[p]
                if 'square_avg' not in st:     # same layout torch.optim.RMSprop creates lazily
                    st['step'] = torch.tensor(0.0, dtype=torch.float32)
                    st['square_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                sq.append(st['square_avg'])
                self._state_tensors.append(st['square_avg'])
            desc.d_user_emb_v, desc.d_item_emb_v = sq[0].data_ptr(), sq[1].data_ptr()
            desc.d_user_bias_v, desc.d_item_bias_v = sq[2].data_ptr(), sq[3].data_ptr()
        if kind == N.OPT_ADAM:
            moments = []
            for p in (ue, ie, ub, ib):
                st = optimizer.state[p]
                if 'exp_avg' not in st:       # same layout torch.optim.Adam creates lazily
                    st['step'] = torch.tensor(0.0, dtype=torch.float32)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                moments.append((st['exp_avg'], st['exp_avg_sq']))
                self._state_tensors += [st['exp_avg'], st['exp_avg_sq']]
            (desc.d_user_emb_m, desc.d_user_emb_v) = (moments[0][0].data_ptr(), moments[0][1].data_ptr())
            (desc.d_item_emb_m, desc.d_item_emb_v) = (moments[1][0].data_ptr(), moments[1][1].data_ptr())
            (desc.d_user_bias_m, desc.d_user_bias_v) = (moments[2][0].data_ptr(), moments[2][1].data_ptr())
            (desc.d_item_bias_m, desc.d_item_bias_v) = (moments[3][0].data_ptr(), moments[3][1].data_ptr())
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            N.check(self._lib.mfb_model_create(ctypes.byref(desc), ctypes.byref(self._handle)), 'model_create')
        if kind in (N.OPT_ADAM, N.OPT_RMSPROP):
            start = int(float(optimizer.state[ue]['step']))
            if start:
                self._call('mfb_model_set_step', start)

    # -- optimiser introspection (implicit.py:182-192 builds it through optimizer_func) --------
    @staticmethod
    def _parse_optimizer(opt):
        hyper = dict(lr=0.0, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0)
        if opt is None:
            return N.OPT_SGD, hyper
        if len(opt.param_groups) != 1:
            raise NotImplementedError('mfb200: a single parameter group is supported')
        g = opt.param_groups[0]
        if g.get('maximize', False) or g.get('differentiable', False):
            raise NotImplementedError('mfb200: maximize/differentiable optimisers are not supported')
        hyper['lr'] = float(g['lr'])
        hyper['weight_decay'] = float(g.get('weight_decay', 0.0))
        if isinstance(opt, torch.optim.Adam) and not isinstance(opt, torch.optim.AdamW):
            if g.get('amsgrad', False) or g.get('decoupled_weight_decay', False):
                raise NotImplementedError('mfb200: amsgrad / decoupled weight decay are not supported')
            hyper['beta1'], hyper['beta2'] = float(g['betas'][0]), float(g['betas'][1])
            hyper['eps'] = float(g['eps'])
            return N.OPT_ADAM, hyper
        if isinstance(opt, torch.optim.RMSprop):
            if g.get('momentum', 0) != 0 or g.get('centered', False):
                raise NotImplementedError('mfb200: RMSprop momentum / centered are not supported')
            hyper['beta2'] = float(g['alpha'])          # the C ABI carries alpha in beta2
            hyper['eps'] = float(g['eps'])
            return N.OPT_RMSPROP, hyper
        if isinstance(opt, torch.optim.SGD):
            if g.get('momentum', 0) != 0 or g.get('dampening', 0) != 0 or g.get('nesterov', False):
                raise NotImplementedError('mfb200: SGD momentum/dampening/nesterov are not supported')
            return N.OPT_SGD, hyper
        raise NotImplementedError('mfb200: optimiser %s is outside the fused path (Adam, RMSprop and plain SGD '
                                  'are supported)' % type(opt).__name__)

    def _call(self, name, *args):
        if not self._handle:
            raise RuntimeError('mfb200: engine already closed')
        with torch.cuda.device(self.device):
            N.check(getattr(self._lib, name)(self._handle, *args), name)

    def close(self):
        if getattr(self, '_handle', None):
            self._lib.mfb_model_destroy(self._handle)
            self._handle = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __deepcopy__(self, memo):
        return None    # a copied net gets its own (forward-only) engine lazily

    @property
    def step(self):
        return int(self._lib.mfb_model_step(self._handle))

    def _sync_optimizer_step(self):
        if self._optimizer is not None and self.optimizer_kind in (N.OPT_ADAM, N.OPT_RMSPROP):
            t = float(self.step)
            for p in self._params:
                self._optimizer.state[p]['step'] = torch.tensor(t, dtype=torch.float32)

    # -- negatives ----------------------------------------------------------------------------
    def draw_negative_pairs(self, pop_users, pop_items, k, rng=None):
        """random.choices(neg_examples, k=k) (implicit.py:352) continuing `rng` (default: global random)."""
        return draw_negative_pairs_device(pop_users, pop_items, k, rng)

    # -- training / validation ------------------------------------------------------------------
    def _steps(self, fn, loss, pos_users, pos_items, batch, n_neg, neg_users, neg_items):
        pos_users = _as_i64_cuda(pos_users, self.device)
        pos_items = _as_i64_cuda(pos_items, self.device)
        n_pos = pos_users.numel()
        nsteps = (n_pos + batch - 1) // batch
        if n_neg > 0:
            neg_users = _as_i64_cuda(neg_users, self.device)
            neg_items = _as_i64_cuda(neg_items, self.device)
            need = nsteps * n_neg * batch
            if neg_users.numel() < need or neg_items.numel() < need:
                raise ValueError('need %d negative pairs, got %d' % (need, neg_users.numel()))
        else:
            neg_users = neg_items = None
        losses = torch.empty(nsteps, dtype=torch.float32, device=self.device)
        self._call(fn, N.LOSS[loss], N.dptr(pos_users), N.dptr(pos_items), n_pos, int(batch), int(n_neg),
                   N.dptr(neg_users), N.dptr(neg_items), N.dptr(losses), N.stream_ptr())
        return losses

    def train_steps(self, loss, pos_users, pos_items, batch, n_neg, neg_users=None, neg_items=None):
        out = self._steps('mfb_train_steps', loss, pos_users, pos_items, batch, n_neg, neg_users, neg_items)
        self._sync_optimizer_step()
        return out

    def loss_steps(self, loss, pos_users, pos_items, batch, n_neg, neg_users=None, neg_items=None):
        return self._steps('mfb_loss_steps', loss, pos_users, pos_items, batch, n_neg, neg_users, neg_items)

    # -- whole epochs with on-device negative sampling (no host round trip per draw) ---------------
    def rng_seed(self, rng=None):
        """Continue `rng`'s (default: the global `random` module's) MT19937 stream on the device."""
        rng = rng or _py_random
        state, self._rng_extra = _get_py_state(rng)
        self._call('mfb_model_rng_seed', N.hptr(state), N.stream_ptr())

    def rng_sync(self, rng=None):
        """Leave `rng` where the reference's draws would have left it (synchronises)."""
        rng = rng or _py_random
        state = np.empty(625, dtype=np.uint32)
        self._call('mfb_model_rng_state', N.hptr(state), N.stream_ptr())
        _set_py_state(rng, state, self._rng_extra)

    def _epoch(self, fn, loss, pos_users, pos_items, batch, n_neg, pop_users, pop_items):
        pos_users = _as_i64_cuda(pos_users, self.device)
        pos_items = _as_i64_cuda(pos_items, self.device)
        n_pos = pos_users.numel()
        nsteps = (n_pos + batch - 1) // batch
        losses = torch.empty(nsteps, dtype=torch.float32, device=self.device)
        pop_len = 0 if (pop_users is None or n_neg == 0) else pop_users.numel()
        self._call(fn, N.LOSS[loss], N.dptr(pos_users), N.dptr(pos_items), n_pos, int(batch), int(n_neg),
                   N.dptr(pop_users), N.dptr(pop_items), pop_len, N.dptr(losses), N.stream_ptr())
        return losses

    def train_epoch(self, loss, pos_users, pos_items, batch, n_neg, pop_users=None, pop_items=None):
        out = self._epoch('mfb_train_epoch', loss, pos_users, pos_items, batch, n_neg, pop_users, pop_items)
        self._sync_optimizer_step()
        return out

    def loss_epoch(self, loss, pos_users, pos_items, batch, n_neg, pop_users=None, pop_items=None):
        return self._epoch('mfb_loss_epoch', loss, pos_users, pos_items, batch, n_neg, pop_users, pop_items)

    def train_epoch_host(self, loss, pos_users_host, pos_items_host, batch, n_neg, pop_users=None, pop_items=None,
                         rng=None):
        """End-to-end entry: HOST int64 id arrays in, per-step losses (numpy) out; H2D/D2H inside."""
        rng = rng or _py_random
        pu = np.ascontiguousarray(pos_users_host, dtype=np.int64)
        pi = np.ascontiguousarray(pos_items_host, dtype=np.int64)
        nsteps = (len(pu) + batch - 1) // batch
        losses = np.empty(nsteps, dtype=np.float32)
        state, extra = _get_py_state(rng)
        self._call('mfb_train_epoch_host', N.LOSS[loss], N.hptr(pu), N.hptr(pi), len(pu), int(batch), int(n_neg),
                   N.hptr(state), N.dptr(pop_users), N.dptr(pop_items),
                   0 if pop_users is None else pop_users.numel(), N.hptr(losses), N.stream_ptr())
        _set_py_state(rng, state, extra)
        self._sync_optimizer_step()
        return losses

    # -- measurement hooks ------------------------------------------------------------------------
    def profile(self, on=True):
        self._call('mfb_profile_enable', 1 if on else 0)

    def profile_read(self):
        """{kernel class: (summed device ms, launches)} since the last read."""
        ms = np.zeros(N.PROFILE_CLASSES, dtype=np.float64)
        cnt = np.zeros(N.PROFILE_CLASSES, dtype=np.int64)
        self._call('mfb_profile_read', N.hptr(ms), N.hptr(cnt))
        return {self._lib.mfb_profile_name(c).decode(): (float(ms[c]), int(cnt[c]))
                for c in range(N.PROFILE_CLASSES) if cnt[c]}

    @property
    def launches(self):
        return int(self._lib.mfb_model_launches(self._handle))

    def flush(self):
        self._call('mfb_flush', N.stream_ptr())

    # -- predict / evaluate -----------------------------------------------------------------------
    def predict_pairs(self, users, items):
        users = _as_i64_cuda(users, self.device)
        items = _as_i64_cuda(items, self.device)
        if users.numel() != items.numel():
            raise ValueError('user and item id arrays differ in length')
        out = torch.empty(users.numel(), dtype=torch.float32, device=self.device)
        self._call('mfb_predict_pairs', N.dptr(users), N.dptr(items), users.numel(), N.dptr(out), N.stream_ptr())
        return out

    def predict_user(self, user):
        out = torch.empty(self.num_items, dtype=torch.float32, device=self.device)
        self._call('mfb_predict_user', int(user), N.dptr(out), N.stream_ptr())
        return out

    def topk(self, user_ids, k, train_indptr=None, train_indices=None, with_scores=False, plan_key=0):
        """plan_key != 0: the caller promises that this key always comes with the same user list and train CSR; the
        model-independent train-mask preprocessing is then reused between calls (mfb_topk_keyed)."""
        user_ids = _as_i64_cuda(user_ids, self.device)
        n = user_ids.numel()
        ids = torch.empty((n, k), dtype=torch.int32, device=self.device)
        scores = torch.empty((n, k), dtype=torch.float32, device=self.device) if with_scores else None
        self._call('mfb_topk_keyed', N.dptr(user_ids), n, N.dptr(train_indptr), N.dptr(train_indices), int(k),
                   N.dptr(ids), N.dptr(scores), ctypes.c_uint64(int(plan_key) & 0xFFFFFFFFFFFFFFFF), N.stream_ptr())
        return (ids, scores) if with_scores else ids

    @property
    def topk_last_redo(self):
        return int(self._lib.mfb_topk_last_redo(self._handle))

    def debug_tc_stats(self, n_users):
        out = np.zeros(5, dtype=np.int64)
        self._call('mfb_debug_tc_stats', int(n_users), N.hptr(out), N.stream_ptr())
        return dict(users=int(out[0]), candidates=int(out[1]), max=int(out[2]), over_cap=int(out[3]),
                    rescored=int(out[4]))

    def debug_tc_scores(self, user_ids):
        """Raw tensor-core scores [num_items, n_users_padded] (test hook)."""
        user_ids = _as_i64_cuda(user_ids, self.device)
        n = user_ids.numel()
        npad = (n + 255) // 256 * 256
        out = torch.zeros((self.num_items, npad), dtype=torch.float32, device=self.device)
        self._call('mfb_debug_tc_scores', N.dptr(user_ids), n, N.dptr(out), N.stream_ptr())
        return out[:, :n]

    def rank_test_items(self, user_ids, test_indptr, test_indices, train_indptr=None, train_indices=None):
        """Average ranks (scipy rankdata of -predict(user), train items last) of the listed users' test items, aligned
        with test_indices (evaluation.py:52-58)."""
        user_ids = _as_i64_cuda(user_ids, self.device)
        out = torch.zeros(test_indices.numel(), dtype=torch.float32, device=self.device)
        self._call('mfb_rank_test_items', N.dptr(user_ids), user_ids.numel(), N.dptr(test_indptr), N.dptr(test_indices),
                   N.dptr(train_indptr), N.dptr(train_indices), N.dptr(out), N.stream_ptr())
        return out

    def topk_hits(self, topk_ids, user_ids, test_indptr, test_indices, ks):
        return topk_hits_device(topk_ids, user_ids, test_indptr, test_indices, ks)
