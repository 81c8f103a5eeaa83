"""Row-sharded implicit-MF training for catalogs whose tables are split over the GPUs of one box
(BASELINE cfg5: 10M users x 2M items x 128).

One process per GPU.  Rank r owns the rows {g : g % world == r} of the four BilinearNet tables (local index
g // world) with their Adam moments, and computes a contiguous block of every minibatch.  A step is
run_train_iteration (implicit.py:347-364) distributed over the ranks:

    owner    gather requested rows        -> all-to-all of rows (NVLink)      -> computing rank
    compute  forward, [all-reduce MAX of the adaptive-hinge maximum], backward
    compute  gradient rows                -> all-to-all back                  -> owner: ordered reduction + optimiser

All ranks hold the same ids (positives replicated; negatives drawn from the same MT19937 stream, so the sampled
indices stay bit-exact with the reference's `random.choices`), hence ids are never exchanged and the per-pair
row counts are known to every rank from the plan.  The kernels live behind the C ABI (`mfb_shard_*`,
include/mfb200.h); torch.distributed (NCCL) moves the buffers.

`ShardedMF.train_steps` is the same code for every transport: `DistComm` (torch.distributed: NCCL on GPUs, gloo
in the CPU tests) or `LocalComm` (G virtual ranks as threads of one process sharing one GPU and one stream -- how the
1-GPU parity tests exercise G > 1; the direct exchange then runs phase by phase in lockstep so that no flag wait is
ever launched before its signals).  The kernel backend is `CudaShardBackend`; there is no CPU product path (the CPU tests
inject tests/shard_spec_backend.py, an executable numpy specification, to check the exchange logic under gloo).
"""
import ctypes
import threading

import numpy as np
import torch

from . import _native as N


def local_rows(n, rank, world):
    """Rows of an n-row table owned by `rank` (rows g with g % world == rank)."""
    return (int(n) - rank + world - 1) // world


def slice_tables(tables, rank, world):
    """Local tables of `rank` from full [U,D],[I,D],[U,1],[I,1] tables."""
    return [t[rank::world] for t in tables]


def assemble_tables(per_rank_tables, world):
    """Inverse of slice_tables: full tables from the ranks' local ones (numpy arrays)."""
    out = []
    for k in range(4):
        parts = [np.asarray(per_rank_tables[r][k]) for r in range(world)]
        n = sum(p.shape[0] for p in parts)
        full = np.empty((n,) + parts[0].shape[1:], dtype=parts[0].dtype)
        for r in range(world):
            full[r::world] = parts[r]
        out.append(full)
    return out


# ---------------------------------------------------------------------------------------------
# transports
# ---------------------------------------------------------------------------------------------
class DistComm(object):
    """torch.distributed transport (NCCL over NVLink on GPUs; gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_to_all(self, out, out_counts, inp, in_counts):
        self._dist.all_to_all_single(out, inp, [int(c) for c in out_counts], [int(c) for c in in_counts],
                                     group=self.group)

    def all_reduce_max(self, t):
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)

    def all_reduce_sum(self, t):
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)

    same_process = False

    def all_gather_object(self, obj):
        out = [None] * self.world
        self._dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self._dist.barrier(group=self.group)


class LocalGroup(object):
    """Shared state of `world` virtual ranks running as threads of one process on one device."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def comm(self, rank):
        return LocalComm(self, rank)


class LocalComm(object):
    def __init__(self, group, rank):
        self.g, self.rank, self.world = group, rank, group.world

    def _exchange(self, item):
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()    # virtual ranks run on their own streams
        self.g.slots[self.rank] = item
        self.g.barrier.wait()
        items = list(self.g.slots)
        self.g.barrier.wait()
        return items

    def all_to_all(self, out, out_counts, inp, in_counts):
        items = self._exchange((inp, [int(c) for c in in_counts]))
        at = 0
        for src in range(self.world):
            sbuf, scounts = items[src]
            off = sum(scounts[:self.rank])
            n = scounts[self.rank]
            assert n == int(out_counts[src]), 'all_to_all counts disagree between ranks'
            out[at:at + n].copy_(sbuf[off:off + n])
            at += n
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        self.g.barrier.wait()    # nobody reuses its input buffer before every rank has copied from it

    def all_reduce_max(self, t):
        items = self._exchange(t.clone())
        t.copy_(torch.stack(items).max(0).values)

    def all_reduce_sum(self, t):
        items = self._exchange(t.clone())
        acc = items[0].clone()
        for x in items[1:]:
            acc += x
        t.copy_(acc)

    same_process = True

    def all_gather_object(self, obj):
        return self._exchange(obj)

    def barrier(self):
        self.g.barrier.wait()


# ---------------------------------------------------------------------------------------------
# CUDA kernel backend (the product path)
# ---------------------------------------------------------------------------------------------
class CudaShardBackend(object):
    """Owns this rank's local tables (torch CUDA tensors), their torch optimiser and the native handles."""

    def __init__(self, rank, world, num_users, num_items, dim, local_tables=None, optimizer='adam', lr=1e-3, l2=0.0,
                 betas=(0.5, 0.999), fast_math=None, device=None, seed=0):
        from spotlight.factorization.representations import BilinearNet   # the drop-in (package __init__ put it on sys.path)
        from .engine import MFEngine
        if min(local_rows(num_users, rank, world), local_rows(num_items, rank, world)) <= 0:
            raise ValueError('rank %d of %d would own no rows of a %d x %d model: use at most min(users, items) ranks'
                             % (rank, world, num_users, num_items))
        N.require_cuda()
        self.rank, self.world = int(rank), int(world)
        self.num_users, self.num_items, self.dim = int(num_users), int(num_items), int(dim)
        self.device = device or torch.device('cuda', torch.cuda.current_device())
        lu, li = local_rows(num_users, rank, world), local_rows(num_items, rank, world)
        with torch.device('meta'):
            net = BilinearNet(lu, li, dim)
        net = net.to_empty(device=self.device)
        with torch.no_grad():
            if local_tables is not None:
                for p, t in zip((net.user_embeddings.weight, net.item_embeddings.weight, net.user_biases.weight,
                                 net.item_biases.weight), local_tables):
                    p.copy_(torch.as_tensor(np.ascontiguousarray(t, dtype=np.float32)).reshape(p.shape).to(self.device))
            else:
                # ScaledEmbedding / ZeroEmbedding initialisation (layers.py:30-35,49-56) drawn on the device per rank
                gen = torch.Generator(device=self.device)
                gen.manual_seed(int(seed) * 1000003 + self.rank)
                net.user_embeddings.weight.normal_(0, 1.0 / dim, generator=gen)
                net.item_embeddings.weight.normal_(0, 1.0 / dim, generator=gen)
                net.user_biases.weight.zero_()
                net.item_biases.weight.zero_()
        self.net = net
        if optimizer == 'adam':
            self.optimizer = torch.optim.Adam(net.parameters(), lr=lr, betas=betas, weight_decay=l2)
        elif optimizer == 'sgd':
            self.optimizer = torch.optim.SGD(net.parameters(), lr=lr, weight_decay=l2)
        elif optimizer == 'rms':
            self.optimizer = torch.optim.RMSprop(net.parameters(), lr=lr, weight_decay=l2)
        else:
            raise NotImplementedError('optimizer %r' % (optimizer,))
        self.engine = MFEngine(net, self.optimizer, fast_math=fast_math)
        self._lib = N.load_library()
        self._shard = ctypes.c_void_p(0)
        with torch.cuda.device(self.device):
            N.check(self._lib.mfb_shard_create(self.engine._handle, self.rank, self.world, self.num_users,
                                               self.num_items, ctypes.byref(self._shard)), 'shard_create')
        self.stride = int(self._lib.mfb_shard_row_stride(self._shard))

    # -- buffers ---------------------------------------------------------------------------------
    def zeros(self, n, dtype):
        return torch.zeros(int(n), dtype=dtype, device=self.device)

    def empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def ids(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=torch.int64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.int64)).to(self.device)

    def draw_negatives(self, state625, pop_users, pop_items, k):
        """random.choices(neg_examples, k) (implicit.py:352) continuing the MT19937 stream in state625 (in place)."""
        out_u, out_i = self.empty(k, torch.int64), self.empty(k, torch.int64)
        with torch.cuda.device(self.device):
            N.check(self._lib.mfb_mt_choices_pairs(N.hptr(state625), N.dptr(pop_users), N.dptr(pop_items),
                                                   pop_users.numel(), int(k), N.dptr(out_u), N.dptr(out_i),
                                                   N.stream_ptr()), 'draw_negatives')
        return out_u, out_i

    # -- phases ----------------------------------------------------------------------------------
    def _call(self, name, *args):
        with torch.cuda.device(self.device):
            N.check(getattr(self._lib, name)(self._shard, *args), name)

    def plan_begin(self):
        """Planning runs on its own stream so that chunk c+1 is planned while chunk c executes; everything queued on
        the current stream so far (the id tensors) is ordered before the first plan."""
        if getattr(self, '_plan_stream', None) is None:
            self._plan_stream = torch.cuda.Stream(device=self.device)
        self._plan_stream.wait_stream(torch.cuda.current_stream(self.device))

    def plan(self, pos_users, pos_items, batch, n_neg, neg_users, neg_items, step0, nsteps):
        counts = np.zeros((nsteps, self.world, self.world), dtype=np.int64)
        if getattr(self, '_plan_stream', None) is None:
            self.plan_begin()
        with torch.cuda.stream(self._plan_stream):
            self._call('mfb_shard_plan', N.dptr(pos_users), N.dptr(pos_items), pos_users.numel(), int(batch),
                       int(n_neg), N.dptr(neg_users), N.dptr(neg_items), int(step0), int(nsteps), N.hptr(counts),
                       N.stream_ptr())
        return counts

    def gather(self, s, send):
        self._call('mfb_shard_gather', int(s), N.dptr(send), N.stream_ptr())

    def forward(self, loss, s, recv, cell):
        self._call('mfb_shard_forward', N.LOSS[loss], int(s), N.dptr(recv), N.dptr(cell), N.stream_ptr())

    def backward(self, loss, s, recv, cell, gsend, partial):
        self._call('mfb_shard_backward', N.LOSS[loss], int(s), N.dptr(recv), N.dptr(cell), N.dptr(gsend),
                   ctypes.c_void_p(partial.data_ptr()), N.stream_ptr())

    def update(self, s, grecv):
        self._call('mfb_shard_update', int(s), N.dptr(grecv), N.stream_ptr())

    # -- direct exchange over peer memory (no collective on the step's critical path) ----------------
    def enable_direct(self, batch, n_neg, comm):
        """Allocates this rank's exchange buffer for the minibatch geometry and maps every peer's (CUDA IPC across
        processes, plain pointers between virtual ranks of one process).  Collective: all ranks call it together."""
        self._close_peers()
        ptr, nbytes = ctypes.c_void_p(0), ctypes.c_int64(0)
        self._call('mfb_shard_xbuf_alloc', int(batch), int(n_neg), ctypes.byref(ptr), ctypes.byref(nbytes))
        if comm.same_process:
            ptrs = comm.all_gather_object(int(ptr.value))
        else:
            handle = (ctypes.c_ubyte * 64)()
            with torch.cuda.device(self.device):
                N.check(self._lib.mfb_ipc_export(ptr, ctypes.cast(handle, ctypes.c_void_p)), 'ipc_export')
            handles = comm.all_gather_object(bytes(handle))
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(int(ptr.value))
                    continue
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                peer = ctypes.c_void_p(0)
                with torch.cuda.device(self.device):
                    N.check(self._lib.mfb_ipc_open(ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(peer)), 'ipc_open')
                self._peer_maps.append(peer)
                ptrs.append(int(peer.value))
        arr = (ctypes.c_void_p * self.world)(*ptrs)
        self._call('mfb_shard_xbuf_set_peers', ctypes.cast(arr, ctypes.c_void_p))
        self.direct_geometry = (int(batch), int(n_neg))
        comm.barrier()        # every rank's flags are zeroed and mapped before anyone signals

    def run_steps(self, loss, s_begin, s_end, partial):
        self._call('mfb_shard_run_steps', N.LOSS[loss], int(s_begin), int(s_end), ctypes.c_void_p(partial.data_ptr()),
                   N.stream_ptr())

    def run_phase(self, loss, s, phase, partial):
        self._call('mfb_shard_run_phase', N.LOSS[loss], int(s), int(phase), ctypes.c_void_p(partial.data_ptr()),
                   N.stream_ptr())

    def direct_check(self):
        self._call('mfb_shard_direct_check', N.stream_ptr())

    def _close_peers(self):
        for peer in getattr(self, '_peer_maps', []):
            self._lib.mfb_ipc_close(peer)
        self._peer_maps = []
        self.direct_geometry = None

    def flush(self):
        self.engine.flush()

    def local_tables(self):
        self.flush()
        return [p.detach().cpu().numpy() for p in (self.net.user_embeddings.weight, self.net.item_embeddings.weight,
                                                   self.net.user_biases.weight, self.net.item_biases.weight)]

    @property
    def launches(self):
        return int(self._lib.mfb_shard_launches(self._shard)) + int(self._lib.mfb_model_launches(self.engine._handle))

    def close(self):
        if getattr(self, '_shard', None):
            torch.cuda.synchronize(self.device)
            self._close_peers()
            self._lib.mfb_shard_destroy(self._shard)
            self._shard = ctypes.c_void_p(0)
        self.engine.close()


# ---------------------------------------------------------------------------------------------
# orchestration (transport- and backend-agnostic)
# ---------------------------------------------------------------------------------------------
class ShardedMF(object):
    """One rank of the row-sharded model.  `loss` names follow ImplicitFactorizationModel: 'bpr' and
    'adaptive_hinge' both train adaptive hinge against the batch's single largest negative (implicit.py:202-212,
    losses.py:133-172 on 1-D negatives), 'pointwise' and 'hinge' are themselves; 'bpr_pairwise' is losses.bpr_loss."""

    LOSS_KERNEL = {'pointwise': 'pointwise', 'bpr': 'adaptive_hinge', 'adaptive_hinge': 'adaptive_hinge',
                   'hinge': 'hinge', 'bpr_pairwise': 'bpr'}

    def __init__(self, backend, comm, chunk_steps=64, direct=False):
        if backend.rank != comm.rank or backend.world != comm.world:
            raise ValueError('backend is rank %d of %d, transport is rank %d of %d'
                             % (backend.rank, backend.world, comm.rank, comm.world))
        self.backend, self.comm = backend, comm
        self.rank, self.world = comm.rank, comm.world
        self.chunk_steps = int(chunk_steps)
        # direct=True: rows and gradients move by peer-memory stores inside the kernels (NVLink), no collective per step
        self.direct = bool(direct)
        self.stride = backend.stride
        self._bufs = {}

    def _buf(self, name, nfloats):
        cur = self._bufs.get(name)
        if cur is None or cur.numel() < nfloats:
            cur = self.backend.empty(int(nfloats * 1.25) + 64, torch.float32)
            self._bufs[name] = cur
        return cur

    def train_steps(self, loss, pos_users, pos_items, batch, n_neg, neg_users, neg_items, step0=0, nsteps=None):
        """Steps [step0, step0+nsteps) of the epoch over the positives; neg_* hold n_neg*batch pairs per step
        starting at step0's.  Returns the per-step losses (numpy float64), identical on every rank."""
        kind = self.LOSS_KERNEL[loss]
        be, comm, stride = self.backend, self.comm, self.stride
        pos_users, pos_items = be.ids(pos_users), be.ids(pos_items)
        neg_users, neg_items = be.ids(neg_users), be.ids(neg_items)
        n_pos = pos_users.numel()
        total = (n_pos + batch - 1) // batch
        nsteps = total - step0 if nsteps is None else int(nsteps)
        m = n_neg * batch
        if neg_users.numel() < nsteps * m:
            raise ValueError('need %d negative pairs, got %d' % (nsteps * m, neg_users.numel()))
        losses = np.zeros(nsteps, dtype=np.float64)
        cell = be.zeros(1, torch.int64)
        if self.direct and getattr(be, 'direct_geometry', None) != (int(batch), int(n_neg)):
            be.enable_direct(batch, n_neg, comm)
        chunks = [(c0, min(self.chunk_steps, nsteps - c0)) for c0 in range(0, nsteps, self.chunk_steps)]

        def plan(idx):
            c0, ns = chunks[idx]
            return be.plan(pos_users, pos_items, batch, n_neg, neg_users[c0 * m:(c0 + ns) * m],
                           neg_items[c0 * m:(c0 + ns) * m], step0 + c0, ns)
        if hasattr(be, 'plan_begin'):
            be.plan_begin()
        counts_next = plan(0)
        partials = []
        for idx, (c0, ns) in enumerate(chunks):
            counts = counts_next
            send_counts = counts[:, self.rank, :] * stride     # [ns, world] floats to each computing rank
            recv_counts = counts[:, :, self.rank] * stride     # [ns, world] floats from each owner
            n_send, n_recv = send_counts.sum(1), recv_counts.sum(1)
            partial = be.zeros(2 * ns, torch.float64)
            partials.append(partial)
            if self.direct and comm.same_process and comm.world > 1:
                # virtual ranks share one device and one stream: phase p of every rank is enqueued before phase p+1 of
                # any, so a wait kernel is never launched before the signals it waits for (nothing ever spins)
                for s in range(ns):
                    for phase in range(4):
                        comm.barrier()
                        be.run_phase(kind, s, phase, partial[2 * s:2 * s + 2])
            elif self.direct:
                be.run_steps(kind, 0, ns, partial)
            else:
                send, grecv = self._buf('send', n_send.max()), self._buf('grecv', n_send.max())
                recv, gsend = self._buf('recv', n_recv.max()), self._buf('gsend', n_recv.max())
                for s in range(ns):
                    be.gather(s, send)
                    comm.all_to_all(recv[:n_recv[s]], recv_counts[s], send[:n_send[s]], send_counts[s])
                    be.forward(kind, s, recv, cell)
                    if kind == 'adaptive_hinge':
                        comm.all_reduce_max(cell)
                    be.backward(kind, s, recv, cell, gsend, partial[2 * s:2 * s + 2])
                    comm.all_to_all(grecv[:n_send[s]], send_counts[s], gsend[:n_recv[s]], recv_counts[s])
                    be.update(s, grecv)
            if idx + 1 < len(chunks):
                counts_next = plan(idx + 1)    # planned on the planning stream while this chunk executes
        if self.direct:
            be.direct_check()
        for (c0, ns), partial in zip(chunks, partials):
            comm.all_reduce_sum(partial)
            p = partial.cpu().numpy().reshape(ns, 2)
            for s in range(ns):
                b = min(batch, n_pos - (step0 + c0 + s) * batch)
                losses[c0 + s] = p[s, 0] / b + (p[s, 1] / m if (kind == 'pointwise' and m > 0) else 0.0)
        return losses

    def train_epoch(self, loss, pos_users, pos_items, batch, n_neg, pop_users, pop_items, state625):
        """One epoch of implicit.py:289-298 with the negatives drawn chunk by chunk from the MT19937 stream in
        state625 (`random.getstate()[1]` layout, advanced in place; every rank passes the same state)."""
        be = self.backend
        pos_users, pos_items = be.ids(pos_users), be.ids(pos_items)
        pop_users, pop_items = be.ids(pop_users), be.ids(pop_items)
        n_pos = pos_users.numel()
        total = (n_pos + batch - 1) // batch
        m = n_neg * batch
        out = []
        block = 16 * self.chunk_steps           # negatives are drawn for 16 chunks at a time
        for c0 in range(0, total, block):
            ns = min(block, total - c0)
            neg_u, neg_i = be.draw_negatives(state625, pop_users, pop_items, ns * m)
            out.append(self.train_steps(loss, pos_users, pos_items, batch, n_neg, neg_u, neg_i, step0=c0, nsteps=ns))
        return np.concatenate(out)

    def flush(self):
        self.backend.flush()

    def local_tables(self):
        return self.backend.local_tables()

    def close(self):
        self.backend.close()


def run_local_ranks(world, make_rank, work):
    """Runs `work(shard)` on `world` virtual ranks as threads (LocalComm); make_rank(rank, comm) -> ShardedMF.
    Returns the list of results in rank order.  Test/bring-up helper for boxes with fewer GPUs than ranks."""
    group = LocalGroup(world)
    results, errors = [None] * world, []

    def body(rank):
        try:
            shard = make_rank(rank, group.comm(rank))      # all virtual ranks enqueue on the same (current) stream
            results[rank] = work(shard)
        except BaseException as exc:   # noqa: BLE001 -- re-raised in the caller; abort the barrier so peers stop
            errors.append(exc)
            group.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
        raise (real or errors)[0]
    return results
