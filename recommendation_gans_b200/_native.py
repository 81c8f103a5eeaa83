"""ctypes binding of the mfb200 C ABI (include/mfb200.h).

There is deliberately no fallback: if the CUDA library is missing or no CUDA device is
present, every entry point raises.  torch is used only for device memory and streams.
"""
import ctypes
import os

import numpy as np
import torch

from . import build as _build

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_void = ctypes.c_void_p

MFB_OK = 0
ERR_INVALID, ERR_CUDA, ERR_RANGE, ERR_SHAPE, ERR_UNSUPPORTED, ERR_NOMEM = -1, -2, -3, -4, -5, -6
LOSS = {'pointwise': 0, 'bpr': 1, 'hinge': 2, 'adaptive_hinge': 3}
OPT_SGD, OPT_ADAM, OPT_RMSPROP = 0, 1, 2
MAX_TOPK = 32


class ModelDesc(ctypes.Structure):
    _fields_ = [('num_users', ctypes.c_int32), ('num_items', ctypes.c_int32), ('dim', ctypes.c_int32),
                ('optimizer', ctypes.c_int32),
                ('lr', ctypes.c_double), ('beta1', ctypes.c_double), ('beta2', ctypes.c_double),
                ('eps', ctypes.c_double), ('weight_decay', ctypes.c_double),
                ('d_user_emb', c_void), ('d_item_emb', c_void), ('d_user_bias', c_void), ('d_item_bias', c_void),
                ('d_user_emb_m', c_void), ('d_user_emb_v', c_void), ('d_item_emb_m', c_void), ('d_item_emb_v', c_void),
                ('d_user_bias_m', c_void), ('d_user_bias_v', c_void), ('d_item_bias_m', c_void),
                ('d_item_bias_v', c_void),
                ('fast_math', ctypes.c_int32), ('reserved', ctypes.c_int32)]


# name -> (restype, argtypes); must list every symbol include/mfb200.h declares
SIGNATURES = {
    'mfb_version': (ctypes.c_int, []),
    'mfb_last_error': (ctypes.c_char_p, []),
    'mfb_model_create': (ctypes.c_int, [ctypes.POINTER(ModelDesc), ctypes.POINTER(c_void)]),
    'mfb_model_destroy': (ctypes.c_int, [c_void]),
    'mfb_model_step': (ctypes.c_int64, [c_void]),
    'mfb_model_set_step': (ctypes.c_int, [c_void, ctypes.c_int64]),
    'mfb_mt_choices_pairs': (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int64, ctypes.c_int64, c_void, c_void, c_void]),
    'mfb_mt_choices_indices': (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int64, c_void, c_void]),
    'mfb_mt_sample_items': (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int64, c_void, c_void]),
    'mfb_mt_words': (ctypes.c_int, [c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_mt_words_parallel': (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int64, c_void, c_void]),
    'mfb_mt_jump_poly': (ctypes.c_int, [ctypes.c_int64, c_void]),
    'mfb_negative_pairs': (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_void, c_void, c_void,
                                          c_void, c_void, c_void, c_void, c_void]),
    'mfb_predict_pairs': (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_predict_user': (ctypes.c_int, [c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_loss_forward_backward': (ctypes.c_int, [ctypes.c_int, c_void, ctypes.c_int64, c_void, ctypes.c_int64,
                                                 c_void, c_void, c_void, c_void]),
    'mfb_loss_forward_backward_ex': (ctypes.c_int, [ctypes.c_int, c_void, ctypes.c_int64, c_void, ctypes.c_int64,
                                                    ctypes.c_int64, c_void, c_void, c_void, c_void, c_void]),
    'mfb_train_steps': (ctypes.c_int, [c_void, ctypes.c_int, c_void, c_void, ctypes.c_int64, ctypes.c_int32,
                                       ctypes.c_int32, c_void, c_void, c_void, c_void]),
    'mfb_loss_steps': (ctypes.c_int, [c_void, ctypes.c_int, c_void, c_void, ctypes.c_int64, ctypes.c_int32,
                                      ctypes.c_int32, c_void, c_void, c_void, c_void]),
    'mfb_flush': (ctypes.c_int, [c_void, c_void]),
    'mfb_train_epoch_host': (ctypes.c_int, [c_void, ctypes.c_int, c_void, c_void, ctypes.c_int64, ctypes.c_int32,
                                            ctypes.c_int32, c_void, c_void, c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_topk': (ctypes.c_int, [c_void, c_void, ctypes.c_int64, c_void, c_void, ctypes.c_int32, c_void, c_void,
                                c_void]),
    'mfb_topk_keyed': (ctypes.c_int, [c_void, c_void, ctypes.c_int64, c_void, c_void, ctypes.c_int32, c_void, c_void,
                                      ctypes.c_uint64, c_void]),
    'mfb_topk_scores': (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int64, c_void, c_void, c_void, ctypes.c_int32, c_void,
                                       c_void, c_void, c_void]),
    'mfb_topk_hits': (ctypes.c_int, [c_void, c_void, ctypes.c_int64, ctypes.c_int32, c_void, c_void, c_void,
                                     ctypes.c_int32, c_void, c_void, c_void]),
    'mfb_rank_test_items': (ctypes.c_int, [c_void, c_void, ctypes.c_int64, c_void, c_void, c_void, c_void, c_void, c_void]),
    'mfb_model_rng_seed': (ctypes.c_int, [c_void, c_void, c_void]),
    'mfb_model_rng_state': (ctypes.c_int, [c_void, c_void, c_void]),
    'mfb_train_epoch': (ctypes.c_int, [c_void, ctypes.c_int, c_void, c_void, ctypes.c_int64, ctypes.c_int32,
                                       ctypes.c_int32, c_void, c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_loss_epoch': (ctypes.c_int, [c_void, ctypes.c_int, c_void, c_void, ctypes.c_int64, ctypes.c_int32,
                                      ctypes.c_int32, c_void, c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_topk_last_redo': (ctypes.c_int, [c_void]),
    'mfb_debug_tc_stats': (ctypes.c_int, [c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_debug_tc_scores': (ctypes.c_int, [c_void, c_void, ctypes.c_int64, c_void, c_void]),
    'mfb_shard_create': (ctypes.c_int, [c_void, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.POINTER(c_void)]),
    'mfb_shard_destroy': (ctypes.c_int, [c_void]),
    'mfb_shard_row_stride': (ctypes.c_int32, [c_void]),
    'mfb_shard_launches': (ctypes.c_int64, [c_void]),
    'mfb_shard_plan': (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, c_void,
                                      c_void, ctypes.c_int64, ctypes.c_int32, c_void, c_void]),
    'mfb_shard_gather': (ctypes.c_int, [c_void, ctypes.c_int32, c_void, c_void]),
    'mfb_shard_forward': (ctypes.c_int, [c_void, ctypes.c_int, ctypes.c_int32, c_void, c_void, c_void]),
    'mfb_shard_backward': (ctypes.c_int, [c_void, ctypes.c_int, ctypes.c_int32, c_void, c_void, c_void, c_void,
                                          c_void]),
    'mfb_shard_update': (ctypes.c_int, [c_void, ctypes.c_int32, c_void, c_void]),
    'mfb_shard_xbuf_alloc': (ctypes.c_int, [c_void, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(c_void),
                                            ctypes.POINTER(ctypes.c_int64)]),
    'mfb_shard_xbuf_set_peers': (ctypes.c_int, [c_void, c_void]),
    'mfb_ipc_export': (ctypes.c_int, [c_void, c_void]),
    'mfb_ipc_open': (ctypes.c_int, [c_void, ctypes.POINTER(c_void)]),
    'mfb_ipc_close': (ctypes.c_int, [c_void]),
    'mfb_shard_run_steps': (ctypes.c_int, [c_void, ctypes.c_int, ctypes.c_int32, ctypes.c_int32, c_void, c_void]),
    'mfb_shard_run_phase': (ctypes.c_int, [c_void, ctypes.c_int, ctypes.c_int32, ctypes.c_int32, c_void, c_void]),
    'mfb_shard_direct_check': (ctypes.c_int, [c_void, c_void]),
    'mfb_profile_enable': (ctypes.c_int, [c_void, ctypes.c_int]),
    'mfb_profile_read': (ctypes.c_int, [c_void, c_void, c_void]),
    'mfb_profile_name': (ctypes.c_char_p, [ctypes.c_int]),
    'mfb_model_launches': (ctypes.c_int64, [c_void]),
    'mfb_library_launches': (ctypes.c_int64, []),
}
PROFILE_CLASSES = 12

_lib = None


def load_library():
    """Loads (building if the sources are newer) lib/libmfb200.so.  Does not need a GPU."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    override = os.environ.get('MFB_LIB_PATH')
    if override:
        path = override
    elif _build.is_stale():
        try:
            path = _build.build_library()
        except Exception as exc:  # no nvcc on this box and no prebuilt library
            if not os.path.exists(_build.LIB_PATH):
                raise RuntimeError('mfb200: CUDA library missing and cannot be built (%s). '
                                   'There is no CPU fallback.' % exc)
            path = _build.LIB_PATH
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('mfb200 runs on CUDA (sm_100a) only and no CUDA device is visible; '
                           'there is no CPU fallback.')


class MfbError(RuntimeError):
    pass


def check(rc, what=''):
    """Maps C status codes onto the exception types the reference raises."""
    if rc == MFB_OK:
        return
    msg = load_library().mfb_last_error().decode('utf-8', 'replace')
    text = ('%s: %s' % (what, msg)) if what else msg
    if rc == ERR_RANGE:
        raise ValueError(text)                 # implicit.py:222-236
    if rc == ERR_SHAPE:
        raise RuntimeError(text)               # torch broadcasting error in losses.py
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    if rc == ERR_NOMEM:
        raise MemoryError(text)
    raise MfbError('mfb200 status %d: %s' % (rc, text))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return ctypes.c_void_p(t.data_ptr())


def hptr(a):
    """Host pointer of a contiguous numpy array."""
    assert isinstance(a, np.ndarray) and a.flags['C_CONTIGUOUS']
    return ctypes.c_void_p(a.ctypes.data)
