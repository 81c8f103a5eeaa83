"""predict() id handling (reference: spotlight/factorization/_components.py:8-25): a scalar user is
paired with every item; arrays are taken pairwise.  Returns flat int64 CUDA tensors."""
import numpy as np
import torch


def _predict_process_ids(user_ids, item_ids, num_items, use_cuda=True):
    if item_ids is None:
        item_ids = np.arange(num_items, dtype=np.int64)
    items = np.asarray(item_ids).reshape(-1).astype(np.int64)
    users = np.asarray(user_ids).reshape(-1).astype(np.int64)
    if users.shape[0] != items.shape[0]:
        users = np.broadcast_to(users, items.shape).copy()   # the reference expands a size-1 user
    return torch.from_numpy(users).cuda(), torch.from_numpy(items).cuda()
