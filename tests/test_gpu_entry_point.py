"""The reference's own training script, unmodified, against the drop-in package ("mf_spotlight.py must
keep running unmodified"; /root/reference/mf_spotlight.py:55-72).

`__graft_entry__.build()` stages the script byte for byte under tests/fixtures/_ref/ (git-ignored; the SHA-256 below
proves identity).  The test builds the cached CSV / PKL inputs the script's data_provider expects from synthetic
ML-100K-shaped ratings, runs `python mf_spotlight.py ...` with only the drop-in directory on PYTHONPATH and checks what
the reference would leave behind: summary.csv, configuration.json, test_summary.json, best_model."""
import csv
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, 'tests', 'fixtures', '_ref', 'mf_spotlight.py')
SCRIPT_SHA256 = 'd1c6e4b1bad5bc165f4e7fc2a4cdbcd93d5c740c923a56cd72f4518f3c8d4040'


def _write_cache(folder):
    """Synthetic ML-100K-shaped ratings through the drop-in's own preprocessing (rating filter is part of the raw
    loader, so ratings are drawn above 3.5) into the cache layout of utils/data_provider.py:157-178."""
    sys.path.insert(0, os.path.join(ROOT, 'recommendation_gans_b200', 'dropin'))
    import pandas as pd
    from spotlight.interactions import Interactions
    from spotlight.dataset_manilupation import train_test_timebased_split
    from utils.data_provider import data_provider
    from utils.helper_functions import make_implicit
    rs = np.random.RandomState(0)
    U, I, n = 300, 400, 20000
    users = rs.randint(0, U, n).astype(np.int64)
    items = ((users * 13 + rs.randint(0, 12, n)) % I).astype(np.int64)
    data = Interactions(users, items, rs.choice([4.0, 5.0], n), np.arange(n), num_users=U, num_items=I)
    data = make_implicit(data)
    train, test = train_test_timebased_split(data, test_percentage=0.1)
    train, valid = train_test_timebased_split(train, test_percentage=0.1)
    neg = [(int(u), int(i)) for u, i in zip(rs.randint(0, U, len(train)), rs.randint(0, I, len(train)))]
    pop = pd.Series(np.bincount(items, minlength=I), index=np.arange(I))
    w = data_provider.__new__(data_provider)
    w.movies_to_keep = -1
    os.makedirs(folder, exist_ok=True)
    rel = os.path.join(folder, 'movielens_100K')
    w.save_statistics(rel, U, I, n)
    w.create_cvs_files(rel, train, valid, test, neg, pop)
    return U, I, len(train)


@pytest.mark.skipif(not os.path.isfile(SCRIPT), reason='reference entry point not staged (__graft_entry__.build())')
def test_unmodified_mf_spotlight_runs_against_the_dropin(tmp_path):
    with open(SCRIPT, 'rb') as f:
        assert hashlib.sha256(f.read()).hexdigest() == SCRIPT_SHA256      # byte-identical to the reference's file
    U, I, n_train = _write_cache(str(tmp_path / 'datasets' / 'movielens'))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'recommendation_gans_b200', 'dropin'), ROOT]))
    cmd = [sys.executable, SCRIPT, '--dataset', '100K', '--mf_embedding_dim', '32', '--batch_size', '256',
           '--neg_examples', '1', '--k', '5', '--training_epochs', '3', '--use_gpu', 'True',
           '--experiment_name', 'entry_point']
    out = subprocess.run(cmd, env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    logs = tmp_path / 'experiments_results' / 'entry_point' / 'result_outputs'
    with open(logs / 'summary.csv') as f:
        rows = list(csv.reader(f))
    assert rows[0] == ['train_loss', 'validation_loss', 'curr_epoch'] and len(rows) == 4
    losses = np.array([[float(x) for x in r] for r in rows[1:]])
    assert np.isfinite(losses).all() and losses[-1, 0] < losses[0, 0]
    with open(logs / 'configuration.json') as f:
        cfg = json.load(f)
    assert cfg == {'num_users': U, 'num_items': I, 'weight_decay': 1e-05, 'lr': 0.001, 'embedding_dim': 32,
                   'batch_size': 256, 'epochs': 3}
    with open(logs / 'test_summary.json') as f:
        res = json.load(f)
    assert set(res) == {'k', 'bce', 'precision', 'recall', 'rand_prec', 'rand_rec', 'pop_prec', 'pop_rec', 'at_k',
                        'map'}
    assert res['k'] == 5 and 0.0 <= res['precision'] <= 1.0 and 0.0 <= res['map'] <= 1.0
    import torch
    ck = torch.load(tmp_path / 'experiments_results' / 'entry_point' / 'saved_models' / 'best_model')
    sd = ck['network']
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        'user_embeddings.weight': (U, 32), 'item_embeddings.weight': (I, 32), 'user_biases.weight': (U, 1),
        'item_biases.weight': (I, 1)}
    assert 'Model chosen from epoch' in out.stderr and 'precision@5' in out.stderr
