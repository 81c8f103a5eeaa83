"""CPU-only: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/mfb200.h declares (no compute calls here)."""
import os
import re

import recommendation_gans_b200  # noqa: F401
from recommendation_gans_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mfb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mfb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_header_symbols():
    lib = N.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(N.SIGNATURES), set(names) ^ set(N.SIGNATURES)
    assert lib.mfb_version() == 100


def test_dropin_modules_import_without_gpu():
    import implicit
    import spotlight.evaluation
    import spotlight.losses
    import spotlight.sampling
    from spotlight.factorization.representations import BilinearNet
    assert implicit.__file__.startswith(recommendation_gans_b200.DROPIN_PATH)
    net = BilinearNet(7, 5, 4)
    assert sorted(net.state_dict()) == ['item_biases.weight', 'item_embeddings.weight',
                                        'user_biases.weight', 'user_embeddings.weight']


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from implicit import ImplicitFactorizationModel
    with pytest.raises(RuntimeError):
        ImplicitFactorizationModel(experiment_name='cpu_refusal')
    from spotlight.factorization.representations import BilinearNet
    net = BilinearNet(7, 5, 4)
    with pytest.raises(RuntimeError):
        net(torch.tensor([0, 1]), torch.tensor([1, 2]))


def test_library_staleness_is_decided_by_content(tmp_path, monkeypatch):
    """The prebuilt library travels to the GPU box in a snapshot that does not keep a meaningful mtime order; a rebuild
    there would have every rank of a multi-GPU run rewriting the .so under the others.  Staleness therefore compares a
    content hash of the sources with the one recorded at build time."""
    import os
    from recommendation_gans_b200 import build as B
    B.build_library()
    assert not B.is_stale()
    for path in B.sources():                     # touching every source must not make the library stale
        os.utime(path, None)
    assert not B.is_stale()
    extra = tmp_path / 'extra.cuh'
    extra.write_text('// a change in any dependency')
    real = B._dependencies
    monkeypatch.setattr(B, '_dependencies', lambda: real() + [str(extra)])
    assert B.is_stale()


def test_committed_bench_lines_keep_the_contract():
    """The bench lines committed under profiles/ (written by bench.py on a B200) carry every key the driver reads."""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    paths = [p for p in glob.glob(os.path.join(root, 'profiles', 'r01_bench_*gpu.json'))] + \
        [os.path.join(root, 'profiles', 'r01_bench_session2.json')]
    assert len(paths) >= 3
    for path in paths:
        d = json.load(open(path))
        for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                    'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline'):
            assert key in d, (path, key)
        assert d['config']['workload'] and 'model' not in d['config']
        assert set(d['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
        assert set(d['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
        assert d['gpu_launches'] > 0 and d['e2e']['h2d_bytes_per_step'] > 0
        assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
        if d['n_gpus'] == 1:
            assert set(d['cpu_baseline']) >= {'value', 'unit', 'cores', 'kind', 'sample'}


def test_dropin_is_standalone_for_mf_spotlight_and_falls_through_for_the_rest():
    """Every module the reference's mf_spotlight.py imports (mf_spotlight.py:1-13) resolves INSIDE the drop-in directory
    -- no reference checkout needed -- while the parts of the `spotlight` / `utils` packages the drop-in does not
    provide (sequence models, slate data, ...) still come from a checkout that follows it on sys.path.  The second
    half runs only where the reference is mounted (this container); nothing is read from it on the GPU box."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dropin = os.path.join(root, 'recommendation_gans_b200', 'dropin')
    code = (
        "import spotlight.optimizers, spotlight.dataset_manilupation, implicit, spotlight.sampling, utils.arg_extractor\n"
        "import spotlight.factorization.representations, spotlight.dnn_models.mlp, spotlight.dnn_models.neuMF\n"
        "import utils.data_provider, utils.helper_functions, spotlight.datasets.movielens, spotlight.evaluation\n"
        "import sys\n"
        "mods = [m for n, m in sys.modules.items() if n.split('.')[0] in ('spotlight', 'utils', 'implicit')]\n"
        "print('\\n'.join(sorted(getattr(m, '__file__', None) or '' for m in mods)))\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([dropin, root]))
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    files = [f for f in out.stdout.strip().splitlines() if f]
    assert files and all(f.startswith(dropin) for f in files), files
    ref = os.environ.get('REF_PATH', '/root/reference')
    if not os.path.isdir(os.path.join(ref, 'spotlight')):
        return
    code = ("import spotlight.interactions, spotlight.sequence.representations as sr, utils.slate_data_provider as sd\n"
            "print(spotlight.interactions.__file__); print(sr.__file__); print(sd.__file__)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([dropin, root, ref]))
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith(dropin) and lines[1].startswith(ref) and lines[2].startswith(ref)
