# measured accumulation error of the evaluation GEMM when the accumulator is initialised (extra K = 16 MMA step) with item
# biases much larger than the dot products: max |tensor-core score - float64 value| in units of 2^-24 * |b|max
import sys, numpy as np, torch
sys.path.insert(0, '.')
import recommendation_gans_b200
from tests.gpu_helpers import make_net
from recommendation_gans_b200.engine import MFEngine
U, I, D = 512, 4096, 128
for bias_scale in (1.0, 50.0, 3000.0):
    for emb_scale in (0.05, 1.0):
        rs = np.random.RandomState(3)
        tabs = [rs.normal(0, emb_scale, (U, D)).astype(np.float32), rs.normal(0, emb_scale, (I, D)).astype(np.float32),
                rs.normal(0, 0.1, (U, 1)).astype(np.float32), rs.normal(0, bias_scale, (I, 1)).astype(np.float32)]
        eng = MFEngine(make_net(tabs))
        users = np.arange(U, dtype=np.int64)
        got = eng.debug_tc_scores(users).cpu().numpy().astype(np.float64)
        ub = torch.from_numpy(tabs[0]).cuda().half().double(); vb = torch.from_numpy(tabs[1]).cuda().half().double()
        ref = (vb @ ub.T).cpu().numpy() + tabs[3].astype(np.float64)
        err = np.abs(got - ref)
        mag = np.abs(tabs[3]).max() + np.abs(ref - tabs[3]).max()
        print('bias scale %7.1f  embedding scale %.2f: max error %.3e = %.2f * 2^-24 * (|b|max + |dot|max = %.3g); budget 2^-16 = 256 * 2^-24'
              % (bias_scale, emb_scale, err.max(), err.max() / (2.0 ** -24 * mag), mag))
