"""Times the small-config forward (chain vs per-layer) and train step at several batch sizes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalautoencoder_b200 import Engine, EngineConfig

def mk(chain, B, layers=(128, 64), tie=False):
    os.environ['MMAE_CHAIN'] = '1' if chain else '0'
    cfg = EngineConfig(num_feats=320, layer_sizes=list(layers), modality_starts=[0, 200, 220, 240, 270, 320],
                       modality_names=['phys', 'call', 'sms', 'screen', 'location'], tie_weights=tie, variational=False,
                       activation='softsign', loss_func='sigmoid_cross_entropy', learning_rate=1e-3, seed=0,
                       precision='tf32', max_batch=B)
    e = Engine(cfg)
    rng = np.random.default_rng(0)
    for vname, shp in e.variables():
        e.set_variable(vname, np.full(shp, 0.1, np.float32) if len(shp) == 1 else
                       (np.clip(rng.standard_normal(shp), -2, 2) / np.sqrt(shp[0])).astype(np.float32))
    return e

def timeit(fn, n=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for B in [int(x) for x in (sys.argv[1:] or ['4096', '65536', '2000000'])]:
    X = torch.rand((B, 320), device='cuda')
    for chain in (1, 0):
        e = mk(chain, B)
        ms_f = timeit(lambda: e.forward(X, recon=True, loss=True))
        def tr():
            e.gen_noise(B); e.train_step(X, noise=True)
        ms_t = timeit(tr) if B <= 262144 else float('nan')
        gbs = B * 2560 / ms_f / 1e6
        print('B=%8d chain=%d  forward %.3f ms (%.1f M samples/s, %.0f GB/s algorithmic)   train %.3f ms (%.1f M samples/s)'
              % (B, chain, ms_f, B / ms_f / 1e3, gbs, ms_t, B / ms_t / 1e3), flush=True)
        e.close()
