"""Steps/s of MultimodalAutoencoder.train() at the reference's own batch sizes (B = 20, grid-search fit)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodalautoencoder_b200 import MultimodalAutoencoder
from multimodalautoencoder_b200.data_funcs import DataLoader
from multimodalautoencoder_b200.synthetic import make_frame

df = make_frame(5000, seed=1)
dl = DataLoader(df=df, supervised=False, cross_validation=True, normalize_and_fill=False, suppress_output=True)
for mode in ('numpy', 'philox'):
    for layers, B in (([128, 64], 20), ([1000, 100], 20), ([128, 64], 4096)):
        m = MultimodalAutoencoder(data_loader=dl, layer_sizes=layers, variational=False, tie_weights=False, batch_size=B,
                                  learning_rate=1e-3, weight_initialization='normal', loss_func='sigmoid_cross_entropy',
                                  verbose=False, precision='tf32', rng_mode=mode)
        np.random.seed(0)
        m.train(200, record_every_nth=100000, save_every_nth=10 ** 9)
        torch.cuda.synchronize()
        n = 1500 if B == 20 else 300
        t0 = time.perf_counter()
        m.train(n, record_every_nth=100000, save_every_nth=10 ** 9)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print('%-7s layers %-12s B=%5d: %7.1f us/step  %10.0f samples/s  graph replays %d / launches %d'
              % (mode, layers, B, 1e6 * dt / n, B * n / dt, m.engine.graph_replays, m.engine.kernel_launches), flush=True)
