"""Data-parallel parity: N ranks on row shards of one global batch == 1 rank on the whole batch.
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from multimodalautoencoder_b200 import Engine, EngineConfig, dp

rank, world, local = dp.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
STARTS, NAMES = [0, 200, 220, 240, 270, 320], ['phys', 'call', 'sms', 'screen', 'location']
ok = True
for name, kw in (('rmse-dropout', dict(loss_func='mean_squared', tie_weights=True)),
                 ('vae', dict(variational=True, tie_weights=False, layer_sizes=[128, 64, 32])),
                 ('sce-untied-tf32', dict(tie_weights=False, precision='tf32')),
                 # wide first layer: its weight gradient is all-reduced in four row blocks while it is still being computed
                 ('wide-first-layer-tf32', dict(tie_weights=False, precision='tf32', num_feats=1024, layer_sizes=[4096, 64],
                                                modality_starts=[0, 256, 512, 640, 768, 1024]))):
    base = dict(num_feats=320, layer_sizes=[128, 64], modality_starts=STARTS, modality_names=NAMES, weight_penalty=0.001,
                learning_rate=1e-3, seed=5, precision='fp32')
    base.update(kw)
    cfg = EngineConfig(**base)
    Bg = 1024
    rng = np.random.default_rng(0)
    eng = Engine(cfg)
    params = {}
    for vname, shp in eng.variables():
        params[vname] = np.full(shp, 0.1, np.float32) if len(shp) == 1 else (rng.standard_normal(shp) / np.sqrt(shp[0])).astype(np.float32)
    eng.set_params(params)
    first, rows = dp.attach_data_parallel(eng, Bg, dist)
    X = rng.uniform(0, 1, (Bg, cfg.num_feats)).astype(np.float32)
    keep = 0.5 if 'dropout' in name else 1.0
    losses = []
    for s in range(3):
        eng.set_rng_step(10 + s)
        eng.gen_noise(rows, first)
        eng.train_step(X[first:first + rows], noise=True, keep=keep)
        losses.append(eng.scalars()['recon_loss'])
    if rank == 0:
        ref = Engine(cfg)
        ref.set_params(params)
        rl = []
        for s in range(3):
            ref.set_rng_step(10 + s)
            ref.gen_noise(Bg, 0)
            ref.train_step(X, noise=True, keep=keep)
            rl.append(ref.scalars()['recon_loss'])
        worst = 0.0
        for vname, _ in eng.variables():
            a, b = eng.get_variable(vname), ref.get_variable(vname)
            moved = np.linalg.norm(b.astype(np.float64) - params[vname]) + 1e-12
            worst = max(worst, np.linalg.norm(a.astype(np.float64) - b) / moved)
        lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, rl))
        tol = 5e-2 if 'tf32' in name else 2e-3
        good = lerr < 1e-5 and worst < tol
        ok = ok and good
        print('dp_check[%s] world=%d: loss rel err %.2e, param movement mismatch %.2e -> %s' % (name, world, lerr, worst, 'OK' if good else 'FAIL'))
    dist.barrier()
    eng.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
