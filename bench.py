#!/usr/bin/env python
"""bench.py -- MMAE train samples/sec (fwd + bwd + Adam) on B200, the metric BASELINE.json names.

Workload (config.workload): "wide" = BASELINE.json configs[3], the configuration the metric's
"1/2/4/8 B200" is quoted on -- F = 4096 in 16 modality blocks of 256, encoder [2048, 1024, 256],
untied, softsign, sigmoid-CE, global batch 65536, data-parallel over the ranks (strong scaling:
65536/N rows per rank, one sum-allreduce of the flat gradient per step).  One "step" = Philox
block-mask noise + forward + backward + fused Adam over one synthetic batch.

  value    device-resident DATASET, CUDA-event timed, whole-job samples/s (max over ranks).  One step = what
           MultimodalAutoencoder.train does per iteration (multimodal_autoencoder.py:565-590): sample the batch rows
           (data_funcs.py:167), block-mask noise (:668-702), forward, backward, Adam -- all on the device
           (mmae_train_step_resident)
  e2e      the same step fed from pinned HOST memory through the C ABI (mmae_train_step_host): every
           step's H2D copy and the D2H read of the loss are inside the timed region
  api      MultimodalAutoencoder.train(rng_mode='philox') timed through the drop-in Python class
  others   (default invocation, 1 GPU) short legs of the other BASELINE.json configs: small / cls / infer
  roofline tcgen05 GEMM family: algorithmic FLOPs / device time of those launches, measured live with
           CUDA events on the engine's stream during the timed region, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU port of the reference's step (oracle/cpu_port.py; the
           reference itself is Python-2 + TensorFlow-1.x and cannot run in this image) on all host cores,
           on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: F, block width list, layers, loss, global batch, head
    'wide': dict(F=4096, blocks=[256] * 16, layers=[2048, 1024, 256], B=65536, tie=False, act='softsign',
                 loss='sigmoid_cross_entropy'),                                         # BASELINE.json configs[3]
    'small': dict(F=320, blocks=[200, 20, 20, 30, 50], layers=[128, 64], B=65536, tie=False, act='softsign',
                  loss='sigmoid_cross_entropy'),                                        # configs[0] shape, large batch
    'cls': dict(F=320, blocks=[200, 20, 20, 30, 50], layers=[200, 100], B=4096, tie=False, act='relu',
                loss='sigmoid_cross_entropy', head=[50, 20], labels=3),                 # configs[1]
    'infer': dict(F=320, blocks=[200, 20, 20, 30, 50], layers=[128, 64], B=10_000_000, tie=False, act='softsign',
                  loss='sigmoid_cross_entropy'),                                        # configs[4]
}
# GEMM FLOPs per sample of one step (cls: reconstruction step + classification step)
FLOPS_PER_SAMPLE = {'wide': 112197632.0, 'small': 507904.0, 'cls': 880000.0 + 412360.0, 'infer': 196608.0}
BYTES_PER_SAMPLE = {'small': 1280.0, 'cls': 1292.0, 'infer': 2560.0}


def modality_names(n):
    base = ['call', 'sms', 'screen', 'location']
    return base + ['phys%02d' % i for i in range(n - 4)] if n > 4 else ['phys', 'call', 'sms', 'screen', 'location'][:n]


def workload_cfg(name):
    w = WORKLOADS[name]
    starts = [0]
    for b in w['blocks']:
        starts.append(starts[-1] + b)
    names = ['phys', 'call', 'sms', 'screen', 'location'] if name == 'small' else modality_names(len(w['blocks']))
    return w, starts, names


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith('active'):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def cpu_port_run(name, steps, warmup, sample_rows):
    """samples/s of the CPU port on `sample_rows` rows per step (bounded sample of the workload)."""
    import numpy as np
    import torch
    from oracle import mmae_oracle as O
    from oracle.cpu_port import CpuPort
    w, starts, names = workload_cfg(name)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(num_feats=w['F'], layer_sizes=list(w['layers']), modality_starts=starts, modality_names=names,
                         tie_weights=w['tie'], activation=w['act'], loss_func=w['loss'], learning_rate=1e-3,
                         cls_layer_sizes=w.get('head'), num_labels=w.get('labels', 3))
    rng = np.random.default_rng(0)
    port = CpuPort(cfg, O.init_params(cfg, rng), threads=cores)
    X = rng.uniform(0, 1, (sample_rows, w['F']))
    np.random.seed(0)
    if name == 'infer':
        Xm = X.copy()
        for m in range(len(w['blocks'])):
            Xm[rng.uniform(size=sample_rows) < 0.2, starts[m]:starts[m + 1]] = -1.0
        fn, what = (lambda: port.predict(Xm)), 'predict() + per-row fill loop'
    elif name == 'cls':
        Yc = (rng.uniform(size=(sample_rows, w['labels'])) < 0.5).astype(np.float64)

        def fn():
            _, t1 = port.step(X)
            _, t2 = port.cls_step(X, Yc)
            return 0.0, t1 + t2
        what = 'noise loop + recon step, noise loop + classification step'
    else:
        fn, what = (lambda: port.step(X)), 'noise loop + fwd + bwd + Adam'
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    t_noise = 0.0
    for _ in range(steps):
        _, tn = fn()
        t_noise += tn
    dt = time.perf_counter() - t0
    return dict(value=sample_rows * steps / dt, seconds=dt, noise_share=t_noise / dt, cores=cores,
                sample='%d steps x %d rows of the %s workload (of %d per step), %s'
                       % (steps, sample_rows, name, w['B'], what))


def reference_code_run(name, rows, steps=2):
    """The reference's OWN train loop (oracle/_ref: its multimodal_autoencoder.py converted to py3, on the TF-1 API shim in
    fp32 over torch's CPU kernels) on a small bounded sample.  Reported beside the port so that the choice of baseline is
    visible: the port is the faster CPU implementation of the two and is therefore the one used as the baseline."""
    import contextlib
    import io
    import numpy as np
    import torch
    from oracle.ref_loader import load_reference
    ref = load_reference(torch.float32)
    if ref is None:
        return None
    w, starts, names = workload_cfg(name)
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (rows, w['F']))
    dl = ref.make_loader(X, X[:200], starts, names)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.mmae.MultimodalAutoencoder(data_loader=dl, layer_sizes=list(w['layers']), variational=False, tie_weights=w['tie'],
                                           batch_size=rows, learning_rate=1e-3, activation_func=w['act'], loss_func=w['loss'],
                                           weight_initialization='normal', verbose=False)
        np.random.seed(0)
        m.train(1, record_every_nth=10 ** 9, save_every_nth=10 ** 9)
        t0 = time.perf_counter()
        m.train(steps, record_every_nth=10 ** 9, save_every_nth=10 ** 9)
        dt = time.perf_counter() - t0
    ref.tf.set_default_dtype(torch.float64)
    return {'value': rows * steps / dt, 'unit': 'samples/s',
            'sample': "%d steps x %d rows through the reference's own MultimodalAutoencoder.train() (one record step inside, :572-575)" % (steps, rows),
            'note': 'reference Python (py3-converted) on the TF-1 API shim, fp32 torch CPU kernels; TensorFlow itself is not installable here'}


METRIC = {'wide': 'MMAE train samples/sec (fwd+bwd+Adam)', 'small': 'MMAE train samples/sec (fwd+bwd+Adam)',
          'cls': 'MMAE train samples/sec (fwd+bwd+Adam), reconstruction step + classification-head step',
          'infer': 'MMAE fill-in inference samples/sec (reconstruction forward + missing-block fill)'}
CPU_ROWS = {'wide': 2048, 'small': 16384, 'cls': 4096, 'infer': 65536}      # cpu_baseline leg inside the GPU arm (10-30 s)
CPU_BUDGET_S = 200.0          # --impl reference: the whole --steps K --warmup W run stays within a few minutes


def run_reference(args):
    """The reference's CPU path for the same workload, on all host cores (rank 0 only).  TensorFlow cannot run in this
    image, so this is the CPU port (oracle/cpu_port.py: the same op sequence on torch's multithreaded fp32 CPU kernels,
    preceded by the reference's own per-row NumPy noise loop).  Rows per step: the workload's full batch when K + W such
    steps fit CPU_BUDGET_S (measured with a calibration step), otherwise the largest power-of-two sample that does."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    name = args.workload if args.workload != 'grid' else 'small'
    w, _, _ = workload_cfg(name)
    full = args.rows or (w['B'] if name != 'infer' else 1_000_000)
    cal_rows = min(full, 1024)
    cal = cpu_port_run(name, 1, 1, cal_rows)
    per_row = cal['seconds'] / cal_rows
    rows = full
    while rows > 256 and per_row * rows * (args.steps + max(args.warmup, 1)) > CPU_BUDGET_S:
        rows //= 2
    r = cpu_port_run(name, args.steps, max(args.warmup, 1), rows)
    own = None
    if name in ('wide', 'small'):
        try:
            own = reference_code_run(name, 1024 if name == 'wide' else 8192)
        except Exception as e:      # noqa: BLE001 -- an extra, never the line itself
            own = {'error': repr(e)[:200]}
    line = {
        'impl': 'reference', 'metric': METRIC[name], 'value': r['value'], 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * r['seconds'] / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': name, 'global_batch': w['B'], 'features': w['F'], 'modality_blocks': len(w['blocks']),
                   'encoder': w['layers'], 'head': w.get('head'), 'loss': w['loss'], 'activation': w['act'],
                   'rows_per_step': rows, 'same_rows_as_gpu_arm': rows == w['B'],
                   'note': 'CPU port of the reference step on all host cores (the TensorFlow-1.x reference cannot run here)'},
        'cpu_baseline': {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample'],
                         'noise_loop_share': r['noise_share'], 'reference_code_via_shim': own},
        'e2e': {'value': r['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def measure_tf32_peak(torch):
    """cuBLAS tf32 8192^3 on this GPU, back to back for ~0.5 s: the tensor-pipe ceiling of kind::tf32 kernels."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device='cuda'); b = torch.randn((n, n), device='cuda')
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b)
        e1.record(); torch.cuda.synchronize()
        return 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) / 1e3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def make_engine(name, B, precision, rng_seed=0):
    import numpy as np
    from multimodalautoencoder_b200 import Engine, EngineConfig
    w, starts, names = workload_cfg(name)
    cfg = EngineConfig(num_feats=w['F'], layer_sizes=list(w['layers']), modality_starts=starts, modality_names=names,
                       tie_weights=w['tie'], variational=False, activation=w['act'], loss_func=w['loss'],
                       learning_rate=1e-3, weight_penalty=0.0, seed=0, precision=precision, max_batch=B,
                       cls_layer_sizes=w.get('head'), num_labels=w.get('labels', 3))
    eng = Engine(cfg)
    rng = np.random.default_rng(rng_seed)       # random-init weights of the named architecture ('normal' init, :44)
    for vname, shp in eng.variables():
        if len(shp) == 1:
            eng.set_variable(vname, np.full(shp, 0.1, np.float32))
        else:
            eng.set_variable(vname, (np.clip(rng.standard_normal(shp), -2, 2) / np.sqrt(shp[0])).astype(np.float32))
    return eng


def run_leg(name, args, rank, world, local, dist, steps, warmup, want_e2e, clocks=True):
    """One workload on this rank's GPU.  Returns the dict of measurements (rank 0 assembles the JSON line)."""
    import numpy as np
    import torch
    from multimodalautoencoder_b200 import Engine
    w, starts, names = workload_cfg(name)
    train = name != 'infer'
    Bg = args.rows or w['B']
    assert Bg % world == 0
    B = Bg // world
    F = w['F']
    head = w.get('head')
    eng = make_engine(name, B, args.precision)
    dp = world > 1 and train          # inference shards rows with no collective
    if dp:
        idt = torch.zeros(128, dtype=torch.uint8, device='cuda')
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        eng.set_shard(Bg, rank * B)
    # synthetic SNAPSHOT-shaped data: U[0,1) features.  Training: a device-resident dataset of 2 x B rows per
    # rank (each rank holds its own shard and samples it locally); inference: this rank's rows of the matrix to fill.
    gen = torch.Generator(device='cuda').manual_seed(1234 + rank)
    n_data = max(2 * B, int(300e6 / (4 * F))) if train else B      # dataset > 2 x L2 (126 MB): sampled rows come from HBM
    X = torch.rand((n_data, F), device='cuda', generator=gen)
    Y = None
    if head:
        Y = (torch.rand((n_data, w['labels']), device='cuda', generator=gen) < 0.5).float()
    if not train:        # each modality of each row independently missing (-1.0) with p = 0.2
        drop = torch.rand((B, len(w['blocks'])), device='cuda', generator=gen) < 0.2
        for m in range(len(w['blocks'])):
            X[:, starts[m]:starts[m + 1]] = torch.where(drop[:, m:m + 1], torch.full_like(X[:, starts[m]:starts[m + 1]], -1.0),
                                                        X[:, starts[m]:starts[m + 1]])
        del drop
    filled = torch.empty((B, F), device='cuda') if not train else None
    if train:
        eng.set_dataset_device(0, X, Y)
        if head:
            eng.set_dataset_device(1, X, Y)

    def step(i):
        eng.set_rng_step(i)
        if name == 'infer':
            eng.forward_into(X, filled=filled)
        else:
            eng.train_step_resident(0, B, gen_noise=True)
            if name == 'cls':
                eng.train_step_resident(1, B, gen_noise=True, classification=True)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local) if (rank == 0 and clocks) else None
    if sampler:
        sampler.start()
    # The wide workload is timed with per-launch CUDA events on (profiling) inside the timed region.  The small-batch
    # workloads replay a captured CUDA graph per step, which per-launch events would break up: they are timed
    # un-instrumented, and the kernel times for the roofline come from a second, instrumented pass of the same steps.
    live_profile = name in ('wide', 'infer') and world == 1      # (multi-GPU steps replay CUDA graphs at small per-rank batches)
    if live_profile:
        eng.set_profiling(True)
    l0, c0, g0 = eng.kernel_launches, eng.chain_launches, eng.graph_replays
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        step(warmup + i)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches - l0
    chains = eng.chain_launches - c0
    replays = eng.graph_replays - g0
    if not live_profile:
        eng.set_profiling(True)
        for i in range(steps):
            step(warmup + steps + i)
        sync_all()
    prof = eng.read_profile()
    eng.set_profiling(False)
    sc = eng.scalars()
    clk = sampler.stop() if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = Bg * steps / (ms / 1e3)

    # ---- end to end: host-fed steps through the C ABI (pinned memory, H2D + result D2H inside the timing)
    e2e = None
    if want_e2e:
        Be = B if train else min(B, 1_000_000)       # inference: a bounded 1 M-row slice per call (host RAM)
        hx = [torch.rand((Be, F)).pin_memory() for _ in range(2)]
        hy = (torch.rand((Be, w['labels'])) < 0.5).float().pin_memory() if head else None
        hs = torch.zeros((steps + warmup + 1, 8), dtype=torch.float64).pin_memory()
        hout = torch.empty((Be, F)).pin_memory() if not train else None

        def estep(i):
            eng.set_rng_step(2000 + i)
            if train:
                eng.train_step_host(hx[i % 2], gen_noise=True)
                if name == 'cls':
                    eng.cls_train_step_host(hx[i % 2], hy, gen_noise=True)
                eng.read_scalars_async(hs[i])
            else:
                eng.forward_host(hx[i % 2].numpy(), filled=True, out={'filled': hout.numpy()})

        for i in range(min(warmup, 3)):
            estep(i)
        sync_all()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        for i in range(steps):
            estep(i)
        t1.record()
        eng.synchronize()
        sync_all()
        wall_ms = (time.perf_counter() - w0) * 1e3
        ems = max(t0.elapsed_time(t1), wall_ms)         # copies run on a side stream: take the wall clock if larger
        if dist is not None:
            t = torch.tensor([ems], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        rows_e = Be * world
        e2e = {'value': rows_e * steps / (ems / 1e3), 'unit': 'samples/s',
               'h2d_bytes_per_step': rows_e * F * 4 + (rows_e * w['labels'] * 4 if head else 0),
               'd2h_bytes_per_step': 64 * world if train else rows_e * F * 4, 'ms_per_step': ems / steps,
               'rows_per_step': rows_e, 'api': 'mmae_train_step_host / mmae_forward_host (C ABI, pinned host buffers)'}
        if train:
            e2e['last_loss'] = float(hs[steps - 1][0])
        del hx, hy, hout
    eng.close()
    del X, Y, filled
    torch.cuda.empty_cache()
    return dict(name=name, w=w, starts=starts, Bg=Bg, B=B, F=F, n_data=n_data, head=head, train=train, ms=ms, value=value, launches=launches,
                chains=chains, replays=replays, prof=prof, sc=sc, clocks=clk, e2e=e2e, live_profile=live_profile, steps=steps,
                warmup=warmup, world=world)


def roofline_of(m, peaks, pk, tf32_peak):
    name, world, value, ms, prof = m['name'], m['world'], m['value'], m['ms'], m['prof']
    peak_tf = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops')))
    peak_bw = float(peaks.get('hbm_gbs'))
    kern_ms = prof['gemm_ms']
    achieved_tf = prof['gemm_flops'] / (kern_ms / 1e3) / 1e12 if kern_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(name)
    if name == 'infer':
        # dominant kernel = the whole-network chain kernel; algorithmic bytes = read X + write filled X
        algo = BYTES_PER_SAMPLE[name] * m['B'] * m['steps']
        ach = algo / (kern_ms / 1e3) / 1e9 if kern_ms > 0 else 0.0
        return {'bound': 'hbm', 'achieved': ach, 'peak': peak_bw, 'unit': 'GB/s', 'frac': ach / peak_bw, 'traffic': traffic,
                'kernel': 'chain_tc_kernel (whole network, activations in TMEM)', 'peak_source': 'copy bandwidth, %s (MEASURED_PEAKS.json)' % pk,
                'algorithmic_bytes_per_launch': BYTES_PER_SAMPLE[name] * m['B'],
                'kernel_share_of_step': kern_ms / ms if ms > 0 else None, 'kernel_launches': prof['gemm_launches'],
                'step_frac_of_peak': (BYTES_PER_SAMPLE[name] * value / world / 1e9) / peak_bw, 'tensor_tflops': achieved_tf}
    r = {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf,
         'traffic': traffic,
         'kernel': 'tcgen05 kind::tf32 family: gemm_tc2_kernel / gemm_tc_kernel (per-layer GEMMs), chain_tc_kernel<fwd|bwd> '
                   '(whole network / all dgrads), wgrad_group_kernel (all weight gradients)',
         'peak_source': 'bf16 dense sustained, %s (MEASURED_PEAKS.json); kind::tf32 issues at half the bf16 rate, '
                        'so 0.5 is this kernel family\'s ceiling against this denominator' % pk,
         'tf32_peak_measured': tf32_peak, 'frac_of_tf32_peak': achieved_tf / tf32_peak if tf32_peak > 0 else None,
         'tf32_peak_source': 'cuBLAS tf32 8192^3 timed in this run (SURVEY 8d: kind::tf32 kernels are normalised by a measured TF32 peak)',
         'gemm_share_of_step': kern_ms / ms if ms > 0 else None,
         'kernel_times': 'CUDA events inside the timed region' if m['live_profile'] else 'CUDA events in a second pass of the same steps (the timed pass replays CUDA graphs)',
         'gemm_launches': prof['gemm_launches'],
         'step_frac_of_peak': (FLOPS_PER_SAMPLE[name] * value / world / 1e12) / peak_tf,
         'step_frac_of_tf32_peak': (FLOPS_PER_SAMPLE[name] * value / world / 1e12) / tf32_peak if tf32_peak > 0 else None}
    if name in BYTES_PER_SAMPLE:      # small configs: the HBM bound beside the tensor bound (SURVEY 8d reports both)
        r['hbm_bound'] = {'algorithmic_GBps': BYTES_PER_SAMPLE[name] * value / world / 1e9, 'peak': peak_bw,
                          'frac': BYTES_PER_SAMPLE[name] * value / world / 1e9 / peak_bw}
        r['binding_roof'] = 'tensor: 60 %% of the HBM roof (%.2f G samples/s) lies above the tf32 compute bound (%.2f G samples/s)' % (
            0.6 * peak_bw / BYTES_PER_SAMPLE[name], tf32_peak * 1e3 / FLOPS_PER_SAMPLE[name]) if tf32_peak > 0 else None
    return r


def config_of(m):
    w, B, F = m['w'], m['B'], m['F']
    return {'workload': m['name'], 'global_batch': m['Bg'], 'features': F, 'modality_blocks': len(w['blocks']),
            'encoder': w['layers'], 'head': m['head'], 'loss': w['loss'], 'activation': w['act'],
            'parallelism': ('dp%d' % m['world']) if m['train'] else ('rows sharded over %d GPU(s), no collective' % m['world']),
            'batch_source': ('device-resident dataset of %d rows per rank, %d rows sampled per step on the device (Philox)' % (m['n_data'], B))
                            if m['train'] else 'device-resident matrix',
            'l2_policy': 'inputs larger than L2: %s %.2f GB per rank (L2 = 126 MB), %s' % (
                'dataset' if m['train'] else 'matrix', m['n_data'] * F * 4 / 1e9,
                'rows sampled at random every step' if m['train'] else 'streamed once per pass'),
            'noise': 'philox block-mask + 5% zero noise drawn on device every step' if m['train'] else
                     'each modality block of each row missing (-1) with p = 0.2'}


def api_leg(name, steps):
    """MultimodalAutoencoder.train(rng_mode='philox') through the drop-in Python class at the workload's shape."""
    import numpy as np
    import torch
    from multimodalautoencoder_b200 import MultimodalAutoencoder
    from types import SimpleNamespace
    w, starts, names = workload_cfg(name)
    B, F = w['B'], w['F']
    rng = np.random.default_rng(7)
    n = 2 * B
    X = rng.random((n, F), dtype=np.float32)
    dl = SimpleNamespace(train_X=X, val_X=X[:400], test_X=X[:400], train_Y=None, val_Y=None, test_Y=None, num_feats=F,
                         num_labels=None, modality_start_indices=list(starts), modality_names=list(names),
                         num_modalities=len(names), fold=None, wanted_feats=None,
                         get_unsupervised_train_batch=lambda b: X[np.random.choice(n, size=b)],
                         get_unsupervised_val_batch=lambda b: X[np.random.choice(400, size=b)])
    m = MultimodalAutoencoder(data_loader=dl, layer_sizes=list(w['layers']), variational=False, tie_weights=w['tie'],
                              batch_size=B, learning_rate=1e-3, activation_func=w['act'], loss_func=w['loss'],
                              weight_initialization='normal', verbose=False, rng_mode='philox', precision='tf32')
    m.train(3, record_every_nth=10 ** 9, save_every_nth=10 ** 9)          # uploads the dataset, captures / warms the step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.train(steps, record_every_nth=steps, save_every_nth=10 ** 9)        # one record step (loss read-back) in the timed region
    m.engine.synchronize()
    dt = time.perf_counter() - t0
    out = {'value': B * steps / dt, 'unit': 'samples/s', 'ms_per_step': 1e3 * dt / steps, 'steps': steps,
           'call': "MultimodalAutoencoder(..., rng_mode='philox').train(%d)" % steps, 'train_loss_per_sample': float(m.train_loss[-1])}
    m.close()
    return out


def run_grid(args, rank, world, local, dist):
    """BASELINE.json configs[2]: 64 MMAE settings fitted independently, round-robin over the ranks, no collective."""
    import itertools
    import numpy as np
    import torch
    from multimodalautoencoder_b200 import MultimodalAutoencoder
    from multimodalautoencoder_b200.data_funcs import DataLoader
    from multimodalautoencoder_b200.synthetic import make_frame
    steps = args.grid_steps
    df = make_frame(4000, seed=3)
    dl = DataLoader(df=df, supervised=False, cross_validation=False, normalize_and_fill=False, suppress_output=True)
    settings = list(itertools.product([[1000, 100], [500, 100], [300, 100], [128, 64]], [True, False], [1.0, 0.5], [0.0, 0.001],
                                      ['softsign', 'relu']))
    mine = settings[rank::world]
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    losses = []
    for (arch, tie, keep, lam, act) in mine:
        m = MultimodalAutoencoder(data_loader=dl, layer_sizes=arch, variational=False, tie_weights=tie, batch_size=20,
                                  learning_rate=1e-3, dropout_prob=keep, weight_penalty=lam, activation_func=act,
                                  loss_func='sigmoid_cross_entropy', weight_initialization='normal', verbose=False,
                                  rng_mode='philox', precision=args.precision)
        m.train(steps, record_every_nth=max(steps // 2, 1), save_every_nth=10 ** 9)
        losses.append(m.get_performance_on_data(dl.val_X))
        m.close()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps({
        'metric': 'MMAE hyper-parameter grid: 64 settings x %d train steps at batch 20, wall time' % steps,
        'value': 64 * steps * 20 / dt, 'unit': 'samples/s', 'n_gpus': world, 'steps': steps, 'warmup': 0,
        'ms_per_step': 1e3 * dt / (steps * ((64 + world - 1) // world)), 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'tf32', 'data': 'synthetic',
        'config': {'workload': 'grid', 'settings': 64, 'steps_per_fit': steps, 'batch': 20,
                   'grid': 'enc {[1000,100],[500,100],[300,100],[128,64]} x tie {T,F} x keep {1.0,0.5} x lambda {0,.001} x act {softsign,relu}',
                   'parallelism': 'settings round-robin over %d GPU(s), no collective (generic_wrapper.py:246-256)' % world},
        'grid_wall_s': dt, 'fits_per_s': 64 / dt, 'mean_val_loss_rank0': float(np.mean(losses)), 'gpu_launches': None}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='wide', choices=sorted(WORKLOADS) + ['grid'])
    ap.add_argument('--precision', default='tf32', choices=['tf32', 'fp32'])
    ap.add_argument('--rows', type=int, default=0, help='override the rows per step (infer / small)')
    ap.add_argument('--grid-steps', type=int, default=300, help='train steps per setting of --workload grid')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-others', action='store_true', help='skip the short small / cls / infer legs of the default invocation')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if args.workload == 'grid':
        return run_grid(args, rank, world, local, dist)

    name = args.workload
    m = run_leg(name, args, rank, world, local, dist, args.steps, args.warmup, not args.no_e2e)
    others = None
    api = None
    if world == 1 and name == 'wide' and not args.no_others and not args.rows:
        others = {}
        saved_rows = args.rows
        for oname in ('small', 'cls', 'infer'):
            try:
                om = run_leg(oname, args, rank, world, local, None, 20, 5, False, clocks=False)
                others[oname] = om
            except Exception as e:          # a failed side leg must not take the headline line with it
                others[oname] = {'error': repr(e)[:300]}
        args.rows = saved_rows
        try:
            api = api_leg('wide', 10)
        except Exception as e:
            api = {'error': repr(e)[:300]}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks, pk = measured_peaks()
    tf32_peak = measure_tf32_peak(torch)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_port_run(name, 3, 1, CPU_ROWS[name])
        cpu = {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample'],
               'noise_loop_share': r['noise_share']}
    line = {
        'metric': METRIC[name], 'value': m['value'], 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': m['ms'] / args.steps, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'tf32' if args.precision == 'tf32' else 'f32',
        'data': 'synthetic', 'config': config_of(m),
        'e2e': m['e2e'], 'gpu_launches': m['launches'], 'whole_network_launches': m['chains'], 'graph_replays': m['replays'],
        'clocks': m['clocks'], 'roofline': roofline_of(m, peaks, pk, tf32_peak), 'cpu_baseline': cpu,
    }
    if m['train']:
        line['final_loss_per_sample'] = m['sc']['recon_loss'] / m['Bg']
    if api is not None:
        line['api'] = api
    if others is not None:
        line['others'] = {}
        for oname, om in others.items():
            if 'error' in om:
                line['others'][oname] = om
                continue
            line['others'][oname] = {'metric': METRIC[oname], 'value': om['value'], 'unit': 'samples/s', 'ms_per_step': om['ms'] / om['steps'],
                                     'steps': om['steps'], 'warmup': om['warmup'], 'config': config_of(om), 'gpu_launches': om['launches'],
                                     'whole_network_launches': om['chains'], 'graph_replays': om['replays'],
                                     'roofline': roofline_of(om, peaks, pk, tf32_peak)}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
