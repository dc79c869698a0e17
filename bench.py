#!/usr/bin/env python
"""bench.py -- MMAE train samples/sec (fwd + bwd + Adam) on B200, the metric BASELINE.json names.

Workload (config.workload): "wide" = BASELINE.json configs[3], the configuration the metric's
"1/2/4/8 B200" is quoted on -- F = 4096 in 16 modality blocks of 256, encoder [2048, 1024, 256],
untied, softsign, sigmoid-CE, global batch 65536, data-parallel over the ranks (strong scaling:
65536/N rows per rank, one sum-allreduce of the flat gradient per step).  One "step" = Philox
block-mask noise + forward + backward + fused Adam over one synthetic batch.

  value    device-resident inputs, CUDA-event timed, whole-job samples/s (max over ranks)
  e2e      the same step fed from pinned HOST memory through the C ABI (mmae_train_step_host): every
           step's H2D copy and the D2H read of the loss are inside the timed region
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
                 loss='sigmoid_cross_entropy'),
    'small': dict(F=320, blocks=[200, 20, 20, 30, 50], layers=[128, 64], B=65536, tie=False, act='softsign',
                  loss='sigmoid_cross_entropy'),
}
FLOPS_PER_SAMPLE = {'wide': 112197632.0, 'small': 507904.0}


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
                         tie_weights=w['tie'], activation=w['act'], loss_func=w['loss'], learning_rate=1e-3)
    rng = np.random.default_rng(0)
    port = CpuPort(cfg, O.init_params(cfg, rng), threads=cores)
    X = rng.uniform(0, 1, (sample_rows, w['F']))
    np.random.seed(0)
    for _ in range(warmup):
        port.step(X)
    t0 = time.perf_counter()
    t_noise = 0.0
    for _ in range(steps):
        _, tn = port.step(X)
        t_noise += tn
    dt = time.perf_counter() - t0
    return dict(value=sample_rows * steps / dt, seconds=dt, noise_share=t_noise / dt, cores=cores,
                sample='%d steps x %d rows of the %s workload (of %d per step), noise loop + fwd + bwd + Adam'
                       % (steps, sample_rows, name, w['B']))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    name = args.workload
    w, _, _ = workload_cfg(name)
    rows = 2048 if name == 'wide' else 16384
    r = cpu_port_run(name, args.steps, max(args.warmup, 1), rows)
    line = {
        'impl': 'reference', 'metric': 'MMAE train samples/sec (fwd+bwd+Adam)', 'value': r['value'], 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * r['seconds'] / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': name, 'global_batch': w['B'], 'features': w['F'], 'encoder': w['layers'],
                   'note': 'CPU port of the reference step (TensorFlow-1.x reference cannot run here); bounded sample'},
        'cpu_baseline': {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample'],
                         'noise_loop_share': r['noise_share']},
        'e2e': {'value': r['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='wide', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='tf32', choices=['tf32', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
    import torch
    from multimodalautoencoder_b200 import Engine, EngineConfig

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    name = args.workload
    w, starts, names = workload_cfg(name)
    Bg = w['B']
    assert Bg % world == 0
    B = Bg // world
    F = w['F']
    cfg = EngineConfig(num_feats=F, layer_sizes=list(w['layers']), modality_starts=starts, modality_names=names,
                       tie_weights=w['tie'], variational=False, activation=w['act'], loss_func=w['loss'],
                       learning_rate=1e-3, weight_penalty=0.0, seed=0, precision=args.precision, max_batch=B)
    eng = Engine(cfg)
    # random-init weights of the named architecture ('normal' init, multimodal_autoencoder.py:44)
    rng = np.random.default_rng(0)
    for vname, shp in eng.variables():
        if len(shp) == 1:
            eng.set_variable(vname, np.full(shp, 0.1, np.float32))
        else:
            eng.set_variable(vname, (np.clip(rng.standard_normal(shp), -2, 2) / np.sqrt(shp[0])).astype(np.float32))
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device='cuda')
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
        eng.set_shard(Bg, rank * B)
    # synthetic SNAPSHOT-shaped data: U[0,1) features, this rank's rows of the global batch
    gen = torch.Generator(device='cuda').manual_seed(1234 + rank)
    X = torch.rand((B, F), device='cuda', generator=gen)

    def step(i):
        eng.set_rng_step(i)
        eng.gen_noise(B, rank * B)
        eng.train_step(X, noise=True, keep=1.0)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.set_profiling(True)
    l0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches - l0
    prof = eng.read_profile()
    eng.set_profiling(False)
    loss = eng.scalars()['recon_loss']
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = Bg * args.steps / (ms / 1e3)

    # ---- end to end: host-fed steps through the C ABI (pinned memory, H2D + loss D2H inside the timing)
    e2e = None
    if not args.no_e2e:
        hx = [torch.rand((B, F)).pin_memory() for _ in range(2)]
        hs = torch.zeros((args.steps + args.warmup + 1, 8), dtype=torch.float64).pin_memory()
        for i in range(min(args.warmup, 3)):
            eng.set_rng_step(1000 + i)
            eng.train_step_host(hx[i % 2], gen_noise=True)
            eng.read_scalars_async(hs[i])
        sync_all()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        for i in range(args.steps):
            eng.set_rng_step(2000 + i)
            eng.train_step_host(hx[i % 2], gen_noise=True)
            eng.read_scalars_async(hs[i])
        t1.record()
        eng.synchronize()
        sync_all()
        wall_ms = (time.perf_counter() - w0) * 1e3
        ems = max(t0.elapsed_time(t1), wall_ms)         # copies run on a side stream: take the wall clock if larger
        if dist is not None:
            t = torch.tensor([ems], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {'value': Bg * args.steps / (ems / 1e3), 'unit': 'samples/s', 'h2d_bytes_per_step': Bg * F * 4,
               'd2h_bytes_per_step': 64 * world, 'ms_per_step': ems / args.steps,
               'last_loss': float(hs[args.steps - 1][0])}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks, pk = measured_peaks()
    peak_tf = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops')))
    achieved = prof['gemm_flops'] / (prof['gemm_ms'] / 1e3) / 1e12 if prof['gemm_ms'] > 0 else 0.0
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                'traffic': None, 'kernel': 'gemm_tc_kernel (tcgen05 kind::tf32)',
                'peak_source': 'bf16 dense sustained, %s (MEASURED_PEAKS.json); kind::tf32 issues at half the bf16 rate, '
                               'so 0.5 is this kernel family\'s ceiling against this denominator' % pk,
                'gemm_share_of_step': prof['gemm_ms'] / ms if ms > 0 else None,
                'gemm_launches': prof['gemm_launches'],
                'step_frac_of_peak': (FLOPS_PER_SAMPLE[name] * value / 1e12) / peak_tf}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_port_run(name, 3, 1, 2048 if name == 'wide' else 16384)
        cpu = {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample'],
               'noise_loop_share': r['noise_share']}
    line = {
        'metric': 'MMAE train samples/sec (fwd+bwd+Adam)', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'tf32' if args.precision == 'tf32' else 'f32',
        'data': 'synthetic',
        'config': {'workload': name, 'global_batch': Bg, 'features': F, 'modality_blocks': len(w['blocks']),
                   'encoder': w['layers'], 'loss': w['loss'], 'activation': w['act'], 'parallelism': 'dp%d' % world,
                   'l2_policy': 'inputs larger than L2 (batch X = %.2f GB per rank, re-read every step)' % (B * F * 4 / 1e9),
                   'noise': 'philox block-mask + 5% zero noise drawn on device every step'},
        'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
        'final_loss_per_sample': loss / Bg,
    }
    print(json.dumps(line))


if __name__ == '__main__':
    main()
