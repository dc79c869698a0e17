"""All-reduce time of the wide model's gradient buckets on the GPUs of this box (NCCL via torch.distributed), for the
NCCL_MAX_CTAS given in the environment.  Run under torchrun."""
import os
import torch
import torch.distributed as dist

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank = dist.get_rank()
sizes = {'sums (8 floats)': 8, 'L2 bucket 1 MB': 262144 + 256, 'L1 bucket 8.4 MB': 2048 * 1024 + 1024,
         'L0 bucket 33.5 MB': 4096 * 2048 + 2048, 'all 86 MB': 21506304}
for name, n in sizes.items():
    t = torch.ones(n, device='cuda')
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dist.all_reduce(t)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    if rank == 0:
        print('NCCL_MAX_CTAS=%s world=%d  %-20s %8.1f us   algbw %6.1f GB/s' % (os.environ.get('NCCL_MAX_CTAS', 'default'), dist.get_world_size(),
                                                                                name, ms * 1e3, n * 4 / ms / 1e6), flush=True)
dist.destroy_process_group()
