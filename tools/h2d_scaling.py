"""Per-rank pinned host -> device copy rate with N ranks copying at the same time (development aid / evidence for the
e2e scaling limit):  torchrun --nproc-per-node N tools/h2d_scaling.py [--bind]
Every rank copies a 256 MiB pinned buffer to its GPU 10 times after a barrier; the table lists the per-rank and
aggregate GB/s, the GPUs' PCIe link and the NUMA node / local cpus sysfs reports for each device."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, '.')
from detprocess_b200.utils.affinity import bind_to_gpu, gpu_local_cpus

rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
info = {}
if '--bind' in sys.argv:
    info = bind_to_gpu(local, local, world)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)                                  # first touch by this (bound) process
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
rate = 10 * n / dt / 1e9
cpus, node = gpu_local_cpus(local)
row = (rank, rate, node, len(cpus), info.get('bound_cpus', [None])[0] if info.get('bound_cpus') else None, len(info.get('bound_cpus', [])))
rows = [None] * world
if world > 1:
    dist.all_gather_object(rows, row)
else:
    rows = [row]
if rank == 0:
    import subprocess
    link = subprocess.run(['nvidia-smi', '--query-gpu=index,pcie.link.gen.current,pcie.link.width.current', '--format=csv,noheader'],
                          capture_output=True, text=True).stdout.strip().replace('\n', ' | ')
    tot = sum(r[1] for r in rows)
    print(f'N={world} bind={"--bind" in sys.argv}: aggregate {tot:.1f} GB/s, per rank ' + ' '.join(f'{r[1]:.1f}' for r in rows) +
          f' | numa nodes {[r[2] for r in rows]} local cpus {[r[3] for r in rows]} bound {[(r[4], r[5]) for r in rows]} | pcie {link}')
if world > 1:
    dist.destroy_process_group()
