import torch, time
for name, dt, n in (('i16 268MB', torch.int16, 4096*32768), ('f64 1GiB', torch.float64, 4096*32768), ('i16 38MB x8', torch.int16, 592*32768)):
    h = torch.empty(n, dtype=dt).pin_memory()
    d = torch.empty(n, dtype=dt, device='cuda')
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    reps = 8 if 'x8' in name else 1
    t0 = time.perf_counter()
    for _ in range(5 * reps): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt_ = (time.perf_counter() - t0) / (5 * reps)
    print(f'{name}: {h.numel()*h.element_size()/dt_/1e9:.1f} GB/s ({dt_*1e3:.2f} ms)')
