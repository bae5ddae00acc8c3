"""World-size-2 gloo test (CPU) of the N>1 host path: contiguous event shards per rank, no
data-path collective, one gather of the small feature tables."""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from detprocess_b200.process.features import shard_range, gather_frames, dist_info


def test_shard_range_partitions_events():
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_events, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    assert dist_info() == (rank, world)
    lo, hi = shard_range(n_events, rank, world)
    # stand-in for the per-rank feature table: a deterministic function of the event number
    ev = np.arange(lo, hi)
    df = pd.DataFrame({'event_number': ev, 'amp_x': np.sin(ev) * 1e-7, 'rank': rank})
    full = gather_frames(df)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_events', [101, 1])
def test_two_rank_gather_equals_single(n_events):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_events, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ev = np.arange(n_events)
    assert np.array_equal(full['event_number'].to_numpy(), ev)          # rank order == event order
    assert np.array_equal(full['amp_x'].to_numpy(), np.sin(ev) * 1e-7)  # sharded == single
    assert sorted(set(full['rank'])) == ([0, 1] if n_events > 1 else [1])


def _psd_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from detprocess_b200.core.noise import allreduce_sums, two_sided_from_sums
    from oracle import psd as P
    n, fs, n_tr = 1024, 1.25e6, 90
    tr = np.random.default_rng(21).standard_normal((n_tr, n)) * 1e-10
    cut = np.random.default_rng(22).random(n_tr) < 0.8
    lo, hi = shard_range(n_tr, rank, world)
    s, c = P.periodogram_sums(tr[lo:hi], cut[lo:hi])       # what one GPU accumulates for its shard
    sums, count = allreduce_sums(torch.from_numpy(s.copy()), torch.tensor([c], dtype=torch.int64))
    psd = two_sided_from_sums(sums, int(count.item()), n, fs).numpy()
    if rank == 0:
        q.put((psd, int(count.item()), P.calc_psd(tr, fs, cut)[1], int(cut.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_psd_allreduce_equals_single():
    """C5 host path: per-rank periodogram sums + counts all-reduced == PSD of the whole set."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_psd_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    psd, count, ref, nref = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert count == nref
    assert np.allclose(psd, ref, rtol=1e-12, atol=0)


def _csd_component_sums(x):
    """what one GPU accumulates for its shard, in the CSDPlan.sums layout"""
    ne, n, N = x.shape
    X = np.fft.rfft(x, axis=-1)
    rows = [np.sum(np.abs(X[:, a]) ** 2, axis=0) for a in range(n)]
    for a in range(n):
        for b in range(a + 1, n):
            z = np.sum(X[:, a] * np.conj(X[:, b]), axis=0)
            rows += [z.real, z.imag]
    return np.stack(rows)


def _csd_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from detprocess_b200.core.noise import allreduce_sums, csd_from_sums
    from oracle import psd as P
    n, fs, n_ev = 512, 1.25e6, 70
    x = np.random.default_rng(31).standard_normal((n_ev, 3, n)) * 1e-10
    x[:, 1] += 0.5 * np.roll(x[:, 0], 2, axis=-1)             # correlated, delayed: complex off-diagonal terms
    lo, hi = shard_range(n_ev, rank, world)
    sums, count = allreduce_sums(torch.from_numpy(_csd_component_sums(x[lo:hi])), torch.tensor([hi - lo], dtype=torch.int64))
    csd = csd_from_sums(sums.numpy(), int(count.item()), 3, n, fs)
    if rank == 0:
        q.put((csd, P.calc_csd(x, fs)[1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_csd_allreduce_equals_single():
    """Noise.calc_csd host path: per-rank component sums all-reduced and unfolded == CSD of the whole set."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_csd_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    csd, ref = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.allclose(csd, ref, rtol=1e-11, atol=1e-40)


def _autocut_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from detprocess_b200.core.noise import global_autocut, sigma_clip_cut
    rng = np.random.default_rng(77)
    stats = rng.standard_normal((101, 3))
    stats[::9] += 15.0                                 # outliers: pile-up / baseline jumps
    lo, hi = shard_range(len(stats), rank, world)
    mask = global_autocut(torch.from_numpy(stats[lo:hi]), sigma_clip_cut)
    got = [None] * world
    dist.all_gather_object(got, (lo, hi, mask))
    if rank == 0:
        full = np.zeros(len(stats), dtype=bool)
        for a, b, m in got:
            full[a:b] = m
        q.put((full, sigma_clip_cut(stats)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_global_autocut_equals_single():
    """the population-level noise cut (reference noise.py:331) on rank-sharded statistics: per-rank statistics ->
    all_gather -> identical global mask on every rank; sharded == single process (SURVEY hard part 6)"""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_autocut_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, ref = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(full, ref)
    assert (~ref)[::9].all() and ref.sum() > 20          # every outlier is gone, a population survives
    # a per-rank cut would NOT be the same thing: the two halves have different populations
    from detprocess_b200.core.noise import sigma_clip_cut
    rng = np.random.default_rng(77)
    stats = rng.standard_normal((101, 3))
    stats[::9] += 15.0
    assert ref.shape == (101,) and sigma_clip_cut(stats[:50]).shape == (50,)
