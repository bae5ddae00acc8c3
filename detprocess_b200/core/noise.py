"""
Noise PSD estimation on the GPU -- the ``Noise.calc_psd`` hot path of the reference
(detprocess/core/noise.py:216-370), config C5 of BASELINE.json.

The reference holds every trace in RAM and calls ``qp.calc_psd(traces[cut], fs,
folded_over=False)`` (noise.py:344).  Here traces stream through ``PSDPlan.accumulate`` in
batches (each trace is read from HBM exactly once); every rank accumulates
``sum |fft(x)_k|^2`` for its shard, the ``[N/2+1]`` sums and the accepted-trace counts are
all-reduced (NCCL over NVLink when ``torch.distributed`` is initialised with the nccl
backend) and the two-sided PSD is formed on every rank.

The noise autocut (``qp.autocuts_noise``, noise.py:331) is a POPULATION-level, data-dependent cut: it does not
decompose over ranks.  ``trace_statistics`` + ``global_autocut`` give the two-pass form that does (SURVEY.md hard part 6):
pass 1 computes per-trace statistics on every rank (window reductions of the fused kernels, bit-identical to numpy, so
the numbers do not depend on the sharding), one ``all_gather`` makes the whole population visible to every rank, every
rank derives the SAME global mask and keeps its own slice; pass 2 accumulates the periodograms of the accepted traces.
The cut function itself is pluggable: QETpy's ``autocuts_noise`` is not part of the reference tree (parity unpinned);
the default below is an iterative n-sigma clip on baseline, slope and range in the manner of ``qetpy.cut.IterCut``.
A caller-supplied boolean ``cut`` mask still works as before.
"""
import numpy as np

from .plans import PSDPlan


def _torch():
    import torch
    return torch


def two_sided_from_sums(sums, count, nb_samples, fs):
    """psd[k] = sums[|k|] / (count * N * fs) in fftfreq order (numpy or torch input)."""
    half = sums / (float(count) * nb_samples * fs)
    if isinstance(half, np.ndarray):
        return np.concatenate([half, half[1:nb_samples - nb_samples // 2][::-1]])
    torch = _torch()
    return torch.cat([half, torch.flip(half[1:nb_samples - nb_samples // 2], dims=[0])])


def allreduce_sums(sums, count, group=None):
    """Sum the per-rank periodogram sums [N/2+1] and accepted-trace counts over all ranks (NCCL over
    NVLink for CUDA tensors, gloo for the CPU tests).  No-op without an initialised process group."""
    torch = _torch()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(sums, op=torch.distributed.ReduceOp.SUM, group=group)
        torch.distributed.all_reduce(count, op=torch.distributed.ReduceOp.SUM, group=group)
    return sums, count


class NoisePSD:
    """Streaming, multi-GPU ``calc_psd``.

    >>> est = NoisePSD(nb_samples=65536, fs=1.25e6)
    >>> for batch, cut in shard:            # CUDA float64 [n, N], optional bool [n]
    ...     est.update(batch, cut)
    >>> freqs, psd = est.finalize()          # all-reduce across ranks, two-sided, A^2/Hz
    """

    def __init__(self, nb_samples, fs, precision='f64', device=None, typical_rms=None):
        self.nb_samples = int(nb_samples)
        self.fs = float(fs)
        self.plan = PSDPlan(nb_samples, fs, precision=precision, device=device)
        if typical_rms is not None:
            self.plan.set_scale(typical_rms)
        self._median_sum = None
        self._n_median = 0

    def update(self, traces, cut=None, with_offset=False):
        """Accumulate one batch.  ``with_offset`` also accumulates the per-trace medians that
        ``Noise.calc_psd`` averages into the channel offset (noise.py:349)."""
        self.plan.accumulate(traces, cut)
        if with_offset:
            torch = _torch()
            sel = traces if cut is None else traces[cut.to(torch.bool)]
            if sel.shape[0]:
                # np.median of an even-length trace is the mean of the two middle samples
                n = sel.shape[1]
                srt = torch.sort(sel, dim=1).values
                med = srt[:, n // 2] if n % 2 else 0.5 * (srt[:, n // 2 - 1] + srt[:, n // 2])
                s = med.sum().reshape(1)
                self._median_sum = s if self._median_sum is None else self._median_sum + s
                self._n_median += sel.shape[0]

    def local_sums(self):
        return self.plan.sums()

    def finalize(self, group=None):
        """All-reduce the per-rank sums/counts and return (freqs, psd) as numpy float64 [N]."""
        torch = _torch()
        sums, count = self.plan.sums()
        off = self._median_sum
        sums, count = allreduce_sums(sums, count, group)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            if off is not None:
                pack = torch.cat([off, torch.tensor([float(self._n_median)], dtype=off.dtype, device=off.device)])
                torch.distributed.all_reduce(pack, op=torch.distributed.ReduceOp.SUM, group=group)
                off, self._n_median = pack[:1], float(pack[1].item())
        n = int(count.item())
        if n == 0:
            raise ValueError('ERROR: No events selected after noise autocut! Unable to calculate PSD')
        psd = two_sided_from_sums(sums, n, self.nb_samples, self.fs).cpu().numpy()
        freqs = np.fft.fftfreq(self.nb_samples, 1.0 / self.fs)
        self.count = n
        self.offset = None if off is None else float(off.item()) / self._n_median
        return freqs, psd


# ------------------------------------------------------------------------------------------ global autocut (two pass)
def trace_statistics(traces, fs):
    """Per-trace statistics of a CUDA batch [n, N] (float64 amps or int16 ADC counts are not distinguished here: pass
    amps): baseline of the first and last eighth, their difference (slope) and the peak-to-peak range.  Computed with
    the window-reduction kernels (numpy's pairwise sums, bit exact), so a trace gives the same numbers on any rank.
    Returns a CUDA float64 tensor [n, 3] = (baseline, slope, range)."""
    torch = _torch()
    from .plans import ReducePlan
    n = int(traces.shape[-1])
    key = (n, float(fs), traces.device.index)
    plan = _stat_plans.get(key)
    if plan is None:
        plan = ReducePlan(n, fs, 1)
        plan.add(0, 'baseline', 0, n // 8)
        plan.add(0, 'baseline', n - n // 8, n)
        plan.add(0, 'maximum', 0, n - 1)
        plan.add(0, 'minimum', 0, n - 1)
        plan.finalize(traces.device)
        _stat_plans[key] = plan
    o = plan.run(traces.to(torch.float64))
    c = [plan.column(i) for i in range(4)]
    return torch.stack([o[:, c[0]], o[:, c[1]] - o[:, c[0]], o[:, c[2]] - o[:, c[3]]], dim=1)


_stat_plans = {}


def sigma_clip_cut(stats, nsig=2.0, max_iter=20):
    """Iterative n-sigma clipping on every column of ``stats`` [n, k] (numpy): a trace survives when all of its statistics
    stay within nsig standard deviations of the surviving population's mean; repeated until the selection is stable
    (the scheme of qetpy.cut.IterCut, which autocuts_noise is built from).  Deterministic function of the population."""
    stats = np.asarray(stats, dtype=np.float64)
    keep = np.all(np.isfinite(stats), axis=1)
    for _ in range(max_iter):
        if keep.sum() < 2:
            break
        mu, sd = stats[keep].mean(axis=0), stats[keep].std(axis=0)
        new = keep & np.all(np.abs(stats - mu) <= nsig * np.where(sd > 0, sd, np.inf), axis=1)
        if np.array_equal(new, keep):
            break
        keep = new
    return keep


def global_autocut(stats_local, cut_fn=sigma_clip_cut, group=None):
    """The population-wide cut for a rank-sharded sample: ``stats_local`` [n_local, k] (torch tensor, any device) of the
    rank's traces in shard order -> boolean numpy mask [n_local].  All ranks all_gather the statistics (padded to the
    longest shard), evaluate ``cut_fn`` on the concatenation in rank order -- i.e. on exactly the array a single process
    would have -- and keep their own slice: the sharded PSD equals the single-process one."""
    torch = _torch()
    import torch.distributed as dist
    st = stats_local.detach()
    if not (dist.is_available() and dist.is_initialized()):
        return cut_fn(st.cpu().numpy())
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = st.device if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    n_local = torch.tensor([st.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(x.item()) for x in sizes]
    nmax, k = max(sizes + [1]), st.shape[1]
    pad = torch.zeros((nmax, k), dtype=torch.float64, device=dev)
    pad[:st.shape[0]] = st.to(dev)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    population = np.concatenate([p[:m].cpu().numpy() for p, m in zip(parts, sizes)], axis=0)
    mask = cut_fn(population)
    lo = sum(sizes[:rank])
    return mask[lo:lo + sizes[rank]]


def calc_psd(traces, fs, cut=None, precision='f64', batch=4096):
    """One-call form on a CUDA tensor [n, N] (what ``qp.calc_psd(traces[cut], fs, folded_over=False)`` returns)."""
    est = NoisePSD(traces.shape[-1], fs, precision=precision, device=traces.device)
    for i in range(0, traces.shape[0], batch):
        est.update(traces[i:i + batch], None if cut is None else cut[i:i + batch])
    return est.finalize()


def csd_from_sums(sums, count, n_chan, nb_samples, fs):
    """[n*n, N/2+1] component sums (CSDPlan.sums layout) -> two-sided complex csd [n, n, N] in fftfreq order:
    csd[a, b, k] = mean_events X_a[k] conj(X_b[k]) / (N fs); csd[b, a] = conj(csd[a, b]); csd[.., N-k] = conj(csd[.., k])."""
    sums = np.asarray(sums, dtype=np.float64) / (float(count) * nb_samples * fs)
    nh = nb_samples // 2 + 1
    half = np.zeros((n_chan, n_chan, nh), dtype=np.complex128)
    pi = 0
    for a in range(n_chan):
        half[a, a] = sums[a]
        for b in range(a + 1, n_chan):
            half[a, b] = sums[n_chan + 2 * pi] + 1j * sums[n_chan + 2 * pi + 1]
            half[b, a] = np.conj(half[a, b])
            pi += 1
    neg = np.conj(half[:, :, 1:nb_samples - nb_samples // 2][:, :, ::-1])
    return np.concatenate([half, neg], axis=-1)


class NoiseCSD:
    """Streaming, multi-GPU ``calc_csd`` (reference Noise.calc_csd, core/noise.py:374-470: qp.calc_csd(traces[cut], fs,
    folded_over=False) on [n_events, n_chan, N] randoms).  Same flow as ``NoisePSD``: per-rank sums, one all-reduce
    (NCCL over NVLink) of the [n*n, N/2+1] float64 sums and the count, CSD formed on every rank.

    Convention: csd[a, b] = E[X_a conj(X_b)] -- the noise covariance the NxM filter inverts (oracle/ofnxm.py);
    scipy.signal.csd, which QETpy's calc_csd is built on, returns the complex conjugate (= the transpose)."""

    def __init__(self, nb_samples, fs, n_chan, precision='f64', device=None, typical_rms=None):
        from .plans import CSDPlan
        self.nb_samples, self.fs, self.n_chan = int(nb_samples), float(fs), int(n_chan)
        self.plan = CSDPlan(nb_samples, fs, n_chan, precision=precision, device=device)
        if typical_rms is not None:
            self.plan.set_scale(typical_rms)

    def update(self, traces, cut=None):
        self.plan.accumulate(traces, cut)

    def finalize(self, group=None):
        sums, count = self.plan.sums()
        sums, count = allreduce_sums(sums, count, group)
        n = int(count.item())
        if n == 0:
            raise ValueError('ERROR: No events selected after pileup cut!')
        self.count = n
        csd = csd_from_sums(sums.cpu().numpy(), n, self.n_chan, self.nb_samples, self.fs)
        return np.fft.fftfreq(self.nb_samples, 1.0 / self.fs), csd


def calc_csd(traces, fs, cut=None, precision='f64', batch=2048):
    """One-call form on a CUDA tensor [n_events, n_chan, N]."""
    est = NoiseCSD(traces.shape[-1], fs, traces.shape[1], precision=precision, device=traces.device)
    for i in range(0, traces.shape[0], batch):
        est.update(traces[i:i + batch], None if cut is None else cut[i:i + batch])
    return est.finalize()


def calc_psd_from_reader(reader, channel, cut=None, precision='f64', batch=4096, device=None, autocut=False, cut_fn=sigma_clip_cut):
    """``Noise.calc_psd`` on the randoms behind a ``detprocess_b200.io.EventReader`` (reference core/noise.py:216-370:
    the events are read in one shot there, in batches here; ranks take contiguous shards and all-reduce the sums).
    ADC counts become amps on the device.  Returns (freqs, psd two-sided)."""
    torch = _torch()
    from ..process.features import dist_info, shard_range
    ci = reader.channels.index(channel)
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    est = NoisePSD(int(reader.metadata['nb_samples']), reader.sample_rate, precision=precision, device=dev)
    rank, world = dist_info()
    lo, hi = shard_range(len(reader), rank, world)
    if autocut:
        # pass 1: per-trace statistics of this rank's shard -> one all_gather -> the global mask (reference noise.py:331:
        # qp.autocuts_noise on the whole population)
        stats = []
        for i0 in range(lo, hi, batch):
            i1 = min(i0 + batch, hi)
            stats.append(trace_statistics(reader.to_amps(reader.upload(i0, i1, dev))[:, ci].contiguous(), reader.sample_rate))
        st = torch.cat(stats) if stats else torch.zeros((0, 3), dtype=torch.float64, device=dev)
        mask = global_autocut(st, cut_fn)
        full = np.zeros(len(reader), dtype=bool)
        full[lo:hi] = mask
        cut = full if cut is None else (np.asarray(cut, dtype=bool) & full)
    for i0 in range(lo, hi, batch):
        i1 = min(i0 + batch, hi)
        x = reader.to_amps(reader.upload(i0, i1, dev))[:, ci].contiguous()
        est.update(x, None if cut is None else torch.as_tensor(np.asarray(cut[i0:i1]), device=dev))
    return est.finalize()


def calc_csd_from_reader(reader, channels, cut=None, precision='f64', batch=2048, device=None):
    """``Noise.calc_csd`` (core/noise.py:374-470) on the randoms behind an ``EventReader``; ``channels`` is the list (or
    'a|b' string) of joint channels in the order of the csd's rows.  Returns (freqs, csd [n, n, N])."""
    torch = _torch()
    from ..process.features import dist_info, shard_range
    if isinstance(channels, str):
        channels = [c.strip() for c in channels.split('|')]
    if len(channels) < 2:
        raise ValueError('ERROR: At least 2 channels required to calculate csd')
    idx = [reader.channels.index(c) for c in channels]
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    est = NoiseCSD(int(reader.metadata['nb_samples']), reader.sample_rate, len(idx), precision=precision, device=dev)
    rank, world = dist_info()
    lo, hi = shard_range(len(reader), rank, world)
    for i0 in range(lo, hi, batch):
        i1 = min(i0 + batch, hi)
        x = reader.to_amps(reader.upload(i0, i1, dev))[:, idx].contiguous()
        est.update(x, None if cut is None else torch.as_tensor(np.asarray(cut[i0:i1]), device=dev))
    return est.finalize()
