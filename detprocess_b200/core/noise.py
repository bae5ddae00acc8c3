"""
Noise PSD estimation on the GPU -- the ``Noise.calc_psd`` hot path of the reference
(detprocess/core/noise.py:216-370), config C5 of BASELINE.json.

The reference holds every trace in RAM and calls ``qp.calc_psd(traces[cut], fs,
folded_over=False)`` (noise.py:344).  Here traces stream through ``PSDPlan.accumulate`` in
batches (each trace is read from HBM exactly once); every rank accumulates
``sum |fft(x)_k|^2`` for its shard, the ``[N/2+1]`` sums and the accepted-trace counts are
all-reduced (NCCL over NVLink when ``torch.distributed`` is initialised with the nccl
backend) and the two-sided PSD is formed on every rank.

The noise autocut (``qp.autocuts_noise``, noise.py:331) is QETpy code that is not part of
the reference tree; its result enters here as the boolean ``cut`` mask.
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


def calc_psd_from_reader(reader, channel, cut=None, precision='f64', batch=4096, device=None):
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
