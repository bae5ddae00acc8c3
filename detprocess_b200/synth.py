"""
Synthetic TES-like inputs (template, two-sided PSD, coloured-noise traces with
injected pulses) of the shapes BASELINE.json names.  numpy only; used by the
tests, ``bench.py`` and ``__graft_entry__.smoke()``.  There is no network and no
HDF5 library in this environment, so every measurement runs on these arrays
(SURVEY.md 8(d)).
"""

import numpy as np

__all__ = ['make_template', 'make_psd', 'make_noise', 'make_traces', 'make_continuous', 'SynthSetup']

FS_DEFAULT = 1.25e6
SEED_DEFAULT = 12345


def make_template(nb_samples, fs=FS_DEFAULT, nb_pretrigger=None,
                  tau_rise=20e-6, tau_fall=200e-6):
    """Two-pole pulse, peak normalised to 1, baseline 0, onset at nb_pretrigger."""
    if nb_pretrigger is None:
        nb_pretrigger = nb_samples // 2
    t = (np.arange(nb_samples) - nb_pretrigger) / fs
    tp = np.where(t > 0, t, 0.0)
    pulse = np.where(t > 0, np.exp(-tp / tau_fall) - np.exp(-tp / tau_rise), 0.0)
    return pulse / pulse.max()


def make_glitch_template(nb_samples, fs=FS_DEFAULT, nb_pretrigger=None):
    return make_template(nb_samples, fs, nb_pretrigger, tau_rise=2e-6, tau_fall=10e-6)


def make_psd(nb_samples, fs=FS_DEFAULT, sigma=1e-11, f_corner=1e3):
    """Two-sided PSD [N] in A^2/Hz (fftfreq order): white + 1/f, even in f, > 0."""
    f = np.abs(np.fft.fftfreq(nb_samples, d=1.0 / fs))
    df = fs / nb_samples
    return sigma ** 2 * (1.0 + f_corner / np.maximum(f, df))


def make_noise(nb_events, psd, fs=FS_DEFAULT, rng=None):
    """Gaussian noise traces [B, N] whose two-sided PSD expectation is ``psd``."""
    rng = np.random.default_rng(SEED_DEFAULT) if rng is None else rng
    n = psd.shape[-1]
    nh = n // 2 + 1
    # E|fft(x)|^2 = psd * N * fs  (calc_psd convention: |fft|^2/(N fs))
    amp = np.sqrt(psd[:nh] * n * fs / 2.0)
    spec = (rng.standard_normal((nb_events, nh)) + 1j * rng.standard_normal((nb_events, nh))) * amp
    spec[:, 0] = spec[:, 0].real * np.sqrt(2.0)
    if n % 2 == 0:
        spec[:, -1] = spec[:, -1].real * np.sqrt(2.0)
    return np.fft.irfft(spec, n=n, axis=-1)


def make_traces(nb_events, template, psd, fs=FS_DEFAULT, rng=None,
                amp_max=2e-7, max_delay=300, pulse_fraction=0.9, offset=0.0,
                return_truth=False):
    """
    Noise + template*A rolled by an integer delay d on ``pulse_fraction`` of the
    events (A ~ U(0, amp_max), d ~ U{-max_delay..max_delay}); the rest noise only.
    """
    rng = np.random.default_rng(SEED_DEFAULT) if rng is None else rng
    traces = make_noise(nb_events, psd, fs, rng)
    has = rng.random(nb_events) < pulse_fraction
    amps = np.where(has, rng.random(nb_events) * amp_max, 0.0)
    delays = np.where(has, rng.integers(-max_delay, max_delay + 1, nb_events), 0)
    for i in np.nonzero(has)[0]:
        traces[i] += amps[i] * np.roll(template, int(delays[i]))
    if offset:
        traces += offset
    if return_truth:
        return traces, amps, delays
    return traces


def make_continuous(n_samples, template, psd, fs=FS_DEFAULT, rng=None, pulse_rate_hz=5.0, amp_range=(5e-8, 2e-7),
                    nb_pretrigger=None, offset=0.0, return_truth=False):
    """Continuous stream [n_samples]: coloured noise (drawn block-wise from ``psd``, blocks cross-faded so
    the stream has no seams) + Poisson pulses of the template's shape (BASELINE.json config C4)."""
    rng = np.random.default_rng(SEED_DEFAULT) if rng is None else rng
    n = psd.shape[-1]
    if nb_pretrigger is None:
        nb_pretrigger = n // 2
    nblk = n_samples // (n // 2) + 2
    blocks = make_noise(nblk, psd, fs, rng)
    w = np.sqrt(0.5 * (1 - np.cos(2 * np.pi * (np.arange(n) + 0.5) / n)))   # sqrt-Hann: w[i]^2 + w[i+n/2]^2 = 1
    x = np.zeros(n_samples + 2 * n)
    for b in range(nblk):
        x[b * (n // 2): b * (n // 2) + n] += w * blocks[b]
    x = x[n // 2: n // 2 + n_samples].copy()
    n_pulses = rng.poisson(pulse_rate_hz * n_samples / fs)
    t0 = np.sort(rng.integers(2 * n, max(2 * n + 1, n_samples - 2 * n), n_pulses))
    amps = rng.uniform(amp_range[0], amp_range[1], n_pulses)
    shape = template[nb_pretrigger:]
    for t, a in zip(t0, amps):
        m = min(len(shape), n_samples - t)
        x[t:t + m] += a * shape[:m]
    if offset:
        x += offset
    if return_truth:
        return x, t0, amps
    return x


class SynthSetup:
    """Template(s) + PSD + trace factory for one (N, pretrigger) shape."""

    def __init__(self, nb_samples=32768, fs=FS_DEFAULT, nb_pretrigger=None, seed=SEED_DEFAULT):
        self.nb_samples = int(nb_samples)
        self.fs = float(fs)
        self.nb_pretrigger = self.nb_samples // 2 if nb_pretrigger is None else int(nb_pretrigger)
        self.seed = seed
        self.template = make_template(self.nb_samples, fs, self.nb_pretrigger)
        self.template_glitch = make_glitch_template(self.nb_samples, fs, self.nb_pretrigger)
        self.psd = make_psd(self.nb_samples, fs)

    def traces(self, nb_events, rank=0, **kwargs):
        rng = np.random.default_rng(self.seed + rank)
        return make_traces(nb_events, self.template, self.psd, self.fs, rng, **kwargs)


class SynthNxM:
    """n-channel, m-template setup for the NxM optimal filter: templates [n, m, N], a cross-spectral density [n, n, N]
    (independent channel noise + a common-mode source that reaches channel a with gain c_a and a delay of a samples,
    so the off-diagonal terms are complex), and correlated-noise traces [B, n, N] with both templates injected at one
    common delay."""

    def __init__(self, nb_samples=32768, n_chan=2, n_templ=2, fs=FS_DEFAULT, nb_pretrigger=None):
        N = self.nb_samples = int(nb_samples)
        self.fs = float(fs)
        self.n_chan, self.n_templ = int(n_chan), int(n_templ)
        self.nb_pretrigger = N // 2 if nb_pretrigger is None else int(nb_pretrigger)
        shapes = [make_template(N, fs, self.nb_pretrigger),
                  make_glitch_template(N, fs, self.nb_pretrigger),
                  make_template(N, fs, self.nb_pretrigger, tau_rise=5e-6, tau_fall=60e-6)]
        self.templates = np.zeros((self.n_chan, self.n_templ, N))
        for a in range(self.n_chan):
            for i in range(self.n_templ):
                share = 1.0 if (a % self.n_templ) == i else 0.25 / (1 + abs(a - i))
                self.templates[a, i] = share * shapes[i % 3]
        self.psd_chan = [make_psd(N, fs, sigma=1e-11 * (1.0 + 0.3 * a)) for a in range(self.n_chan)]
        self.psd_common = make_psd(N, fs, sigma=0.7e-11, f_corner=3e3)
        self.gain_common = np.array([1.0, -0.6, 0.8, 0.5][:self.n_chan])
        f = np.fft.fftfreq(N, d=1.0 / fs)
        self.delay_phase = [np.exp(-2j * np.pi * f * (a / fs)) for a in range(self.n_chan)]   # delay of a samples
        self.csd = np.zeros((self.n_chan, self.n_chan, N), dtype=np.complex128)
        for a in range(self.n_chan):
            for b in range(self.n_chan):
                self.csd[a, b] = (self.gain_common[a] * self.gain_common[b] * self.psd_common
                                  * self.delay_phase[a] * np.conj(self.delay_phase[b]))
            self.csd[a, a] += self.psd_chan[a]

    def traces(self, nb_events, rng=None, amp_max=2e-7, max_delay=300, pulse_fraction=0.9, return_truth=False):
        rng = np.random.default_rng(SEED_DEFAULT) if rng is None else rng
        N = self.nb_samples
        common = make_noise(nb_events, self.psd_common, self.fs, rng)
        x = np.zeros((nb_events, self.n_chan, N))
        for a in range(self.n_chan):
            x[:, a] = make_noise(nb_events, self.psd_chan[a], self.fs, rng) + self.gain_common[a] * np.roll(common, a, axis=-1)
        has = rng.random(nb_events) < pulse_fraction
        amps = np.where(has[:, None], rng.random((nb_events, self.n_templ)) * amp_max, 0.0)
        delays = np.where(has, rng.integers(-max_delay, max_delay + 1, nb_events), 0)
        for e in np.nonzero(has)[0]:
            for i in range(self.n_templ):
                x[e] += amps[e, i] * np.roll(self.templates[:, i], int(delays[e]), axis=-1)
        if return_truth:
            return x, amps, delays
        return x
