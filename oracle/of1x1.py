"""
CPU oracle: OF1x1 optimal-filter maths (float64 numpy).  TEST INFRASTRUCTURE ONLY.

What this restates
------------------
The reference computes OF1x1 features by driving two QETpy classes:

* ``qp.OFBase``  -- built once per (nb_samples, nb_pretrigger, tag) key in
  ``detprocess/process/processing_data.py:275-381`` (``OFBase(fs)``, ``set_csd``,
  ``add_template``, ``calc_phi``) and refreshed per event in
  ``processing_data.py:712-772`` (``clear_signal``, ``update_signal(calc_fft=True)``,
  ``calc_signal_filt``, ``calc_signal_filt_td``).
* ``qp.OF1x1``   -- built per event per algorithm in
  ``detprocess/core/algorithms.py:331-341`` (nodelay), ``:410-421`` (unconstrained),
  ``:533-558`` (constrained + chi2 nopulse + resolutions).

QETpy (``qetpy>=1.8.6``, reference ``setup.py:73``) is NOT vendored in the reference
and NOT installable in this environment, so the arithmetic below restates QETpy's
published optimum-filter algorithm (``qetpy/core/_of_base.py``, ``_of_1x1.py``; the
same maths as the older ``qetpy.OptimumFilter``):

    df      = fs / N
    J       = two-sided PSD (A^2/Hz), fftfreq order; coupling 'AC' => J[0] = inf
    s       = fft(template) / N / df            (integralnorm: s /= s[0])
    phi     = conj(s) / J
    norm    = Re(sum(phi * s)) * df
    v       = fft(trace) / N / df
    filt    = phi * v / norm
    amps_td = Re(ifft(filt * N)) * df           (index 0 == zero delay)
    chi0    = Re(sum(conj(v) * v / J)) * df
    chi2_td = chi0 - amps_td**2 * norm
    rolled by ``pretrigger_samples`` so that zero delay sits at index pretrigger
    nodelay : values at rolled index ``pretrigger``
    delay   : argmin of rolled chi2 inside the window, t0 = (ind - pretrigger)/fs
    lowchi2 = sum_{|f|<=fcut} df * |v - amp*exp(-2j*pi*t0*f)*s|^2 / J
    ampres  = 1/sqrt(norm)
    timeres = 1/sqrt(amp^2 * sum((2 pi f)^2 |s|^2 / J) * df)

PARITY UNPINNED: the reference ships no tests/golden vectors for these numbers
(SURVEY.md F2/F3).  amp and t0 are independent of the FFT normalisation; chi2,
lowchi2, ampres and timeres depend on it; the convention used here (``/N/df``) is
isolated in ``_fft_norm``.  ``OF_WINDOW_MAX_INCLUSIVE`` isolates the one other
convention that could not be confirmed (whether the constrained-fit window
includes ``window_max_index``).
"""

import numpy as np

__all__ = ['OFBaseOracle', 'OF1x1Oracle', 'of1x1_batch', 'OF_WINDOW_MAX_INCLUSIVE', 'interpolate_parabola']

# Whether the delay-search window [window_min_index, window_max_index] includes its
# upper end.  QETpy builds the candidate set with a python slice (end exclusive),
# like every other window in detprocess (algorithms.py:698); kept as ONE constant.
OF_WINDOW_MAX_INCLUSIVE = False

SENTINEL = -999999.0


def interpolate_parabola(v_prev, v_best, v_next, delta, t_interp=None):
    """The three-point parabola of QETpy's ``interpolate_t0`` option (``_interpolate_parabola`` of the optimum-filter
    classes, as recalled -- QETpy is not in the reference tree: PARITY UNPINNED, the whole convention lives in this one
    function): values at the best delay and one sample before / after it, spacing ``delta``; returns the vertex time
    relative to the best sample and the parabola's value there (or at a given ``t_interp``).  ``calc`` applies it to the
    chi2 series first (-> t_interp, chi2) and then evaluates the amplitude parabola at that same t_interp."""
    sf = 1.0 / (v_best * 100.0)                    # scale factor QETpy applies "for precision purposes"
    a = sf * (v_next - 2.0 * v_best + v_prev) / (2.0 * delta ** 2)
    b = sf * (v_next - v_prev) / (2.0 * delta)
    c = sf * v_best
    if t_interp is None:
        t_interp = -b / (2.0 * a)
    return t_interp, (a * t_interp ** 2 + b * t_interp + c) / sf


def _fft_norm(x, fs):
    """fft(x)/N/df -- the single place the FFT normalisation convention lives."""
    n = x.shape[-1]
    df = fs / n
    return np.fft.fft(x, axis=-1) / n / df


class OFBaseOracle:
    """
    Minimal stand-in for ``qp.OFBase`` exposing the calls detprocess makes
    (SURVEY.md 8(b) "of_base object surface"; processing_data.py:278-381,731-772).
    Single-channel (1x1) only.
    """

    def __init__(self, sample_rate, verbose=False):
        self._fs = float(sample_rate)
        self._verbose = verbose
        self._nbins = None
        self._psd = {}            # chan -> J [N] (coupling applied)
        self._templates = {}      # chan -> tag -> template [N]
        self._templates_fft = {}  # chan -> tag -> s [N]
        self._pretrigger = {}     # chan -> tag -> int
        self._phis = {}           # chan -> tag -> phi [N]
        self._norms = {}          # chan -> tag -> float
        self._signals = {}        # chan -> trace [N]
        self._signals_fft = {}    # chan -> v [N]
        self._signals_filt = {}   # chan -> tag -> filt [N]
        self._signals_filt_td = {}  # chan -> tag -> amps_td [N]

    # ---- setup ---------------------------------------------------------
    def sample_rate(self):
        return self._fs

    def nb_samples(self):
        return self._nbins

    def fft_freqs(self):
        return np.fft.fftfreq(self._nbins, d=1.0 / self._fs)

    def df(self):
        return self._fs / self._nbins

    def _check_n(self, n):
        if self._nbins is None:
            self._nbins = int(n)
        elif self._nbins != int(n):
            raise ValueError('ERROR: inconsistent number of samples')

    def set_csd(self, channel, csd, coupling='AC',
                ignored_frequency_peaks=None, ignore_harmonics=False):
        """csd: two-sided PSD [N] (or [1,1,N]) in fftfreq order."""
        J = np.array(np.real(np.asarray(csd)), dtype=np.float64).reshape(-1).copy()
        self._check_n(J.shape[-1])
        if coupling == 'AC':
            J[0] = np.inf
        if ignored_frequency_peaks is not None:
            f = self.fft_freqs()
            peaks = np.atleast_1d(np.asarray(ignored_frequency_peaks, dtype=float))
            df = self.df()
            for pk in peaks:
                mults = [pk]
                if ignore_harmonics:
                    mults = np.arange(pk, self._fs / 2, pk)
                for fpk in mults:
                    J[np.abs(np.abs(f) - fpk) <= df / 2] = np.inf
        self._psd[channel] = J

    set_psd = set_csd

    def csd(self, channel):
        return self._psd.get(channel)

    psd = csd

    def add_template(self, channel, template, template_tag='default',
                     pretrigger_samples=None, integralnorm=False, overwrite=False):
        template = np.asarray(template, dtype=np.float64).reshape(-1)
        self._check_n(template.shape[-1])
        if (not overwrite and channel in self._templates
                and template_tag in self._templates[channel]):
            raise ValueError('ERROR: template already exists, use overwrite=True')
        s = _fft_norm(template, self._fs)
        if integralnorm:
            s = s / s[0]
        if pretrigger_samples is None:
            pretrigger_samples = self._nbins // 2
        self._templates.setdefault(channel, {})[template_tag] = template
        self._templates_fft.setdefault(channel, {})[template_tag] = s
        self._pretrigger.setdefault(channel, {})[template_tag] = int(pretrigger_samples)
        # invalidate derived
        self._phis.get(channel, {}).pop(template_tag, None)
        self._norms.get(channel, {}).pop(template_tag, None)

    def template(self, channel, template_tag='default'):
        return self._templates.get(channel, {}).get(template_tag)

    def template_fft(self, channel, template_tag='default'):
        return self._templates_fft.get(channel, {}).get(template_tag)

    def template_tags(self, channel):
        return list(self._templates.get(channel, {}).keys())

    def pretrigger_samples(self, channel, template_tag='default'):
        return self._pretrigger[channel][template_tag]

    def calc_phi(self, channel, template_tag='default'):
        s = self._templates_fft[channel][template_tag]
        J = self._psd[channel]
        phi = np.conj(s) / J
        self._phis.setdefault(channel, {})[template_tag] = phi
        self._norms.setdefault(channel, {})[template_tag] = float(
            np.real(np.dot(phi, s)) * self.df())

    def phi(self, channel, template_tag='default'):
        return self._phis.get(channel, {}).get(template_tag)

    def norm(self, channel, template_tag='default'):
        return self._norms.get(channel, {}).get(template_tag)

    # ---- per event -------------------------------------------------------
    def clear_signal(self):
        self._signals = {}
        self._signals_fft = {}
        self._signals_filt = {}
        self._signals_filt_td = {}

    def is_signal_stored(self, channel):
        return channel in self._signals

    def update_signal(self, channel, signal, calc_fft=True, **kwargs):
        signal = np.asarray(signal, dtype=np.float64).reshape(-1)
        if signal.shape[-1] != self._nbins:
            raise ValueError('ERROR: signal length != template/psd length')
        self._signals[channel] = signal
        self._signals_filt.pop(channel, None)
        self._signals_filt_td.pop(channel, None)
        if calc_fft:
            self._signals_fft[channel] = _fft_norm(signal, self._fs)

    def signal(self, channel):
        return self._signals.get(channel)

    def signal_fft(self, channel, **kwargs):
        return self._signals_fft.get(channel)

    def calc_signal_filt(self, channel, template_tag=None):
        tags = [template_tag] if template_tag is not None else self.template_tags(channel)
        v = self._signals_fft[channel]
        for tag in tags:
            if self.phi(channel, tag) is None:
                self.calc_phi(channel, tag)
            self._signals_filt.setdefault(channel, {})[tag] = (
                self._phis[channel][tag] * v / self._norms[channel][tag])

    def calc_signal_filt_td(self, channel, template_tag=None):
        tags = [template_tag] if template_tag is not None else self.template_tags(channel)
        for tag in tags:
            if tag not in self._signals_filt.get(channel, {}):
                self.calc_signal_filt(channel, tag)
            filt = self._signals_filt[channel][tag]
            self._signals_filt_td.setdefault(channel, {})[tag] = (
                np.real(np.fft.ifft(filt * self._nbins)) * self.df())

    def signal_filt(self, channel, template_tag='default'):
        return self._signals_filt.get(channel, {}).get(template_tag)

    def signal_filt_td(self, channel, template_tag='default'):
        return self._signals_filt_td.get(channel, {}).get(template_tag)

    def calc_chisq0(self, channel):
        v = self._signals_fft[channel]
        J = self._psd[channel]
        return float(np.real(np.dot(np.conj(v) / J, v)) * self.df())


class OF1x1Oracle:
    """
    Stand-in for ``qp.OF1x1(of_base=, channel=, template_tag=)`` as used by
    ``detprocess/core/algorithms.py:331-341, 410-421, 533-558``.
    """

    def __init__(self, of_base, channel, template_tag='default'):
        self._of_base = of_base
        self._channel = channel
        self._tag = template_tag
        self._nodelay = None
        self._withdelay = None
        self._chi0 = None

    def _arrays(self):
        ofb, ch, tag = self._of_base, self._channel, self._tag
        if ofb.signal_filt_td(ch, tag) is None:
            ofb.calc_signal_filt_td(ch, tag)
        amps_td = ofb.signal_filt_td(ch, tag)
        norm = ofb.norm(ch, tag)
        if self._chi0 is None:
            self._chi0 = ofb.calc_chisq0(ch)
        chi2_td = self._chi0 - amps_td ** 2 * norm
        pre = ofb.pretrigger_samples(ch, tag)
        return np.roll(amps_td, pre), np.roll(chi2_td, pre), pre

    def get_chisq_lowfreq(self, amp, t0, lowchi2_fcutoff=10000):
        ofb, ch, tag = self._of_base, self._channel, self._tag
        f = ofb.fft_freqs()
        v = ofb.signal_fft(ch)
        s = ofb.template_fft(ch, tag)
        J = ofb.psd(ch)
        chi2tot = ofb.df() * np.abs(v - amp * np.exp(-2.0j * np.pi * t0 * f) * s) ** 2 / J
        return float(np.sum(chi2tot[np.abs(f) <= lowchi2_fcutoff]))

    def calc(self, window_min_from_trig_usec=None, window_max_from_trig_usec=None,
             window_min_index=None, window_max_index=None,
             lowchi2_fcutoff=10000, interpolate_t0=False,
             lgc_outside_window=False, pulse_direction_constraint=0,
             lgc_fit_withdelay=True, lgc_fit_nodelay=True, lgc_plot=False, **kwargs):
        ofb, ch, tag = self._of_base, self._channel, self._tag
        amps, chi2, pre = self._arrays()
        n = amps.shape[-1]
        fs = ofb.sample_rate()

        if lgc_fit_nodelay:
            a0, c0 = float(amps[pre]), float(chi2[pre])
            self._nodelay = (a0, 0.0, c0,
                             self.get_chisq_lowfreq(a0, 0.0, lowchi2_fcutoff))

        if lgc_fit_withdelay:
            wmin, wmax = None, None
            # usec form wins over index form
            if window_min_from_trig_usec is not None:
                wmin = int(np.floor(pre + window_min_from_trig_usec * fs * 1e-6))
            elif window_min_index is not None:
                wmin = int(window_min_index)
            if window_max_from_trig_usec is not None:
                wmax = int(np.ceil(pre + window_max_from_trig_usec * fs * 1e-6))
            elif window_max_index is not None:
                wmax = int(window_max_index)
            lo, hi = of_window_bounds(n, wmin, wmax)
            mask = np.zeros(n, dtype=bool)
            mask[lo:hi] = True
            if lgc_outside_window:
                mask = ~mask
            if pulse_direction_constraint in (1, -1):
                mask &= (amps * pulse_direction_constraint > 0)
            if not mask.any():
                raise ValueError('ERROR: empty OF delay window')
            masked = np.where(mask, chi2, np.inf)
            ind = int(np.argmin(masked))
            a, c = float(amps[ind]), float(chi2[ind])
            t0 = (ind - pre) / fs
            if interpolate_t0 and 0 < ind < n - 1:
                dt, c = interpolate_parabola(float(chi2[ind - 1]), c, float(chi2[ind + 1]), 1.0 / fs)
                _, a = interpolate_parabola(float(amps[ind - 1]), a, float(amps[ind + 1]), 1.0 / fs, t_interp=dt)
                t0 = t0 + dt
            self._withdelay = (a, t0, c, self.get_chisq_lowfreq(a, t0, lowchi2_fcutoff))

    def get_result_nodelay(self):
        return self._nodelay

    def get_result_withdelay(self):
        return self._withdelay

    def get_chisq_nopulse(self):
        if self._chi0 is None:
            self._chi0 = self._of_base.calc_chisq0(self._channel)
        return self._chi0

    def get_energy_resolution(self):
        return 1.0 / np.sqrt(self._of_base.norm(self._channel, self._tag))

    def get_time_resolution(self, amp=None):
        ofb, ch, tag = self._of_base, self._channel, self._tag
        if amp is None:
            amp = self._withdelay[0]
        f = ofb.fft_freqs()
        s = ofb.template_fft(ch, tag)
        J = ofb.psd(ch)
        return float(1.0 / np.sqrt(
            amp ** 2 * np.sum((2 * np.pi * f) ** 2 * np.abs(s) ** 2 / J) * ofb.df()))


def of_window_bounds(n, wmin, wmax):
    """Half-open [lo, hi) candidate range of rolled delay indices."""
    lo = 0 if wmin is None else int(wmin)
    if wmax is None:
        hi = n
    else:
        hi = int(wmax) + (1 if OF_WINDOW_MAX_INCLUSIVE else 0)
    lo = min(max(lo, 0), n)
    hi = min(max(hi, 0), n)
    return lo, hi


def of1x1_batch(traces, template, psd, fs, pretrigger_samples, windows=(),
                coupling='AC', integralnorm=False, lowchi2_fcutoff=10000, interpolate=False):
    """
    Vectorised float64 evaluation of the same maths for a [B, N] batch -- used by
    the parity tests so that thousands of events finish in seconds.

    windows : sequence of (window_min_index, window_max_index, lgc_outside) in
              rolled-index form; ``None`` bounds mean unconstrained.
    Returns a dict of arrays:
      chi0[B]; amp0, chi2_0, lowchi2_0 [B] (no-delay fit);
      amp[W,B], ind[W,B] (rolled index), t0[W,B], chi2[W,B], lowchi2[W,B],
      ampres (scalar), timeres[W,B]
    """
    traces = np.atleast_2d(np.asarray(traces, dtype=np.float64))
    nb, n = traces.shape
    df = fs / n
    f = np.fft.fftfreq(n, d=1.0 / fs)
    J = np.array(psd, dtype=np.float64).copy()
    if coupling == 'AC':
        J[0] = np.inf
    s = _fft_norm(np.asarray(template, dtype=np.float64), fs)
    if integralnorm:
        s = s / s[0]
    phi = np.conj(s) / J
    norm = float(np.real(np.dot(phi, s)) * df)
    v = _fft_norm(traces, fs)
    amps_td = np.real(np.fft.ifft(phi * v / norm * n, axis=-1)) * df
    chi0 = np.real(np.sum(np.conj(v) / J * v, axis=-1)) * df
    chi2_td = chi0[:, None] - amps_td ** 2 * norm
    pre = int(pretrigger_samples)
    amps = np.roll(amps_td, pre, axis=-1)
    chi2 = np.roll(chi2_td, pre, axis=-1)
    low = np.abs(f) <= lowchi2_fcutoff
    tsum = float(np.sum((2 * np.pi * f) ** 2 * np.abs(s) ** 2 / J) * df)

    def lowchi2(a, t0):
        ph = np.exp(-2.0j * np.pi * t0[:, None] * f[None, low])
        r = v[:, low] - a[:, None] * ph * s[None, low]
        return np.sum(df * np.abs(r) ** 2 / J[None, low], axis=-1)

    out = {'chi0': chi0, 'norm': norm, 'ampres': 1.0 / np.sqrt(norm)}
    a0 = amps[:, pre].copy()
    out['amp0'] = a0
    out['chi2_0'] = chi2[:, pre].copy()
    out['lowchi2_0'] = lowchi2(a0, np.zeros(nb))
    nw = len(windows)
    for key in ('amp', 't0', 'chi2', 'lowchi2', 'timeres'):
        out[key] = np.zeros((nw, nb))
    out['ind'] = np.zeros((nw, nb), dtype=np.int64)
    rows = np.arange(nb)
    for iw, (wmin, wmax, outside) in enumerate(windows):
        lo, hi = of_window_bounds(n, wmin, wmax)
        mask = np.zeros(n, dtype=bool)
        mask[lo:hi] = True
        if outside:
            mask = ~mask
        ind = np.argmin(np.where(mask[None, :], chi2, np.inf), axis=-1)
        a = amps[rows, ind]
        t0 = (ind - pre) / fs
        c2 = chi2[rows, ind]
        if interpolate:
            inner = (ind > 0) & (ind < n - 1)
            im, ip = np.clip(ind - 1, 0, n - 1), np.clip(ind + 1, 0, n - 1)
            with np.errstate(all='ignore'):
                dt, ci = interpolate_parabola(chi2[rows, im], c2, chi2[rows, ip], 1.0 / fs)
                _, ai = interpolate_parabola(amps[rows, im], a, amps[rows, ip], 1.0 / fs, t_interp=dt)
            a = np.where(inner, ai, a)
            c2 = np.where(inner, ci, c2)
            t0 = np.where(inner, t0 + dt, t0)
        out['ind'][iw] = ind
        out['amp'][iw] = a
        out['t0'][iw] = t0
        out['chi2'][iw] = c2
        out['lowchi2'][iw] = lowchi2(a, t0)
        with np.errstate(divide='ignore'):
            out['timeres'][iw] = 1.0 / np.sqrt(a ** 2 * tsum)

    def at(ind):
        """the same fit evaluated at a GIVEN rolled delay index per event (near-tie bookkeeping of the fp32 parity
        tests: what the float64 maths says at the index another implementation picked)"""
        ind = np.asarray(ind, dtype=np.int64)
        a = amps[rows, ind]
        t0 = (ind - pre) / fs
        with np.errstate(divide='ignore'):
            tr = 1.0 / np.sqrt(a ** 2 * tsum)
        return {'amp': a, 't0': t0, 'chi2': chi2[rows, ind], 'lowchi2': lowchi2(a, t0), 'timeres': tr}

    out['at'] = at
    return out
