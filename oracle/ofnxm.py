"""
CPU oracle: NxM optimal filter (float64 numpy).  TEST INFRASTRUCTURE ONLY -- nothing under
``detprocess_b200/`` may import this module.

What this restates
------------------
``FeatureExtractors.ofnxm`` (reference ``detprocess/core/algorithms.py:141-274``) drives ``qp.OFnxm``:

    OF = qp.OFnxm(of_base=of_base, channels='a|b', template_tag=tag)      # algorithms.py:246
    OF.calc()                                                             # :252
    amps, t0, chi2 = OF.get_fit_withdelay(window_min_index=..., window_max_index=...,
                                          lgc_outside_window=..., interpolate_t0=...)   # :255
    amps0, t00, chi20 = OF.get_fit_nodelay()                              # :263

with ``of_base.template(channel, template_tag)`` of shape ``[n_chan, n_templ, N]`` (:196-206) and the
``[n, n, N]`` cross-spectral density of ``FilterData.get_csd`` (``filterdata.py:380``) set through
``OFBase.set_csd`` (``processing_data.py:294-326``).

QETpy (``qetpy>=1.8.6``) is not vendored in the reference and not installable here, so the arithmetic restates the
published N-channel, M-template optimum filter with one common time delay, in the same normalisation as
``oracle/of1x1.py`` (to which it reduces for n = m = 1):

    S[a,i,k]   = fft(template[a,i]) / N / df
    iS[k]      = inv(csd[:, :, k])              coupling 'AC' => iS[0] = 0
    Phi[i,a,k] = sum_b conj(S[b,i,k]) iS[k][b,a]
    P[i,j]     = Re sum_{a,k} Phi[i,a,k] S[a,j,k] df
    V[a,k]     = fft(trace[a]) / N / df
    q_i(t)     = Re ifft_k( sum_a Phi[i,a,k] V[a,k] * N ) df          (index 0 == zero delay)
    amps(t)    = P^-1 q(t)
    chi0       = Re sum_k V^H iS V df
    chi2(t)    = chi0 - q(t)^T P^-1 q(t)
    rolled by ``pretrigger_samples``; no-delay = rolled index pretrigger; with delay = argmin of the rolled chi2
    inside (or outside) [window_min_index, window_max_index), t0 = (ind - pretrigger) / fs.

PARITY UNPINNED: the reference holds no tests or golden vectors for these numbers; amplitudes and t0
do not depend on the FFT normalisation, chi2 does (the ``/N/df`` convention is the one of ``oracle/of1x1.py``).
``interpolate_t0`` is not restated.
"""

import numpy as np

from .of1x1 import _fft_norm, of_window_bounds

__all__ = ['ofnxm_setup', 'ofnxm_batch']


def ofnxm_setup(templates, csd, fs, pretrigger_samples, coupling='AC', integralnorm=False):
    """templates [n, m, N] float64, csd [n, n, N] complex (two-sided, fftfreq order)."""
    templates = np.asarray(templates, dtype=np.float64)
    csd = np.asarray(csd, dtype=np.complex128)
    n, m, N = templates.shape
    assert csd.shape == (n, n, N)
    df = fs / N
    S = _fft_norm(templates, fs)                                   # [n, m, N]
    if integralnorm:
        S = S / S[:, :, :1]
    # bins with a non-finite csd diagonal (ignored_frequency_peaks of OFBase.set_csd) carry no weight
    ignored = ~np.all(np.isfinite(np.real(csd[np.arange(n), np.arange(n), :])), axis=0)
    safe = np.where(ignored[None, None, :], np.eye(n)[:, :, None], csd)
    iS = np.linalg.inv(np.transpose(safe, (2, 0, 1)))              # [N, n, n]
    iS[ignored] = 0.0
    if coupling == 'AC':
        iS[0] = 0.0
    Phi = np.einsum('bik,kba->iak', np.conj(S), iS)                # [m, n, N]
    P = np.real(np.einsum('iak,ajk->ij', Phi, S)) * df
    return {'n': n, 'm': m, 'N': N, 'fs': float(fs), 'df': df, 'pre': int(pretrigger_samples), 'S': S, 'iS': iS,
            'Phi': Phi, 'P': P, 'Pinv': np.linalg.inv(P)}


def ofnxm_batch(traces, setup, window=(None, None, False)):
    """traces [B, n, N].  Returns dict: chi0[B]; amps[B, m], ind[B], t0[B], chi2[B] (with delay, inside / outside
    the window given in rolled indices); amps0[B, m], chi2_0[B] (no delay)."""
    traces = np.asarray(traces, dtype=np.float64)
    if traces.ndim == 2:
        traces = traces[None]
    B, n, N = traces.shape
    st = setup
    df, pre, Pinv = st['df'], st['pre'], st['Pinv']
    V = _fft_norm(traces, st['fs'])                                                # [B, n, N]
    Q = np.einsum('iak,bak->bik', st['Phi'], V)                                    # [B, m, N]
    q = np.real(np.fft.ifft(Q * N, axis=-1)) * df
    chi0 = np.real(np.einsum('bak,kac,bck->b', np.conj(V), st['iS'], V)) * df
    q = np.roll(q, pre, axis=-1)
    dchi = np.einsum('bit,ij,bjt->bt', q, Pinv, q)
    chi2 = chi0[:, None] - dchi
    amps_t = np.einsum('ij,bjt->bit', Pinv, q)
    wmin, wmax, outside = window
    lo, hi = of_window_bounds(N, wmin, wmax)
    mask = np.zeros(N, dtype=bool)
    mask[lo:hi] = True
    if outside:
        mask = ~mask
    ind = np.argmin(np.where(mask[None, :], chi2, np.inf), axis=-1)
    rows = np.arange(B)
    return {'chi0': chi0, 'ind': ind, 't0': (ind - pre) / st['fs'], 'amps': amps_t[rows, :, ind], 'chi2': chi2[rows, ind],
            'amps0': amps_t[:, :, pre], 'chi2_0': chi2[:, pre], 'dchi2_td': dchi}
