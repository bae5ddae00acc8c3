"""
CPU oracle of the continuous-stream optimal-filter trigger (TEST INFRASTRUCTURE ONLY, see
oracle/__init__.py).  Restates, for one trigger channel and one amplitude,

  OptimumFilterTrigger.__init__      detprocess/core/oftrigger.py:489-499  (phi_td, norm)
  OptimumFilterTrigger.update_trace  detprocess/core/oftrigger.py:655-679  (oaconvolve, dchi2, padding)
  _getchangeslessthanthresh          detprocess/core/oftrigger.py:29-74
  find_triggers_once                 detprocess/core/oftrigger.py:937-1016 (threshold, groups, arg-max)

The FIR filtering is scipy.signal.oaconvolve itself -- the library call the reference makes --
so that stage is pinned by construction.  phi / W / iW come from QETpy in the reference
(not in the tree, parity unpinned): here they are inputs.
"""
import numpy as np
from scipy import special, stats
from scipy.signal import oaconvolve


def phi_td_from_phi_fd(phi_fd):
    """oftrigger.py:491-493: zero the DC bin, real part of the inverse FFT."""
    phi = np.array(phi_fd, dtype=np.complex128)
    phi[0] = 0
    return np.fft.ifft(phi).real


def filter_trace(trace, phi_td, iw, w, padding=True):
    """oftrigger.py:659-679 for n_channels = m_amplitudes = 1.  Returns (filtered, delta_chi2)."""
    trace = np.asarray(trace, dtype=np.float64)
    v_td = oaconvolve(trace[None, :], np.asarray(phi_td)[None, :], mode='same', axes=-1)[0]
    filtered = iw * v_td
    dchi2 = filtered * w * filtered
    if padding:
        cut_len = len(phi_td)
        dchi2[:cut_len] = 0.0
        dchi2[-(cut_len) + (cut_len + 1) % 2:] = 0.0
    return filtered, dchi2


def chi2_threshold(thresh, m_amplitudes=1):
    """oftrigger.py:961-965."""
    if thresh < 25:
        survival_fraction = stats.norm.sf(thresh) * 2
        return special.gammainccinv(m_amplitudes / 2, survival_fraction) * 2
    return thresh ** 2


def getchangeslessthanthresh(x, threshold):
    """oftrigger.py:29-74."""
    diff = x[1:] - x[:-1]
    inds = np.where(diff > threshold)[0] + 1
    start_inds = np.zeros(len(inds) + 1, dtype=int)
    start_inds[1:] = inds
    end_inds = np.zeros(len(inds) + 1, dtype=int)
    end_inds[-1] = len(x)
    end_inds[:-1] = inds
    return np.array(list(zip(start_inds, end_inds)))


def find_triggers_once(dchi2, filtered, thr_chi2, pileup_window, index_shift, fs):
    """oftrigger.py:967-1016.  Returns dict of arrays (trigger_index, trigger_time, amplitude, delta_chi2)."""
    mask = dchi2 > thr_chi2
    triggers = np.where(mask)[0]
    out = {'trigger_index': [], 'trigger_time': [], 'trigger_amplitude': [], 'trigger_delta_chi2': []}
    ranges = getchangeslessthanthresh(triggers, pileup_window)
    for a, b in ranges:
        if b > a:
            evt_inds = triggers[a:b]
            evt_ind = evt_inds[np.argmax(dchi2[evt_inds])]
            out['trigger_index'].append(evt_ind + index_shift)
            out['trigger_time'].append((evt_ind + index_shift) / fs)
            out['trigger_amplitude'].append(filtered[evt_ind])
            out['trigger_delta_chi2'].append(dchi2[evt_ind])
    return {k: np.asarray(v) for k, v in out.items()}


def getchangeslessthandynamicthresh(x, amplitudes, threshold_function):
    """oftrigger.py:78-141: ranges of `x` whose consecutive gaps stay within a window that depends on the largest
    delta-chi2 met so far in the open range (`threshold_function(max)` -> window in samples)."""
    start_inds, end_inds = [], []
    current_start = 0
    for i in range(1, len(x)):
        max_amplitude = np.max(amplitudes[current_start:i + 1])
        if (x[i] - x[i - 1]) > threshold_function(max_amplitude):
            start_inds.append(current_start)
            end_inds.append(i)
            current_start = i
    start_inds.append(current_start)
    end_inds.append(len(x))
    return np.array(list(zip(start_inds, end_inds)))


def find_triggers_once_dynamic(dchi2, filtered, thr_chi2, threshold_function, index_shift, fs):
    """find_triggers_once with dynamic=True (oftrigger.py:975-979): same arg-max per range, dynamic ranges."""
    mask = dchi2 > thr_chi2
    triggers = np.where(mask)[0]
    out = {'trigger_index': [], 'trigger_time': [], 'trigger_amplitude': [], 'trigger_delta_chi2': []}
    for a, b in getchangeslessthandynamicthresh(triggers, dchi2[mask], threshold_function):
        if b > a:
            evt_inds = triggers[a:b]
            evt_ind = evt_inds[np.argmax(dchi2[evt_inds])]
            out['trigger_index'].append(evt_ind + index_shift)
            out['trigger_time'].append((evt_ind + index_shift) / fs)
            out['trigger_amplitude'].append(filtered[evt_ind])
            out['trigger_delta_chi2'].append(dchi2[evt_ind])
    return {k: np.asarray(v) for k, v in out.items()}


def residual_delta_chi2(dchi2, filtered, first_pass_index, template, phi_td, iw, w, saturated=None):
    """oftrigger.py:772-820 for one channel / one amplitude: for every first-pass trigger (their *shifted* indices, as
    the reference uses them) that is not saturated, the delta-chi2 trace a pulse of the filtered amplitude at that index
    would produce (template x amplitude -> same FIR -> iw, w) is subtracted, aligned so that its maximum sits on the
    trigger index.  Returns the residual trace (a copy)."""
    res = np.array(dchi2, dtype=np.float64)
    nt = len(template)
    for k, trigger_index in enumerate(first_pass_index):
        if saturated is not None and saturated[k]:
            continue
        amp = filtered[trigger_index]
        trigger_trace = np.asarray(template, dtype=np.float64) * amp
        v_td = oaconvolve(trigger_trace[None, :], np.asarray(phi_td)[None, :], mode='same', axes=-1)[0]
        f = iw * v_td
        d = f * w * f
        j = int(np.argmax(d))
        res[trigger_index - j:trigger_index - j + nt] -= d
    return res


def saturated_flags(raw_lpf, first_pass_index, nt, saturation_amplitude, positive_pulses=True):
    """oftrigger.py:776-786: a trigger is saturated when the low-passed raw trace crosses the saturation amplitude
    within nt/4 samples of it."""
    flags = []
    for trigger_index in first_pass_index:
        seg = raw_lpf[trigger_index - int(nt / 4):trigger_index + int(nt / 4)]
        if positive_pulses:
            flags.append(bool(np.sum(seg > saturation_amplitude) > 0))
        else:
            flags.append(bool(np.sum(seg < -1 * saturation_amplitude) > 0))
    return np.asarray(flags, dtype=bool)


def combine_triggers(first, second):
    """combine_trigger_data (oftrigger.py:262-320): the first-pass triggers followed by those second-pass triggers
    whose index is not among the first-pass indices."""
    new = ~np.isin(second['trigger_index'], first['trigger_index'])
    return {k: np.concatenate([np.asarray(first[k]), np.asarray(second[k])[new]]) for k in first}
