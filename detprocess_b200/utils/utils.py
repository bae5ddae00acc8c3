"""
Host-side helpers with the semantics of the reference's ``detprocess/utils/utils.py``
(channel-name algebra :70-184, window indices :189-301) and pytesio's
``convert_length_msec_to_samples`` (used by ``process/config.py:10``).
"""
import numpy as np

ALLOWED_SEPARATORS = [',', '|', '+', '-']


def unique_list(seq):
    seen = set()
    out = []
    for x in seq:
        if x not in seen:
            seen.add(x)
            out.append(x)
    return out


def convert_length_msec_to_samples(length_msec, sample_rate):
    """20 ms @ 1.25 MHz -> 25000 (matches the reference's saved YamlConfig output)."""
    return int(round(float(length_msec) * 1e-3 * float(sample_rate)))


def split_channel_name(channel_name, available_channels=None, separator=None, label=None):
    """
    Split "a,b" / "a+b" / "a-b" / "a|b" into individual channels.
    Returns (channel_list, separator) like the reference (separator None when the name
    is a single channel).  Same checks: unknown separators and unknown channels raise
    ValueError; "-" needs ``available_channels``.
    """
    channel_name = channel_name.replace(' ', '')
    if separator is not None and separator not in ALLOWED_SEPARATORS:
        raise ValueError(f'ERROR: separator "{separator}" not recognized. '
                         f'Allowed separator {ALLOWED_SEPARATORS} ')
    if not any(sep in channel_name for sep in ALLOWED_SEPARATORS):
        return [channel_name], None

    if available_channels is None:
        if separator is None:
            raise ValueError('ERROR: separator required when "available_channels" not provided! ')
        if separator == '-':
            raise ValueError('ERROR: "available_channels" required when using separator "-"')
        if separator == '+' and (',' in channel_name or '|' in channel_name):
            raise ValueError(f'ERROR: Channels cannot be split with {separator} before channels '
                             f'split with "," and "|"')
        return channel_name.split(separator), separator

    if channel_name in available_channels or channel_name == 'all':
        return [channel_name], None

    # strip known channel names; what is left must be separators only
    remainder = channel_name
    found = []
    for chan in available_channels:
        if chan in remainder:
            remainder = remainder.replace(chan, '')
            found.append(chan)
    seps = list(set(remainder))
    bad = [s for s in seps if s not in ALLOWED_SEPARATORS]
    if bad:
        raise ValueError(f'ERROR: Unidentified channel "{channel_name}" in yaml file! '
                         f'Perhaps not in raw data? Available channels = {available_channels}')

    if separator is None:
        if len(seps) == 1:
            sep = seps[0]
            if sep != '-':
                found = channel_name.split(sep)
            return found, sep
        return found, seps

    if separator not in channel_name:
        return [channel_name], None
    if separator != '-':
        return channel_name.split(separator), separator
    if any(s in channel_name for s in ('|', '+', ',')):
        raise ValueError('Multiple separators available, split first with other separators before "-"')
    return list(found), separator


def get_window_indices(nb_samples, nb_pretrigger_samples, fs,
                       window_min_from_start_usec=None, window_min_to_end_usec=None,
                       window_min_from_trig_usec=None, window_max_from_start_usec=None,
                       window_max_to_end_usec=None, window_max_from_trig_usec=None, **kwargs):
    """
    usec -> sample indices exactly as FeatureProcessing._get_window_indices
    (reference process/features.py:1243-1344): priority from_start > to_end > from_trig,
    int() truncation, clamp to [0, N-1], ValueError if max < min.
    """
    def resolve(from_start, to_end, from_trig, default):
        idx = default
        if from_start is not None:
            idx = int(from_start * fs * 1e-6)
        elif to_end is not None:
            idx = nb_samples - abs(int(to_end * fs * 1e-6)) - 1
        elif from_trig is not None:
            idx = nb_pretrigger_samples + int(from_trig * fs * 1e-6)
        return min(max(idx, 0), nb_samples - 1)

    lo = resolve(window_min_from_start_usec, window_min_to_end_usec, window_min_from_trig_usec, 0)
    hi = resolve(window_max_from_start_usec, window_max_to_end_usec, window_max_from_trig_usec, nb_samples - 1)
    if hi < lo:
        raise ValueError('ERROR window calculation: max index smaller than min!Check configuration!')
    return lo, hi


def columns_to_frame(parts):
    """ONE DataFrame from the per-batch column dicts (name -> ndarray [B] or scalar)"""
    import pandas as pd
    parts = [p for p in parts if p]
    if not parts:
        return pd.DataFrame()
    names = list(parts[0].keys())
    sizes = [next((len(v) for v in p.values() if isinstance(v, np.ndarray) and v.ndim), 1) for p in parts]
    out = {}
    for c in names:
        vals = [np.asarray(p[c]) if np.ndim(p[c]) else np.full(nb, p[c]) for p, nb in zip(parts, sizes)]
        out[c] = vals[0] if len(vals) == 1 else np.concatenate(vals)
    return pd.DataFrame(out)


def cleanup_freq_ranges(f_lims):
    """Frequency ranges of ``psd_amp`` / ``psd_peaks`` and their feature-name suffixes (reference utils/utils.py:437-470):
    a number or a one-element list is a single frequency (name ``'<f>'``), a pair is ordered low -> high (name
    ``'<lo>_<hi>'``, rounded); signs are dropped and a name that came before is skipped."""
    if not isinstance(f_lims, list):
        f_lims = [f_lims]
    ranges, names = [], []
    for item in f_lims:
        vals = [item] if isinstance(item, (int, float)) else list(item)
        lo = abs(vals[0])
        if len(vals) == 2:
            hi = abs(vals[1])
            lo, hi = min(lo, hi), max(lo, hi)
            name, rng = f'{round(lo)}_{round(hi)}', [lo, hi]
        else:
            name, rng = f'{round(lo)}', [lo]
        if name not in names:
            names.append(name)
            ranges.append(rng)
    return ranges, names


def get_ind_freq_ranges(freq_ranges, freqs):
    """``[ind_low, ind_high)`` into ``freqs`` for every range (reference utils/utils.py:475-505): nearest bins of the two
    edges (single frequency: its nearest bin), ordered; an empty range is widened by one bin upwards, or downwards at the
    end of the array."""
    freqs = np.asarray(freqs)
    out = []
    for rng in freq_ranges:
        lo = int(np.argmin(np.abs(freqs - abs(rng[0]))))
        hi = lo + 1 if len(rng) != 2 else int(np.argmin(np.abs(freqs - abs(rng[1]))))
        if lo > hi:
            lo, hi = hi, lo
        if lo == hi:
            if hi < len(freqs) - 1:
                hi += 1
            elif lo > 0:
                lo -= 1
            else:
                raise ValueError('Frequency range too narrow or outside bounds.')
        out.append([lo, hi])
    return out
