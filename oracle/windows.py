"""
CPU oracle: window (usec -> sample index) arithmetic.  TEST INFRASTRUCTURE ONLY.

Restates ``FeatureProcessing._get_window_indices``
(``detprocess/process/features.py:1243-1344``; duplicated in
``detprocess/utils/utils.py:189-301``): priority from_start > to_end > from_trig,
python ``int()`` truncation toward zero, clamp to [0, N-1], error if max < min.
Pinned: exactly restatable from the reference tree (integer arithmetic).
"""


def get_window_indices(nb_samples, nb_pretrigger_samples, fs,
                       window_min_from_start_usec=None,
                       window_min_to_end_usec=None,
                       window_min_from_trig_usec=None,
                       window_max_from_start_usec=None,
                       window_max_to_end_usec=None,
                       window_max_from_trig_usec=None,
                       **kwargs):
    def one(from_start, to_end, from_trig, default):
        idx = default
        if from_start is not None:
            idx = int(from_start * fs * 1e-6)
        elif to_end is not None:
            idx = nb_samples - abs(int(to_end * fs * 1e-6)) - 1
        elif from_trig is not None:
            idx = nb_pretrigger_samples + int(from_trig * fs * 1e-6)
        if idx < 0:
            idx = 0
        elif idx > nb_samples - 1:
            idx = nb_samples - 1
        return idx

    min_index = one(window_min_from_start_usec, window_min_to_end_usec,
                    window_min_from_trig_usec, 0)
    max_index = one(window_max_from_start_usec, window_max_to_end_usec,
                    window_max_from_trig_usec, nb_samples - 1)
    if max_index < min_index:
        raise ValueError('ERROR window calculation: max index smaller than min!'
                         'Check configuration!')
    return min_index, max_index
