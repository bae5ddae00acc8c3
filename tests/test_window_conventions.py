"""
The two window conventions that meet in `of1x1_constrained`, enumerated on the README windows
(reference README.md:81-96) with usec x fs products that are NOT integers (VERDICT r1, weak #1 iv).

The pipeline always hands BOTH forms to the extractor (reference features.py:762-788): the usec values of the YAML block and
the indices `_get_window_indices` made of them with python `int()` (truncation toward zero, features.py:1243-1344).
Inside `qp.OF1x1.calc` (not in the reference tree) the usec form is converted again -- floor for the lower, ceil for
the upper edge -- and wins over the index form.  For non-integer products the two differ by one sample per edge; which one
QETpy really uses cannot be read off the reference (parity unpinned), so what this file does is (a) enumerate the cases,
(b) pin that the oracle and the product's host code agree with each other in every one of them, for both precedences.
"""
import numpy as np
import pytest

from oracle.windows import get_window_indices as oracle_indices
from oracle.of1x1 import OFBaseOracle, OF1x1Oracle, of_window_bounds
from detprocess_b200.core import algorithms as A
from detprocess_b200.utils.utils import get_window_indices as extract_window_indices

FS = 1.25e6
N, PRE = 32768, 16384
# README windows (+-400, -1000, +-500 us) and neighbours whose product with 1.25 MHz is fractional
USEC = [-1000.0, -500.0, -400.0, 400.0, 500.0, -400.3, 400.3, -399.9, 399.9, 0.4, -0.4, 123.456, -987.654, 0.0]


class _Base:
    """the part of OFBaseBatch `_of_window` reads"""
    def nb_samples(self): return N
    def sample_rate(self): return FS
    def pretrigger_samples(self, channel, tag): return PRE


def _oracle_usec_window(lo_us, hi_us):
    """candidate range of OF1x1Oracle.calc when the usec form is given (floor / ceil, then the half-open slice)"""
    wmin = int(np.floor(PRE + lo_us * FS * 1e-6))
    wmax = int(np.ceil(PRE + hi_us * FS * 1e-6))
    return of_window_bounds(N, wmin, wmax)


@pytest.mark.parametrize('lo_us', [u for u in USEC if u <= 0])
@pytest.mark.parametrize('hi_us', [u for u in USEC if u >= 0])
def test_usec_form_and_index_form_enumerated(lo_us, hi_us):
    if lo_us == 0.0 and hi_us == 0.0:
        return
    # (1) what the pipeline computes (int() truncation): oracle restatement == the product's host code
    imin, imax = oracle_indices(N, PRE, FS, window_min_from_trig_usec=lo_us, window_max_from_trig_usec=hi_us)
    pmin, pmax = extract_window_indices(N, PRE, FS, window_min_from_trig_usec=lo_us, window_max_from_trig_usec=hi_us)
    assert (imin, imax) == (pmin, pmax)
    assert imin == PRE + int(lo_us * FS * 1e-6) and imax == PRE + int(hi_us * FS * 1e-6)
    # (2) usec form inside the fit: floor / ceil; product == oracle
    olo, ohi = _oracle_usec_window(lo_us, hi_us)
    plo, phi = A._of_window(_Base(), 'c', 'default', lo_us, hi_us, imin, imax)      # both forms given: usec wins
    assert (plo, phi) == (olo, ohi)
    # (3) index form alone (what an external caller of the extractor may pass)
    ilo, ihi = A._of_window(_Base(), 'c', 'default', None, None, imin, imax)
    assert (ilo, ihi) == of_window_bounds(N, imin, imax)
    # (4) the enumeration itself: the two forms agree iff both products are integers
    frac_lo = (lo_us * FS * 1e-6) % 1 != 0
    frac_hi = (hi_us * FS * 1e-6) % 1 != 0
    assert (olo == ilo) == (not frac_lo)         # floor(-x) = -ceil(x): one sample earlier than int() for x < 0
    assert (ohi == ihi) == (not frac_hi)         # ceil(x): one sample later than int() for x > 0
    if frac_lo:
        assert olo == ilo - 1
    if frac_hi:
        assert ohi == ihi + 1


def test_oracle_object_api_applies_the_same_precedence():
    """OF1x1Oracle.calc (the per-event object the CPU baseline drives): usec form wins; the candidate set it searches
    is the one `_oracle_usec_window` describes"""
    n, pre, fs = 2048, 1000, 1.25e6
    rng = np.random.default_rng(2)
    from detprocess_b200.synth import make_template, make_psd, make_traces
    template = make_template(n, fs, nb_pretrigger=pre)
    psd = make_psd(n, fs)
    x = make_traces(1, template, psd, fs, rng, pulse_fraction=0.0)[0]
    ofb = OFBaseOracle(fs)
    ofb.set_csd('c', psd)
    ofb.add_template('c', template, pretrigger_samples=pre)
    ofb.calc_phi('c')
    ofb.update_signal('c', x)
    of = OF1x1Oracle(ofb, 'c')
    amps, chi2, _ = of._arrays()
    for lo_us, hi_us in [(-40.3, 40.3), (-8.0, 8.0), (-0.4, 0.4)]:
        of.calc(window_min_from_trig_usec=lo_us, window_max_from_trig_usec=hi_us, window_min_index=pre - 1, window_max_index=pre + 1,
                lgc_fit_nodelay=False)
        lo = int(np.floor(pre + lo_us * fs * 1e-6))
        hi = int(np.ceil(pre + hi_us * fs * 1e-6))
        ind = lo + int(np.argmin(chi2[lo:hi]))
        assert of.get_result_withdelay()[1] == pytest.approx((ind - pre) / fs, abs=1e-15)
