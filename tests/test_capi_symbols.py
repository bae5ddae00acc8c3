"""The C-ABI library loads and exports every symbol include/detprocess_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'detprocess_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dp_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ('dp_of_plan_create', 'dp_of1x1_batch', 'dp_window_reduce_batch', 'dp_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from detprocess_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(dll, name), f'{name} declared in the header but not exported'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    assert dll.dp_version() >= 100


def test_plan_argument_errors_without_gpu():
    """Host-side validation only (no compute, no device)."""
    import numpy as np
    from detprocess_b200.core.plans import OFPlan, ReducePlan
    with pytest.raises(NotImplementedError):
        OFPlan(25002, 1.25e6)                     # half does not factor into 2, 3, 4, 5
    with pytest.raises(ValueError):
        OFPlan(4096, 1.25e6, precision='f16')
    p = OFPlan(4096, 1.25e6)
    with pytest.raises(ValueError):
        p.set_psd(0, np.ones(100))                # wrong length
    with pytest.raises(ValueError):
        p.set_psd(0, -np.ones(4096))              # non-positive psd
    p.set_psd(0, np.ones(4096))
    t = p.add_template(0, np.exp(-np.arange(4096) / 100.0), pretrigger_samples=2048)
    with pytest.raises(ValueError):
        p.add_fit(0, t, 10, 10)                   # empty window
    with pytest.raises(ValueError):
        p.add_fit(0, 3, 0, 10)                    # unknown template
    r = ReducePlan(4096, 1.25e6)
    with pytest.raises(ValueError):
        r.add(0, 'maximum', 5, 5)                 # numpy: zero-size array to reduction operation
    r.add(0, 'baseline', 5, 5)                    # numpy: nan + warning, allowed


def test_nxm_plan_argument_errors_and_p_matrix_without_gpu():
    """NxM plan: host-side validation and the template matrix P (no device needed before finalize): P equals the
    oracle's, a singular csd and degenerate templates are refused."""
    import numpy as np
    from detprocess_b200 import _lib
    from detprocess_b200.core.plans import NxMPlan
    from detprocess_b200.synth import SynthNxM
    from oracle.ofnxm import ofnxm_setup
    S = SynthNxM(16384, 2, 2)
    with pytest.raises(NotImplementedError):
        NxMPlan(4096, S.fs, 2, 2)                              # unsupported trace length
    with pytest.raises(ValueError):
        NxMPlan(16384, S.fs, 5, 1)                             # more channels than the kernel is built for
    p = NxMPlan(16384, S.fs, 2, 2)
    with pytest.raises(ValueError):
        p.set_filter(S.templates[:, :1], S.csd)                # wrong template shape
    with pytest.raises(_lib.DetprocessB200Error):
        p.p_matrix()                                           # no filter yet
    p.set_filter(S.templates, S.csd, S.nb_pretrigger, 'AC')
    P, Pinv = p.p_matrix()
    st = ofnxm_setup(S.templates, S.csd, S.fs, S.nb_pretrigger)
    assert np.allclose(P, st['P'], rtol=1e-10) and np.allclose(Pinv, st['Pinv'], rtol=1e-8)
    with pytest.raises(ValueError):
        p.set_window(100, 50)                                  # lo > hi
    with pytest.raises(ValueError):
        p.set_window(0, 16385)
    bad = S.csd.copy()
    bad[1] = bad[0]                                            # rank-deficient at every bin
    bad[:, 1] = bad[:, 0]
    with pytest.raises(ValueError):
        NxMPlan(16384, S.fs, 2, 2).set_filter(S.templates, bad)
    same = S.templates.copy()
    same[:, 1] = same[:, 0]                                    # two identical templates: P is singular
    with pytest.raises(ValueError):
        NxMPlan(16384, S.fs, 2, 2).set_filter(same, S.csd)
    with pytest.raises(_lib.DetprocessB200Error):
        p.finalize()                                           # no CUDA device here: no CPU fallback
