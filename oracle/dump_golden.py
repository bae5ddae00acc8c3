"""
Close the "parity unpinned" loop where the real reference stack is installed.

    python oracle/dump_golden.py tests/golden/of1x1_qetpy.npz

Runs UPSTREAM QETpy (``qetpy>=1.8.6``, the package detprocess delegates the OF maths to,
reference setup.py:73) on the seeded synthetic inputs of ``detprocess_b200.synth`` through exactly the
calls the reference makes (processing_data.py:278-381, 731-772; algorithms.py:331-341, 410-421, 533-558;
noise.py:344) and stores inputs' seeds + outputs.  ``tests/test_oracle_of.py::test_against_qetpy_golden``
compares ``oracle/of1x1.py`` with the file when it exists (it is skipped otherwise).  QETpy is NOT
available in this build environment (no network), so the file is not committed: the oracle stays
pinned by its known-answer tests only, as stated in DESIGN.md section 2.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detprocess_b200.synth import SynthSetup, make_traces  # noqa: E402


def main(out_path, nb_samples=4096, n_events=16, seed=12345):
    try:
        import qetpy as qp
    except ImportError as e:                                           # pragma: no cover
        raise SystemExit(f'qetpy is not installed here ({e}); run this where detprocess+QETpy are available')
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    traces = make_traces(n_events, S.template, S.psd, S.fs, np.random.default_rng(seed))
    ofb = qp.OFBase(S.fs)
    ofb.set_csd('ch', S.psd, coupling='AC')
    ofb.add_template('ch', S.template, template_tag='default', pretrigger_samples=pre)
    ofb.calc_phi('ch', template_tag='default')
    res = {k: [] for k in ('amp_nodelay', 'chi2_nodelay', 'amp_un', 't0_un', 'chi2_un', 'lowchi2_un',
                           'amp_con', 't0_con', 'chi2_con', 'chi2nopulse', 'ampres', 'timeres')}
    for x in traces:
        ofb.clear_signal()
        ofb.update_signal('ch', x, calc_fft=True)
        ofb.calc_signal_filt('ch')
        ofb.calc_signal_filt_td('ch')
        OF = qp.OF1x1(of_base=ofb, channel='ch', template_tag='default')
        OF.calc(lowchi2_fcutoff=10000, lgc_fit_withdelay=False, lgc_fit_nodelay=True)
        a, _, c, _ = OF.get_result_nodelay()
        res['amp_nodelay'].append(a), res['chi2_nodelay'].append(c)
        OF.calc(lowchi2_fcutoff=10000, lgc_fit_withdelay=True, lgc_fit_nodelay=False)
        a, t, c, lc = OF.get_result_withdelay()
        res['amp_un'].append(a), res['t0_un'].append(t), res['chi2_un'].append(c), res['lowchi2_un'].append(lc)
        OF.calc(window_min_index=pre - 500, window_max_index=pre + 500, lowchi2_fcutoff=10000,
                lgc_fit_withdelay=True, lgc_fit_nodelay=False)
        a, t, c, _ = OF.get_result_withdelay()
        res['amp_con'].append(a), res['t0_con'].append(t), res['chi2_con'].append(c)
        res['chi2nopulse'].append(OF.get_chisq_nopulse())
        res['ampres'].append(OF.get_energy_resolution()), res['timeres'].append(OF.get_time_resolution())
    freqs, psd = qp.calc_psd(traces, fs=S.fs, folded_over=False)
    np.savez(out_path, nb_samples=nb_samples, n_events=n_events, seed=seed, qetpy_version=getattr(qp, '__version__', '?'),
             psd_of_traces=psd, **{k: np.asarray(v) for k, v in res.items()})
    print('wrote', out_path)
    dump_nxm(qp, os.path.join(os.path.dirname(out_path), 'ofnxm_qetpy.npz'), nb_samples, n_events, seed)
    dump_utils(qp, os.path.join(os.path.dirname(out_path), 'utils_qetpy.npz'), nb_samples, seed)


def dump_nxm(qp, out_path, nb_samples, n_events, seed):
    """NxM filter and CSD through the calls the reference makes: OFBase.set_csd / add_template on the joint channel
    (processing_data.py:294-381), qp.OFnxm(...).calc() / get_fit_withdelay / get_fit_nodelay (algorithms.py:246-263),
    qp.calc_csd (noise.py:452).  Settles the csd convention (E[X_a conj X_b] here, its conjugate in scipy.signal.csd)
    and the chi2 normalisation of oracle/ofnxm.py."""
    from detprocess_b200.synth import SynthNxM
    S = SynthNxM(nb_samples, 2, 2)
    pre = S.nb_pretrigger
    x = S.traces(n_events, np.random.default_rng(seed))
    chan = 'a|b'
    ofb = qp.OFBase(S.fs)
    ofb.set_csd(chan, S.csd, coupling='AC')
    ofb.add_template(chan, S.templates, template_tag='shared', pretrigger_samples=pre)
    res = {k: [] for k in ('amps', 't0', 'chi2', 'amps0', 'chi2_0')}
    for ev in x:
        ofb.clear_signal()
        ofb.update_signal(chan, ev, calc_fft=True)
        OF = qp.OFnxm(of_base=ofb, channels=chan, template_tag='shared', verbose=False)
        OF.calc()
        a, t, c = OF.get_fit_withdelay(window_min_index=pre - 500, window_max_index=pre + 500, lgc_outside_window=False)
        a0, _, c0 = OF.get_fit_nodelay()
        res['amps'].append(a), res['t0'].append(t), res['chi2'].append(c), res['amps0'].append(a0), res['chi2_0'].append(c0)
    noise = S.traces(64, np.random.default_rng(seed + 1), pulse_fraction=0.0)
    _, csd = qp.calc_csd(noise, fs=S.fs, folded_over=False)
    np.savez(out_path, nb_samples=nb_samples, n_events=n_events, seed=seed, csd_of_noise=csd,
             **{k: np.asarray(v) for k, v in res.items()})
    print('wrote', out_path)


def dump_utils(qp, out_path, nb_samples, seed):
    """The two QETpy helpers the round-2 additions lean on: ``qp.utils.fold_spectrum`` (psd_amp, algorithms.py:1019) and
    ``qp.utils.lowpassfilter(x, cut_off_freq=50e3, fs)`` (saturation test of the residual re-trigger, oftrigger.py:622-627).
    Settles oracle/psd.py::fold_spectrum and the first-order Butterworth + filtfilt the product recalls."""
    rng = np.random.default_rng(seed)
    fs = 1.25e6
    spec = rng.random(nb_samples)
    spec_odd = rng.random(nb_samples - 1)
    f_even, fold_even = qp.utils.fold_spectrum(spec, fs)
    f_odd, fold_odd = qp.utils.fold_spectrum(spec_odd, fs)
    x = rng.standard_normal(8 * nb_samples)
    np.savez(out_path, seed=seed, spec=spec, spec_odd=spec_odd, f_even=f_even, fold_even=fold_even, f_odd=f_odd, fold_odd=fold_odd,
             lpf_in=x, lpf_out=qp.utils.lowpassfilter(x, cut_off_freq=50e3, fs=fs))
    print('wrote', out_path)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'tests/golden/of1x1_qetpy.npz')
