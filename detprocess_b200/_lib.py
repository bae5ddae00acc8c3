"""
ctypes binding of libdetprocess_b200.so (C ABI in include/detprocess_b200.h).

There is NO fallback: if the shared library is missing or fails to load, importing
this module raises.  Build it with ``python -m detprocess_b200.build`` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_C', os.environ.get('DP_LIB_NAME', 'libdetprocess_b200.so'))

DP_OK = 0
DP_PREC_F64, DP_PREC_F32 = 0, 1
DP_IN_F64, DP_IN_F32, DP_IN_I16 = 0, 1, 2
DP_OP_BASELINE, DP_OP_INTEGRAL, DP_OP_MAXIMUM, DP_OP_MINIMUM = 0, 1, 2, 3
DP_FIT_NOUT = 5


class DetprocessB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). '
            'Run `python -m detprocess_b200.build`.')
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i, _d, _ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
_ip, _dp, _fp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)

# every symbol include/detprocess_b200.h declares, with its signature
SIGNATURES = {
    'dp_last_error': (C.c_char_p, []),
    'dp_version': (_i, []),
    'dp_device_count': (_i, [_ip]),
    'dp_of_plan_create': (_i, [C.POINTER(_vp), _i, _d, _i, _i]),
    'dp_of_plan_destroy': (None, [_vp]),
    'dp_of_plan_set_psd': (_i, [_vp, _i, _vp, _i]),
    'dp_of_plan_add_template': (_i, [_vp, _i, _vp, _i, _i, _ip]),
    'dp_of_plan_add_fit': (_i, [_vp, _i, _i, _i, _i, _i, _ip]),
    'dp_of_plan_add_fit_ex': (_i, [_vp, _i, _i, _i, _i, _i, _d, _ip]),
    'dp_of_plan_set_lowchi2_fcutoff': (_i, [_vp, _d]),
    'dp_of_plan_set_neighbours': (_i, [_vp, _i]),
    'dp_of_plan_neighbour_offset': (_i, [_vp, _i, _i, _ip]),
    'dp_of_plan_set_adc_conversion': (_i, [_vp, _i, _d, _d]),
    'dp_csd_plan_create': (_i, [C.POINTER(_vp), _i, _d, _i, _i, _i]),
    'dp_csd_plan_destroy': (None, [_vp]),
    'dp_csd_plan_set_scale': (_i, [_vp, _d]),
    'dp_csd_reset': (_i, [_vp, _vp]),
    'dp_csd_accumulate': (_i, [_vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong, _vp, _vp]),
    'dp_csd_get_sums': (_i, [_vp, _vp, _vp, _vp]),
    'dp_csd_plan_last_kernel_ms': (_i, [_vp, C.POINTER(C.c_float)]),
    'dp_nxm_plan_create': (_i, [C.POINTER(_vp), _i, _d, _i, _i, _i]),
    'dp_nxm_plan_destroy': (None, [_vp]),
    'dp_nxm_plan_set_filter': (_i, [_vp, _vp, _vp, _i, _i]),
    'dp_nxm_plan_set_window': (_i, [_vp, _i, _i, _i]),
    'dp_nxm_plan_finalize': (_i, [_vp, _i]),
    'dp_nxm_plan_n_out': (_i, [_vp, _ip]),
    'dp_nxm_plan_get_p_matrix': (_i, [_vp, _vp, _vp]),
    'dp_ofnxm_batch': (_i, [_vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong, _vp, _vp]),
    'dp_nxm_plan_last_kernel_ms': (_i, [_vp, C.POINTER(C.c_float)]),
    'dp_of_plan_finalize': (_i, [_vp, _i]),
    'dp_of_plan_n_out': (_i, [_vp, _ip]),
    'dp_of_plan_fit_offset': (_i, [_vp, _i, _i, _ip]),
    'dp_of_plan_chi0_offset': (_i, [_vp, _i, _ip]),
    'dp_of_plan_get_phi': (_i, [_vp, _i, _i, _vp]),
    'dp_of_plan_get_norm': (_i, [_vp, _i, _i, _dp]),
    'dp_of_plan_get_template_fft': (_i, [_vp, _i, _i, _vp]),
    'dp_of1x1_batch': (_i, [_vp, _vp, _i, _ll, _ll, _vp, _vp]),
    'dp_of1x1_batch_host': (_i, [_vp, _vp, _i, _ll, _ll, _vp]),
    'dp_of1x1_windows': (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _vp]),
    'dp_of1x1_batch_ex': (_i, [_vp, _vp, _i, _ll, _ll, _vp, _ll, _vp, _ll, _vp, _vp]),
    'dp_of_plan_last_kernel_ms': (_i, [_vp, _fp]),
    'dp_of_plan_launch_count': (_i, [_vp, C.POINTER(_ll)]),
    'dp_reduce_plan_create': (_i, [C.POINTER(_vp), _i, _d, _i]),
    'dp_reduce_plan_destroy': (None, [_vp]),
    'dp_reduce_plan_add': (_i, [_vp, _i, _i, _i, _i, _ip]),
    'dp_reduce_plan_column': (_i, [_vp, _i, _i, _ip]),
    'dp_reduce_plan_finalize': (_i, [_vp, _i]),
    'dp_reduce_plan_n_out': (_i, [_vp, _ip]),
    'dp_window_reduce_batch': (_i, [_vp, _vp, _ll, _ll, _vp, _vp]),
    'dp_reduce_plan_set_adc_conversion': (_i, [_vp, _i, _d, _d]),
    'dp_window_reduce_batch_raw': (_i, [_vp, _vp, _i, _ll, _ll, _vp, _vp]),
    'dp_window_reduce_batch_ex': (_i, [_vp, _vp, _i, _ll, _ll, _vp, _ll, _vp, _ll, _vp, _vp]),
    'dp_channel_combine': (_i, [_vp, _i, _ll, _ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'dp_reduce_plan_last_kernel_ms': (_i, [_vp, _fp]),
    'dp_psd_plan_create': (_i, [C.POINTER(_vp), _i, _d, _i, _i]),
    'dp_psd_plan_destroy': (None, [_vp]),
    'dp_psd_plan_set_scale': (_i, [_vp, _d]),
    'dp_psd_reset': (_i, [_vp, _vp]),
    'dp_psd_accumulate': (_i, [_vp, _vp, _i, _ll, _ll, _vp, _vp]),
    'dp_psd_get_sums': (_i, [_vp, _vp, _vp, _vp]),
    'dp_psd_plan_last_kernel_ms': (_i, [_vp, _fp]),
    'dp_trigger_plan_create': (_i, [C.POINTER(_vp), _vp, _i, _d, _d, _i, _ll, _i]),
    'dp_trigger_plan_destroy': (None, [_vp]),
    'dp_trigger_plan_set_scale': (_i, [_vp, _d]),
    'dp_trigger_plan_geometry': (_i, [_vp, _ip, _ip]),
    'dp_trigger_run': (_i, [_vp, _vp, _ll, _d, _ll, _ll, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'dp_trigger_run_raw': (_i, [_vp, _vp, _i, _ll, _d, _ll, _ll, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'dp_trigger_plan_last_kernel_ms': (_i, [_vp, _fp, _fp]),
    'dp_band_plan_create': (_i, [C.POINTER(_vp), _i, _d, _vp, _vp, _i, _i]),
    'dp_band_plan_destroy': (None, [_vp]),
    'dp_band_amplitudes': (_i, [_vp, _vp, _i, _ll, _ll, _d, _d, _vp, _vp]),
    'dp_trigger_candidates': (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp]),
    'dp_trigger_filtered_at': (_i, [_vp, _vp, _i, _ll, _vp, _i, _vp, _vp]),
    'dp_trigger_residual_run': (_i, [_vp, _vp, _vp, _i, _vp, _i, _d, _ll, _ll, _vp, _vp, _vp, _i, _vp, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc):
    """Map C error codes to Python exceptions (DP_ERR_INVALID -> ValueError, like the reference)."""
    if rc == DP_OK:
        return
    msg = lib.dp_last_error().decode('utf-8', 'replace')
    if rc == 1:
        raise ValueError(msg)
    if rc == 4:
        raise NotImplementedError(msg)
    raise DetprocessB200Error(f'[{rc}] {msg}')
