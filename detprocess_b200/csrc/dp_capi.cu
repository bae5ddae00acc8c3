// C-ABI of libdetprocess_b200.so (see include/detprocess_b200.h).
// Host side: plan objects (one per reference qp.OFBase / per reduction feature set),
// device-table upload, persistent-grid launches of the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/detprocess_b200.h"
#include "dp_combine_kernel.cuh"
#include "dp_csd_kernel.cuh"
#include "dp_csd_launch.hpp"
#include "dp_nxm_launch.hpp"
#include "dp_nxm_plan.hpp"
#include "dp_of2_launch.hpp"
#include "dp_of_launch.hpp"
#include "dp_ofg_launch.hpp"
#include "dp_ofg_plan.hpp"
#include "dp_plan.hpp"
#include "dp_plan2.hpp"
#include "dp_psd2_kernel.cuh"
#include "dp_band_kernel.cuh"
#include "dp_band_launch.hpp"
#include "dp_psd_kernel.cuh"
#include "dp_reduce_plan.hpp"
#include "dp_trig_kernel.cuh"
#include "dp_trig_launch.hpp"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

// Every entry point that allocates or launches runs with the plan's device current and puts the caller's device back
// afterwards (the process-wide current device also belongs to PyTorch).
struct DeviceGuard {
    int prev = -1;
    bool changed = false, ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) {
            ok = cudaSetDevice(dev) == cudaSuccess;
            changed = ok;
        }
    }
    ~DeviceGuard() {
        if (changed && prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DP_ON_DEVICE(dev)        \
    DeviceGuard dp_dev_guard(dev); \
    if (!dp_dev_guard.ok) return fail(DP_ERR_CUDA, "cudaSetDevice failed")
#define DP_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(DP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t n = 0;
};

template <class V> int upload(std::vector<void*>& owned, const std::vector<V>& h, const V** out) {
    void* d = nullptr;
    const size_t bytes = std::max<size_t>(sizeof(V) * h.size(), 16);
    DP_CUDA(cudaMalloc(&d, bytes));
    owned.push_back(d);
    if (!h.empty()) DP_CUDA(cudaMemcpy(d, h.data(), sizeof(V) * h.size(), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const V*>(d);
    return DP_OK;
}

}  // namespace

// ============================================================================ OF plan
struct dp_of_plan {
    int N = 0;
    double fs = 0;
    int n_chan = 0;
    int precision = DP_PREC_F64;
    double fcut = 10000.0;
    std::vector<dpplan::Channel> chans;
    bool finalized = false;
    int device = 0;
    dpplan::Geometry geom;
    int v2_r1 = 0;  // != 0: the v2 kernels (dp_of2_kernel.cuh) serve this plan, M = v2_r1 * 4096
    int v2_multi = 0;  // some channel has more than one template
    int v2_narrow = 0;  // every template's delay windows span at most two pass-1' columns per thread (column-wise scan)
    int neighbours = 0;    // every fit also reports the amplitude one sample before / after its best delay (interpolate_t0)
    bool generic = false;  // nb_samples is not a power of two: the mixed-radix kernel (dp_ofg_kernel.cuh) serves this plan
    const void *g_tw = nullptr, *g_wn = nullptr, *g_pos_k = nullptr, *g_pos_m = nullptr;
    std::vector<int> g_radix;
    int g_pairs = 0;
    size_t persist_bytes = 0;  // scratch bytes kept persisting in L2 (0: no access-policy window)
    // device state
    std::vector<void*> owned;
    const void* d_chans = nullptr;
    const void *tw1 = nullptr, *tw2 = nullptr, *twn = nullptr, *twp = nullptr, *tw3 = nullptr, *groups = nullptr;
    const void *chunk3 = nullptr, *zones = nullptr;
    void* scratch = nullptr;
    long long scratch_per_cta = 0;
    int grid_max = 0;
    size_t smem = 0;
    int n_out = 0;
    std::vector<int> chan_out_base;
    int nlow = 0;
    double scale = 1.0;
    int subtract_first = 0;
    std::vector<double> adc_gain, adc_offset;  // per channel, int16 traces only
    long long* d_chan_off = nullptr;           // [n_chan] channel offsets of the last dp_of1x1_batch_ex layout
    std::vector<long long> chan_off_host;      // what d_chan_off holds
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long launches = 0;
    // staging for dp_of1x1_batch_host
    void* stage_dev[2] = {nullptr, nullptr};
    double* stage_out[2] = {nullptr, nullptr};
    double* stage_out_host = nullptr;  // pinned: the result rows of one call (a D2H copy into pageable memory blocks the
    long long stage_out_host_rows = 0; // host until the chunk's kernel is done and the next chunk's H2D would not overlap it)
    long long stage_events = 0;
    int stage_dtype = -1;
    long long stage_stride = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
};

namespace {
// the low-frequency tables (chi0 weights, template spectrum) reach up to the highest lowchi2_fcutoff any fit asks for
static double table_fcut(const dp_of_plan* p) {
    double f = -1.0;
    for (const auto& ch : p->chans)
        for (const auto& ft : ch.fits) f = std::max(f, ft.fcut < 0 ? p->fcut : ft.fcut);
    return f < 0 ? p->fcut : f;
}


template <class T> int of_setup(dp_of_plan* p) {
    const int prec = sizeof(T) == 8 ? 0 : 1;
    for (int in = 0; in < 3; ++in) {
        size_t smem = 0;
        int grid_max = 0, occ = 0;
        const int rc = dp_of_setup_table[prec][in](p->geom.R1, p->geom.P, p->device, &smem, &grid_max, &occ);
        if (rc == -1) return fail(DP_ERR_UNSUPPORTED, "unsupported trace length for this precision");
        if (rc != 0) return fail(DP_ERR_CUDA, std::string("OF kernel setup: ") + cudaGetErrorString((cudaError_t)rc));
        if (occ < 1) return fail(DP_ERR_CUDA, "OF kernel does not fit on an SM");
        p->smem = smem;
        p->grid_max = in == 0 ? grid_max : std::min(p->grid_max, grid_max);
    }
    return DP_OK;
}
template <class T> int of_launch(dp_of_plan* p, const DpOfParams<T>& prm, int grid, cudaStream_t st) {
    const int prec = sizeof(T) == 8 ? 0 : 1;
    const int rc = dp_of_launch_table[prec][prm.in_dtype](p->geom.R1, p->geom.P, &prm, grid, p->smem, st);
    if (rc == -1) return fail(DP_ERR_UNSUPPORTED, "unsupported trace length for this precision");
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("OF kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    return DP_OK;
}

template <class T> int of_finalize(dp_of_plan* p) {
    if (p->neighbours) return fail(DP_ERR_UNSUPPORTED, "interpolate_t0 needs nb_samples 16384, 32768, 65536 or a non power of two");
    dpplan::DeviceTables<T> dt;
    try {
        dt = dpplan::build_tables<T>(p->geom, p->fs, p->chans, table_fcut(p), p->scale);
    } catch (const std::exception& e) {
        return fail(DP_ERR_INVALID, e.what());
    }
    p->nlow = dt.nlow;
    int rc;
    const cx<T>* d;
    if ((rc = upload(p->owned, dt.tw1, &d))) return rc;
    p->tw1 = d;
    if ((rc = upload(p->owned, dt.tw2, &d))) return rc;
    p->tw2 = d;
    if ((rc = upload(p->owned, dt.twn, &d))) return rc;
    p->twn = d;
    if ((rc = upload(p->owned, dt.twp, &d))) return rc;
    p->twp = d;
    std::vector<DpChanDev<T>> cd(p->n_chan);
    p->chan_out_base.assign(p->n_chan, 0);
    int base = 0;
    for (int c = 0; c < p->n_chan; ++c) {
        DpChanDev<T>& dc = cd[c];
        std::memset(&dc, 0, sizeof(dc));
        const T* w;
        if ((rc = upload(p->owned, dt.chans[c].wj, &w))) return rc;
        dc.wj = w;
        if ((rc = upload(p->owned, dt.chans[c].wj_low, &w))) return rc;
        dc.wj_low = w;
        if ((rc = upload(p->owned, dt.chans[c].wj_self, &w))) return rc;
        dc.wj_self = w;
        dc.adc_gain = c < (int)p->adc_gain.size() ? p->adc_gain[c] : 1.0;
        dc.adc_offset = c < (int)p->adc_offset.size() ? p->adc_offset[c] : 0.0;
        dc.n_templ = (int)p->chans[c].templ.size();
        dc.n_slots = (int)p->chans[c].fits.size();
        dc.out_base = base;
        p->chan_out_base[c] = base;
        base += 1 + (DP_SLOT_NOUT + (p->neighbours ? 2 : 0)) * dc.n_slots;
        for (int i = 0; i < dc.n_templ; ++i) {
            auto& h = dt.chans[c].templ[i];
            const cx<T>* ph;
            if ((rc = upload(p->owned, h.phi, &ph))) return rc;
            dc.templ[i].phi = ph;
            if ((rc = upload(p->owned, h.s_low, &ph))) return rc;
            dc.templ[i].s_low = ph;
            if ((rc = upload(p->owned, h.phi_self, &ph))) return rc;
            dc.templ[i].phi_self = ph;
            dc.templ[i].norm = h.norm;
            dc.templ[i].tsum = h.tsum;
            dc.templ[i].pretrigger = h.pretrigger;
        }
        for (int i = 0; i < dc.n_slots; ++i) {
            const auto& f = p->chans[c].fits[i];
            dc.slots[i] = DpSlot{f.templ, f.lo, f.hi, f.outside, std::min(p->nlow, dpplan::count_low_bins(p->N, p->fs, f.fcut < 0 ? p->fcut : f.fcut))};
        }
    }
    p->n_out = base;
    const DpChanDev<T>* dcd;
    if ((rc = upload(p->owned, cd, &dcd))) return rc;
    p->d_chans = dcd;
    if ((rc = of_setup<T>(p))) return rc;
    p->scratch_per_cta = 96LL * p->geom.NT;
    DP_CUDA(cudaMalloc(&p->scratch, sizeof(cx<T>) * (size_t)p->scratch_per_cta * (size_t)p->grid_max));
    p->owned.push_back(p->scratch);
    return DP_OK;
}

// where the first sample of (event, channel) sits: see Dp2Params
struct OfLayout {
    long long event_stride = 0, chan_stride = 0;
    const long long* chan_offset_dev = nullptr;
    const long long* row_start = nullptr;
    long long stream_len = 0;
};
static OfLayout default_layout(int n_chan, long long row_stride) {
    OfLayout l;
    l.event_stride = (long long)n_chan * row_stride;
    l.chan_stride = row_stride;
    return l;
}

template <class T>
int of_run(dp_of_plan* p, const void* traces_dev, int in_dtype, long long n_events, const OfLayout& lay, double* out_dev,
           cudaStream_t st, bool timed) {
    if (lay.row_start != nullptr || in_dtype > DP_IN_I16)
        return fail(DP_ERR_UNSUPPORTED, "windows of a continuous stream need nb_samples 16384, 32768 or 65536");
    DpOfParams<T> prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.traces = traces_dev;
    prm.event_stride = lay.event_stride;
    prm.chan_stride = lay.chan_stride;
    prm.chan_offset = lay.chan_offset_dev;
    prm.n_rows = (int)(n_events * p->n_chan);
    prm.n_chan = p->n_chan;
    prm.chans = reinterpret_cast<const DpChanDev<T>*>(p->d_chans);
    prm.tw1 = reinterpret_cast<const cx<T>*>(p->tw1);
    prm.tw2 = reinterpret_cast<const cx<T>*>(p->tw2);
    prm.twn = reinterpret_cast<const cx<T>*>(p->twn);
    prm.twp = reinterpret_cast<const cx<T>*>(p->twp);
    prm.scratch = reinterpret_cast<cx<T>*>(p->scratch);
    prm.scratch_per_cta = p->scratch_per_cta;
    prm.out = out_dev;
    prm.n_out = p->n_out;
    prm.nlow = p->nlow;
    prm.scale = p->scale;
    prm.subtract_first = p->subtract_first;
    prm.in_dtype = in_dtype;
    const int grid = (int)std::min<long long>(prm.n_rows, p->grid_max);
    if (timed) DP_CUDA(cudaEventRecord(p->ev0, st));
    int rc = of_launch<T>(p, prm, grid, st);
    if (rc) return rc;
    if (timed) DP_CUDA(cudaEventRecord(p->ev1, st));
    p->timed = timed;
    p->launches += 1;
    return DP_OK;
}

// ------------------------------------------------------------------ v2 kernels
template <class T> int of2_finalize(dp_of_plan* p) {
    using S = typename Dp2Traits<T>::S;
    dpplan2::Tables2<T> dt;
    try {
        switch (p->v2_r1) {
            case 2: dt = dpplan2::build_tables2<T, 2>(p->fs, p->chans, table_fcut(p), p->scale); break;
            case 4: dt = dpplan2::build_tables2<T, 4>(p->fs, p->chans, table_fcut(p), p->scale); break;
            default: dt = dpplan2::build_tables2<T, 8>(p->fs, p->chans, table_fcut(p), p->scale); break;
        }
    } catch (const std::invalid_argument& e) {
        return fail(DP_ERR_INVALID, e.what());
    } catch (const std::exception& e) {
        return fail(DP_ERR_STATE, e.what());
    }
    p->nlow = dt.nlow;
    int rc;
    const cx<T>* d;
    if ((rc = upload(p->owned, dt.tw1, &d))) return rc;
    p->tw1 = d;
    if ((rc = upload(p->owned, dt.tw2, &d))) return rc;
    p->tw2 = d;
    if ((rc = upload(p->owned, dt.tw3, &d))) return rc;
    p->tw3 = d;
    const cx<S>* ds;
    if ((rc = upload(p->owned, dt.twn, &ds))) return rc;
    p->twn = ds;
    const int2* dg;
    if ((rc = upload(p->owned, dt.groups, &dg))) return rc;
    p->groups = dg;
    const int* dc3;
    if ((rc = upload(p->owned, dt.chunk3, &dc3))) return rc;
    p->chunk3 = dc3;
    const uint4* dzo;
    if ((rc = upload(p->owned, dt.zones, &dzo))) return rc;
    p->zones = dzo;
    std::vector<Dp2ChanDev<T>> cd(p->n_chan);
    p->chan_out_base.assign(p->n_chan, 0);
    int base = 0, max_templ = 1;
    for (int c = 0; c < p->n_chan; ++c) {
        Dp2ChanDev<T>& dc = cd[c];
        std::memset(&dc, 0, sizeof(dc));
        const T* w;
        const S* ws;
        if ((rc = upload(p->owned, dt.chans[c].wj, &w))) return rc;
        dc.wj = w;
        if ((rc = upload(p->owned, dt.chans[c].wj_low, &ws))) return rc;
        dc.wj_low = ws;
        if ((rc = upload(p->owned, dt.chans[c].wj_self, &ws))) return rc;
        dc.wj_self = ws;
        dc.adc_gain = c < (int)p->adc_gain.size() ? p->adc_gain[c] : 1.0;
        dc.adc_offset = c < (int)p->adc_offset.size() ? p->adc_offset[c] : 0.0;
        dc.n_templ = (int)p->chans[c].templ.size();
        dc.n_slots = (int)p->chans[c].fits.size();
        dc.out_base = base;
        p->chan_out_base[c] = base;
        base += 1 + (DP_SLOT_NOUT + (p->neighbours ? 2 : 0)) * dc.n_slots;
        max_templ = std::max(max_templ, dc.n_templ);
        for (int i = 0; i < dc.n_templ; ++i) {
            auto& h = dt.chans[c].templ[i];
            const cx<T>* ph;
            const cx<S>* phs;
            if ((rc = upload(p->owned, h.phi, &ph))) return rc;
            dc.templ[i].phi = ph;
            if ((rc = upload(p->owned, h.s_low, &phs))) return rc;
            dc.templ[i].s_low = phs;
            if ((rc = upload(p->owned, h.phi_self, &phs))) return rc;
            dc.templ[i].phi_self = phs;
            dc.templ[i].norm = h.norm;
            dc.templ[i].tsum = h.tsum;
            dc.templ[i].pretrigger = h.pretrigger;
        }
        for (int i = 0; i < dc.n_slots; ++i) {
            const auto& f = p->chans[c].fits[i];
            dc.slots[i] = DpSlot{f.templ, f.lo, f.hi, f.outside, std::min(p->nlow, dpplan::count_low_bins(p->N, p->fs, f.fcut < 0 ? p->fcut : f.fcut))};
        }
    }
    p->n_out = base;
    const Dp2ChanDev<T>* dcd;
    if ((rc = upload(p->owned, cd, &dcd))) return rc;
    p->d_chans = dcd;
    const int prec = sizeof(S) == 8 ? 0 : 1;
    for (int in = 0; in < 6; ++in) {
        size_t smem = 0;
        int grid_max = 0, occ = 0, threads = 0;
        const int src = dp_of2_setup_table[prec][in](p->v2_r1, p->device, &smem, &grid_max, &occ, &threads);
        if (src == -1) return fail(DP_ERR_UNSUPPORTED, "unsupported trace length");
        if (src != 0) return fail(DP_ERR_CUDA, std::string("OF kernel setup: ") + cudaGetErrorString((cudaError_t)src));
        if (occ < 1) return fail(DP_ERR_CUDA, "OF kernel does not fit on an SM");
        p->smem = smem;
        p->grid_max = in == 0 ? grid_max : std::min(p->grid_max, grid_max);
    }
    long long per_cta = 0;
    switch (p->v2_r1) {
        case 2: per_cta = Dp2OfKernel<T, 2, 0>::scratch_v(max_templ); break;
        case 4: per_cta = Dp2OfKernel<T, 4, 0>::scratch_v(max_templ); break;
        default: per_cta = Dp2OfKernel<T, 8, 0>::scratch_v(max_templ); break;
    }
    p->scratch_per_cta = per_cta;
    p->v2_multi = max_templ > 1 ? 1 : 0;
    {
        // mirrors the kernel's window union per template (dp_of2_kernel.cuh, "fits of this template"): complex points
        // [nlo, nhi] that some fit can select, widened by one point when the neighbour amplitudes are reported
        const int M = p->N / 2, nb = std::min(p->v2_r1, std::is_same<T, double>::value ? 2 : DP2_NBMAX_F32);   // blocks per phase (Dp2Geom::NB)
        const int bound = 2 * nb * 256 - 1;                                         // 2 * NT * VL - 1
        bool narrow = true, any = false;
        for (int c = 0; c < p->n_chan; ++c)
            for (int t = 0; t < (int)p->chans[c].templ.size(); ++t) {
                int nlo = 0x7fffffff, nhi = -1;
                for (const auto& f : p->chans[c].fits) {
                    if (f.templ != t) continue;
                    const bool everything = f.outside || (f.lo == 0 && f.hi == p->N);
                    const int a = everything ? 0 : (f.lo >> 1), b = everything ? 0x7fffffff : ((f.hi - 1) >> 1);
                    nlo = std::min(nlo, a);
                    nhi = std::max(nhi, b);
                }
                if (nhi < 0) continue;       // template without fits
                any = true;
                if (p->neighbours && nhi >= nlo && nhi - nlo < 4095) {
                    nlo = nlo > 0 ? nlo - 1 : 0;
                    nhi = nhi < M - 1 ? nhi + 1 : M - 1;
                }
                narrow = narrow && nhi >= nlo && (nhi - nlo) < bound;
            }
        p->v2_narrow = (any && narrow) ? 1 : 0;
    }
    const size_t scratch_bytes = sizeof(cx<T>) * (size_t)per_cta * (size_t)p->grid_max;
    DP_CUDA(cudaMalloc(&p->scratch, scratch_bytes));
    p->owned.push_back(p->scratch);
    // L2 residency of the scratch, opt-in with DP_L2_PERSIST=1 (it changes a device-wide limit): reserve a persisting
    // set-aside and tag the scratch window on every launch; measured +2 % on the fp64 two-template plan
    p->persist_bytes = 0;
    {
        const char* e = std::getenv("DP_L2_PERSIST");
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, p->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, p->device);
        const bool uses_scratch = p->v2_multi || scratch_bytes > sizeof(cx<T>) * 16 * 512 * (size_t)p->grid_max;
        if (e && std::string(e) == "1" && uses_scratch && max_persist > 0 && scratch_bytes <= (size_t)max_window) {
            size_t cur = 0;
            cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
            const size_t want = std::min(scratch_bytes, (size_t)max_persist);
            if (cur >= want || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) p->persist_bytes = scratch_bytes;
            cudaGetLastError();
        }
    }
    return DP_OK;
}

template <class T>
int of2_run(dp_of_plan* p, const void* traces_dev, int in_dtype, long long n_events, const OfLayout& lay, double* out_dev,
            cudaStream_t st, bool timed) {
    using S = typename Dp2Traits<T>::S;
    Dp2Params<T> prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.traces = traces_dev;
    prm.event_stride = lay.event_stride;
    prm.chan_stride = lay.chan_stride;
    prm.chan_offset = lay.chan_offset_dev;
    prm.n_rows = (int)(n_events * p->n_chan);
    prm.n_chan = p->n_chan;
    prm.chans = reinterpret_cast<const Dp2ChanDev<T>*>(p->d_chans);
    prm.tw1 = reinterpret_cast<const cx<T>*>(p->tw1);
    prm.tw2 = reinterpret_cast<const cx<T>*>(p->tw2);
    prm.tw3 = reinterpret_cast<const cx<T>*>(p->tw3);
    prm.twn = reinterpret_cast<const cx<S>*>(p->twn);
    prm.groups = reinterpret_cast<const int2*>(p->groups);
    prm.chunk3 = reinterpret_cast<const int*>(p->chunk3);
    prm.zones = reinterpret_cast<const uint4*>(p->zones);
    prm.scratch = reinterpret_cast<cx<T>*>(p->scratch);
    prm.scratch_per_cta = p->scratch_per_cta;
    prm.out = out_dev;
    prm.n_out = p->n_out;
    prm.nlow = p->nlow;
    prm.scale = p->scale;
    prm.subtract_first = p->subtract_first;
    {
        const char* e = std::getenv("DP2_SKEW_NS");  // development switch
        prm.skew_ns = e ? std::atoi(e) : DP2_SKEW_NS;
    }
    prm.row_start = lay.row_start;
    prm.stream_len = lay.stream_len;
    prm.neighbours = p->neighbours;
    const int grid = (int)std::min<long long>(prm.n_rows, p->grid_max);
    if (timed) DP_CUDA(cudaEventRecord(p->ev0, st));
    const int prec = sizeof(S) == 8 ? 0 : 1;
    const int rc = dp_of2_launch_table[prec][in_dtype](p->v2_r1, p->v2_multi | (p->v2_narrow << 1), &prm, grid, p->smem, st, p->persist_bytes);
    if (rc == -1) return fail(DP_ERR_UNSUPPORTED, "unsupported trace length");
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("OF kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    if (timed) DP_CUDA(cudaEventRecord(p->ev1, st));
    p->timed = timed;
    p->launches += 1;
    return DP_OK;
}

// ------------------------------------------------------------------ mixed-radix kernel (nb_samples not 2^k)
template <class T> int ofg_finalize(dp_of_plan* p) {
    dpgen::Tables<T> dt;
    try {
        dt = dpgen::build_tables<T>(p->N, p->fs, p->chans, table_fcut(p), p->scale);
    } catch (const std::invalid_argument& e) {
        return fail(DP_ERR_INVALID, e.what());
    } catch (const std::exception& e) {
        return fail(DP_ERR_STATE, e.what());
    }
    p->nlow = dt.nlow;
    p->g_radix = dt.radix;
    p->g_pairs = dt.n_pairs;
    int rc;
    const cx<T>* d;
    const int* di;
    if ((rc = upload(p->owned, dt.tw, &d))) return rc;
    p->g_tw = d;
    if ((rc = upload(p->owned, dt.wn, &d))) return rc;
    p->g_wn = d;
    if ((rc = upload(p->owned, dt.pos_k, &di))) return rc;
    p->g_pos_k = di;
    if ((rc = upload(p->owned, dt.pos_m, &di))) return rc;
    p->g_pos_m = di;
    std::vector<DpGenChanDev<T>> cd(p->n_chan);
    p->chan_out_base.assign(p->n_chan, 0);
    int base = 0;
    for (int c = 0; c < p->n_chan; ++c) {
        DpGenChanDev<T>& dc = cd[c];
        std::memset(&dc, 0, sizeof(dc));
        const T* w;
        if ((rc = upload(p->owned, dt.chans[c].wj_k, &w))) return rc;
        dc.wj_k = w;
        if ((rc = upload(p->owned, dt.chans[c].wj_m, &w))) return rc;
        dc.wj_m = w;
        if ((rc = upload(p->owned, dt.chans[c].wj_low, &w))) return rc;
        dc.wj_low = w;
        dc.adc_gain = c < (int)p->adc_gain.size() ? p->adc_gain[c] : 1.0;
        dc.adc_offset = c < (int)p->adc_offset.size() ? p->adc_offset[c] : 0.0;
        dc.n_templ = (int)p->chans[c].templ.size();
        dc.n_slots = (int)p->chans[c].fits.size();
        dc.out_base = base;
        p->chan_out_base[c] = base;
        base += 1 + (DP_SLOT_NOUT + (p->neighbours ? 2 : 0)) * dc.n_slots;
        for (int i = 0; i < dc.n_templ; ++i) {
            auto& h = dt.chans[c].templ[i];
            const cx<T>* ph;
            if ((rc = upload(p->owned, h.phi_k, &ph))) return rc;
            dc.templ[i].phi_k = ph;
            if ((rc = upload(p->owned, h.phi_m, &ph))) return rc;
            dc.templ[i].phi_m = ph;
            if ((rc = upload(p->owned, h.s_low, &ph))) return rc;
            dc.templ[i].s_low = ph;
            dc.templ[i].norm = h.norm;
            dc.templ[i].tsum = h.tsum;
            dc.templ[i].pretrigger = h.pretrigger;
        }
        for (int i = 0; i < dc.n_slots; ++i) {
            const auto& f = p->chans[c].fits[i];
            dc.slots[i] = DpSlot{f.templ, f.lo, f.hi, f.outside, std::min(p->nlow, dpplan::count_low_bins(p->N, p->fs, f.fcut < 0 ? p->fcut : f.fcut))};
        }
    }
    p->n_out = base;
    const DpGenChanDev<T>* dcd;
    if ((rc = upload(p->owned, cd, &dcd))) return rc;
    p->d_chans = dcd;
    size_t smem = 0;
    int grid_max = 0;
    const int src = dp_ofg_setup(sizeof(T) == 8 ? 0 : 1, p->N / 2, p->device, &smem, &grid_max);
    if (src == -2) return fail(DP_ERR_UNSUPPORTED, "event does not fit one SM's shared memory");
    if (src != 0) return fail(DP_ERR_CUDA, std::string("OF kernel setup: ") + cudaGetErrorString((cudaError_t)src));
    p->smem = smem;
    p->grid_max = grid_max;
    return DP_OK;
}

template <class T>
int ofg_run(dp_of_plan* p, const void* traces_dev, int in_dtype, long long n_events, const OfLayout& lay, double* out_dev,
            cudaStream_t st, bool timed) {
    DpGenParams<T> prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.traces = traces_dev;
    prm.in_dtype = in_dtype % 3;        // element-wise loads: the pair-aligned / element-aligned variants coincide
    prm.event_stride = lay.event_stride;
    prm.chan_stride = lay.chan_stride;
    prm.chan_offset = lay.chan_offset_dev;
    prm.row_start = lay.row_start;
    prm.stream_len = lay.stream_len;
    prm.n_rows = (int)(n_events * p->n_chan);
    prm.n_chan = p->n_chan;
    prm.chans = reinterpret_cast<const DpGenChanDev<T>*>(p->d_chans);
    prm.M = p->N / 2;
    prm.n_pass = (int)p->g_radix.size();
    for (int j = 0; j < prm.n_pass; ++j) prm.radix[j] = p->g_radix[j];
    prm.tw = reinterpret_cast<const cx<T>*>(p->g_tw);
    prm.wn = reinterpret_cast<const cx<T>*>(p->g_wn);
    prm.pos_k = reinterpret_cast<const int*>(p->g_pos_k);
    prm.pos_m = reinterpret_cast<const int*>(p->g_pos_m);
    prm.n_pairs = p->g_pairs;
    prm.out = out_dev;
    prm.n_out = p->n_out;
    prm.nlow = p->nlow;
    prm.scale = p->scale;
    prm.subtract_first = p->subtract_first;
    prm.neighbours = p->neighbours;
    const int grid = (int)std::min<long long>(prm.n_rows, p->grid_max);
    if (timed) DP_CUDA(cudaEventRecord(p->ev0, st));
    const int rc = dp_ofg_launch(sizeof(T) == 8 ? 0 : 1, &prm, grid, p->smem, st);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("OF kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    if (timed) DP_CUDA(cudaEventRecord(p->ev1, st));
    p->timed = timed;
    p->launches += 1;
    return DP_OK;
}

// precision / kernel-generation dispatch of one batch launch
int of_dispatch(dp_of_plan* p, const void* traces_dev, int in_dtype, long long n_events, const OfLayout& lay, double* out_dev,
                cudaStream_t st, bool timed) {
    if (p->v2_r1) {
        if (p->precision == DP_PREC_F32) return of2_run<f2>(p, traces_dev, in_dtype, n_events, lay, out_dev, st, timed);
        return of2_run<double>(p, traces_dev, in_dtype, n_events, lay, out_dev, st, timed);
    }
    if (p->generic) {
        if (p->precision == DP_PREC_F32) return ofg_run<float>(p, traces_dev, in_dtype, n_events, lay, out_dev, st, timed);
        return ofg_run<double>(p, traces_dev, in_dtype, n_events, lay, out_dev, st, timed);
    }
    if (p->precision == DP_PREC_F32) return of_run<float>(p, traces_dev, in_dtype, n_events, lay, out_dev, st, timed);
    return of_run<double>(p, traces_dev, in_dtype, n_events, lay, out_dev, st, timed);
}

int of_check(const dp_of_plan* p, int chan) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (chan < 0 || chan >= p->n_chan) return fail(DP_ERR_INVALID, "channel index out of range");
    return DP_OK;
}

}  // namespace

extern "C" {

const char* dp_last_error(void) { return g_err.c_str(); }
int dp_version(void) { return 100; }
int dp_device_count(int* count) {
    DP_CUDA(cudaGetDeviceCount(count));
    return DP_OK;
}

int dp_of_plan_create(dp_of_plan** plan, int nb_samples, double sample_rate, int n_chan, int precision) {
    if (!plan) return fail(DP_ERR_INVALID, "null plan pointer");
    if (n_chan < 1) return fail(DP_ERR_INVALID, "n_chan must be >= 1");
    if (!(sample_rate > 0)) return fail(DP_ERR_INVALID, "sample_rate must be > 0");
    if (precision != DP_PREC_F64 && precision != DP_PREC_F32) return fail(DP_ERR_INVALID, "unknown precision");
    auto p = std::make_unique<dp_of_plan>();
    // nb_samples 16384 / 32768 / 65536 run on the v2 kernels (DP_OF_KERNEL=v1 selects the first
    // generation where it supports the length; kept for A/B measurements)
    const char* gen = std::getenv("DP_OF_KERNEL");
    p->v2_r1 = (gen && std::string(gen) == "v1") ? 0 : dpplan2::r1_of(nb_samples);
    if (!p->v2_r1) {
        if (!dpplan::is_pow2(nb_samples) && dpgen::supported(nb_samples, precision == DP_PREC_F64)) {
            p->generic = true;      // e.g. 25000 / 12500 samples (reference examples/processing/process_example.yaml:93-94)
        } else {
            try {
                p->geom = dpplan::pick_geometry(nb_samples, precision == DP_PREC_F64);
            } catch (const std::exception& e) {
                return fail(DP_ERR_UNSUPPORTED, std::string(e.what()) + " (or an even length whose half factors into 2, 3, 4, 5 and fits one SM)");
            }
        }
    }
    p->N = nb_samples;
    p->fs = sample_rate;
    p->n_chan = n_chan;
    p->precision = precision;
    p->chans.resize(n_chan);
    *plan = p.release();
    return DP_OK;
}

void dp_of_plan_destroy(dp_of_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    for (int i = 0; i < 2; ++i) {
        if (p->stage_dev[i]) cudaFree(p->stage_dev[i]);
        if (p->stage_out[i]) cudaFree(p->stage_out[i]);
        if (p->streams[i]) cudaStreamDestroy(p->streams[i]);
    }
    if (p->stage_out_host) cudaFreeHost(p->stage_out_host);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}

int dp_of_plan_set_psd(dp_of_plan* p, int chan, const double* psd, int coupling_ac) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    if (!psd) return fail(DP_ERR_INVALID, "null psd");
    auto& ch = p->chans[chan];
    if (!ch.templ.empty()) return fail(DP_ERR_STATE, "set the psd before adding templates");
    ch.J.assign(psd, psd + p->N);
    for (int i = 0; i < p->N; ++i)
        if (!(ch.J[i] > 0)) return fail(DP_ERR_INVALID, "psd must be strictly positive");
    if (coupling_ac) ch.J[0] = std::numeric_limits<double>::infinity();
    return DP_OK;
}

int dp_of_plan_add_template(dp_of_plan* p, int chan, const double* templ, int pretrigger_samples, int integralnorm,
                            int* templ_index) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    if (!templ) return fail(DP_ERR_INVALID, "null template");
    auto& ch = p->chans[chan];
    if ((int)ch.J.size() != p->N) return fail(DP_ERR_STATE, "set the psd before adding templates");
    if ((int)ch.templ.size() >= DP_MAX_TEMPLATES) return fail(DP_ERR_UNSUPPORTED, "too many templates for one channel");
    if (pretrigger_samples < 0 || pretrigger_samples >= p->N) return fail(DP_ERR_INVALID, "pretrigger_samples out of range");
    dpplan::Template tp;
    tp.trace.assign(templ, templ + p->N);
    tp.pretrigger = pretrigger_samples;
    tp.integralnorm = integralnorm != 0;
    dpplan::finalize_template(tp, ch.J, p->fs);
    if (!(tp.norm > 0)) return fail(DP_ERR_INVALID, "template has zero optimal-filter norm");
    ch.templ.push_back(std::move(tp));
    if (templ_index) *templ_index = (int)ch.templ.size() - 1;
    return DP_OK;
}

int dp_of_plan_add_fit_ex(dp_of_plan* p, int chan, int templ_index, int window_lo, int window_hi, int outside,
                          double lowchi2_fcutoff_hz, int* fit_index) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    auto& ch = p->chans[chan];
    if (templ_index < 0 || templ_index >= (int)ch.templ.size()) return fail(DP_ERR_INVALID, "template index out of range");
    if ((int)ch.fits.size() >= DP_MAX_SLOTS) return fail(DP_ERR_UNSUPPORTED, "too many fits for one channel");
    {
        int nt = 0;
        for (const auto& f : ch.fits) nt += f.templ == templ_index;
        if (nt >= DP_MAX_TSLOTS) return fail(DP_ERR_UNSUPPORTED, "too many fits for one template");
    }
    if (lowchi2_fcutoff_hz >= 0 && !std::isfinite(lowchi2_fcutoff_hz)) return fail(DP_ERR_INVALID, "lowchi2_fcutoff must be finite");
    window_lo = std::min(std::max(window_lo, 0), p->N);
    window_hi = std::min(std::max(window_hi, 0), p->N);
    const int ncand = outside ? p->N - std::max(0, window_hi - window_lo) : window_hi - window_lo;
    if (ncand <= 0) return fail(DP_ERR_INVALID, "empty OF delay window");
    dpplan::Fit f{templ_index, window_lo, window_hi, outside ? 1 : 0};
    f.fcut = lowchi2_fcutoff_hz;
    ch.fits.push_back(f);
    if (fit_index) *fit_index = (int)ch.fits.size() - 1;
    return DP_OK;
}
int dp_of_plan_add_fit(dp_of_plan* p, int chan, int templ_index, int window_lo, int window_hi, int outside, int* fit_index) {
    return dp_of_plan_add_fit_ex(p, chan, templ_index, window_lo, window_hi, outside, -1.0, fit_index);
}

int dp_of_plan_set_lowchi2_fcutoff(dp_of_plan* p, double fcutoff_hz) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    p->fcut = fcutoff_hz;
    return DP_OK;
}

int dp_of_plan_set_adc_conversion(dp_of_plan* p, int chan, double gain, double offset) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    if (!(gain > 0) || !std::isfinite(gain) || !std::isfinite(offset)) return fail(DP_ERR_INVALID, "adc gain must be finite and > 0");
    p->adc_gain.resize(p->n_chan, 1.0);
    p->adc_offset.resize(p->n_chan, 0.0);
    p->adc_gain[chan] = gain;
    p->adc_offset[chan] = offset;
    return DP_OK;
}

int dp_of_plan_finalize(dp_of_plan* p, int device) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    bool all_ac = true;
    double jsum = 0;
    long long jn = 0;
    for (auto& ch : p->chans) {
        if ((int)ch.J.size() != p->N) return fail(DP_ERR_STATE, "a channel has no psd");
        if (ch.templ.empty()) return fail(DP_ERR_STATE, "a channel has no template");
        if (std::isfinite(ch.J[0])) all_ac = false;
        for (int i = 1; i < p->N; ++i)
            if (std::isfinite(ch.J[i])) {
                jsum += ch.J[i];
                ++jn;
            }
    }
    p->device = device;
    DP_ON_DEVICE(device);
    if (p->precision == DP_PREC_F32) {
        // power-of-two pre-scale so fp32 sees O(1) samples; exact, undone in the tables
        const double rms = std::sqrt(std::max(jsum / std::max<long long>(jn, 1) * p->fs, 1e-300));
        p->scale = std::exp2(-std::round(std::log2(rms)));
        p->subtract_first = all_ac ? 1 : 0;
    } else {
        p->scale = 1.0;
        p->subtract_first = 0;
    }
    int rc;
    if (p->v2_r1)
        rc = (p->precision == DP_PREC_F32) ? of2_finalize<f2>(p) : of2_finalize<double>(p);
    else if (p->generic)
        rc = (p->precision == DP_PREC_F32) ? ofg_finalize<float>(p) : ofg_finalize<double>(p);
    else
        rc = (p->precision == DP_PREC_F32) ? of_finalize<float>(p) : of_finalize<double>(p);
    if (rc) return rc;
    DP_CUDA(cudaEventCreate(&p->ev0));
    DP_CUDA(cudaEventCreate(&p->ev1));
    p->finalized = true;
    return DP_OK;
}

int dp_of_plan_n_out(const dp_of_plan* p, int* n_out) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    *n_out = p->n_out;
    return DP_OK;
}
int dp_of_plan_chi0_offset(const dp_of_plan* p, int chan, int* offset) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (!p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    *offset = p->chan_out_base[chan];
    return DP_OK;
}
int dp_of_plan_fit_offset(const dp_of_plan* p, int chan, int fit_index, int* offset) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (!p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    if (fit_index < 0 || fit_index >= (int)p->chans[chan].fits.size()) return fail(DP_ERR_INVALID, "fit index out of range");
    *offset = p->chan_out_base[chan] + 1 + DP_SLOT_NOUT * fit_index;
    return DP_OK;
}
int dp_of_plan_set_neighbours(dp_of_plan* p, int on) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    p->neighbours = on ? 1 : 0;
    return DP_OK;
}
int dp_of_plan_neighbour_offset(const dp_of_plan* p, int chan, int fit_index, int* offset) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (!p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    if (!p->neighbours) return fail(DP_ERR_STATE, "the plan does not report neighbour amplitudes (dp_of_plan_set_neighbours)");
    const int nf = (int)p->chans[chan].fits.size();
    if (fit_index < 0 || fit_index >= nf) return fail(DP_ERR_INVALID, "fit index out of range");
    *offset = p->chan_out_base[chan] + 1 + DP_SLOT_NOUT * nf + 2 * fit_index;
    return DP_OK;
}
int dp_of_plan_get_phi(const dp_of_plan* p, int chan, int ti, double* out) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (ti < 0 || ti >= (int)p->chans[chan].templ.size()) return fail(DP_ERR_INVALID, "template index out of range");
    const auto& phi = p->chans[chan].templ[ti].phi;
    for (int i = 0; i < p->N; ++i) {
        out[2 * i] = phi[i].real();
        out[2 * i + 1] = phi[i].imag();
    }
    return DP_OK;
}
int dp_of_plan_get_template_fft(const dp_of_plan* p, int chan, int ti, double* out) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (ti < 0 || ti >= (int)p->chans[chan].templ.size()) return fail(DP_ERR_INVALID, "template index out of range");
    const auto& s = p->chans[chan].templ[ti].s;
    for (int i = 0; i < p->N; ++i) {
        out[2 * i] = s[i].real();
        out[2 * i + 1] = s[i].imag();
    }
    return DP_OK;
}
int dp_of_plan_get_norm(const dp_of_plan* p, int chan, int ti, double* norm) {
    int rc = of_check(p, chan);
    if (rc) return rc;
    if (ti < 0 || ti >= (int)p->chans[chan].templ.size()) return fail(DP_ERR_INVALID, "template index out of range");
    *norm = p->chans[chan].templ[ti].norm;
    return DP_OK;
}

int dp_of1x1_batch(dp_of_plan* p, const void* traces_dev, int in_dtype, long long n_events, long long row_stride,
                   double* out_dev, void* stream) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    DP_ON_DEVICE(p->device);
    if (n_events < 0) return fail(DP_ERR_INVALID, "negative n_events");
    if (n_events == 0) return DP_OK;
    if (!traces_dev || !out_dev) return fail(DP_ERR_INVALID, "null buffer");
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (row_stride < p->N || (row_stride & 1)) return fail(DP_ERR_INVALID, "row_stride must be even and >= nb_samples");
    const size_t esz = in_dtype == DP_IN_F64 ? 8 : (in_dtype == DP_IN_F32 ? 4 : 2);
    if ((reinterpret_cast<uintptr_t>(traces_dev) % (2 * esz)) != 0) return fail(DP_ERR_INVALID, "trace buffer misaligned");
    if (n_events * p->n_chan > 2000000000LL) return fail(DP_ERR_INVALID, "batch too large; split it");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return of_dispatch(p, traces_dev, in_dtype, n_events, default_layout(p->n_chan, row_stride), out_dev, st, true);
}

int dp_of1x1_batch_ex(dp_of_plan* p, const void* base_dev, int in_dtype, long long n_events, long long event_stride,
                      const long long* chan_offsets, long long chan_stride, const long long* start_index_dev,
                      long long n_stream_samples, double* out_dev, void* stream) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    DP_ON_DEVICE(p->device);
    if (n_events < 0 || n_stream_samples < 0) return fail(DP_ERR_INVALID, "negative size");
    if (n_events == 0) return DP_OK;
    if (!base_dev || !out_dev) return fail(DP_ERR_INVALID, "null buffer");
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (n_events * p->n_chan > 2000000000LL) return fail(DP_ERR_INVALID, "batch too large; split it");
    const size_t esz = in_dtype == DP_IN_F64 ? 8 : (in_dtype == DP_IN_F32 ? 4 : 2);
    if ((reinterpret_cast<uintptr_t>(base_dev) % esz) != 0) return fail(DP_ERR_INVALID, "trace buffer misaligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    OfLayout lay;
    lay.event_stride = event_stride;
    lay.chan_stride = chan_stride;
    lay.row_start = start_index_dev;
    lay.stream_len = n_stream_samples;
    // rows aligned to a sample pair take the vector loads; anything else (windows, odd strides / offsets) the
    // element-aligned instantiation (in_dtype + 3)
    bool pair_aligned = start_index_dev == nullptr && (reinterpret_cast<uintptr_t>(base_dev) % (2 * esz)) == 0 &&
                        (event_stride % 2) == 0 && (chan_stride % 2) == 0;
    if (start_index_dev == nullptr) {
        if (event_stride < 0) return fail(DP_ERR_INVALID, "negative event_stride");
    } else if (n_stream_samples < p->N) {
        // every window leaves the stream: the kernel writes the sentinels
    }
    if (chan_offsets != nullptr) {
        for (int c = 0; c < p->n_chan; ++c) {
            if (chan_offsets[c] < 0) return fail(DP_ERR_INVALID, "negative channel offset");
            pair_aligned = pair_aligned && (chan_offsets[c] % 2) == 0;
        }
        // the offsets live in a small device array owned by the plan, refreshed only when they change
        if (p->d_chan_off == nullptr) {
            DP_CUDA(cudaMalloc(&p->d_chan_off, sizeof(long long) * (size_t)p->n_chan));
            p->owned.push_back(p->d_chan_off);
        }
        if (p->chan_off_host.size() != (size_t)p->n_chan || !std::equal(p->chan_off_host.begin(), p->chan_off_host.end(), chan_offsets)) {
            // (synchronous: a layout change is a configuration step, not part of the steady state)
            DP_CUDA(cudaStreamSynchronize(st));
            DP_CUDA(cudaMemcpy(p->d_chan_off, chan_offsets, sizeof(long long) * (size_t)p->n_chan, cudaMemcpyHostToDevice));
            p->chan_off_host.assign(chan_offsets, chan_offsets + p->n_chan);
        }
        lay.chan_offset_dev = p->d_chan_off;
    }
    const int in = pair_aligned ? in_dtype : in_dtype + 3;
    if (!pair_aligned && !p->v2_r1 && !p->generic)
        return fail(DP_ERR_UNSUPPORTED, "element-aligned rows / windows need nb_samples 16384, 32768, 65536 or a non power of two");
    return of_dispatch(p, base_dev, in, n_events, lay, out_dev, st, true);
}

int dp_of1x1_windows(dp_of_plan* p, const double* stream_dev, long long n_stream_samples, const long long* start_index_dev,
                     long long n_events, double* out_dev, void* stream) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    if (!p->v2_r1) return fail(DP_ERR_UNSUPPORTED, "window mode needs nb_samples 16384, 32768 or 65536");
    if (p->n_chan != 1) return fail(DP_ERR_UNSUPPORTED, "dp_of1x1_windows is single channel; dp_of1x1_batch_ex takes n_chan streams");
    if (!start_index_dev) return fail(DP_ERR_INVALID, "null buffer");
    return dp_of1x1_batch_ex(p, stream_dev, DP_IN_F64, n_events, 0, nullptr, n_stream_samples, start_index_dev, n_stream_samples,
                             out_dev, stream);
}

int dp_of_plan_last_kernel_ms(dp_of_plan* p, float* ms) {
    if (!p || !p->finalized || !p->timed) return fail(DP_ERR_STATE, "no timed launch");
    DP_CUDA(cudaEventSynchronize(p->ev1));
    DP_CUDA(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return DP_OK;
}
int dp_of_plan_launch_count(const dp_of_plan* p, long long* n) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    *n = p->launches;
    return DP_OK;
}

int dp_of1x1_batch_host(dp_of_plan* p, const void* traces_host, int in_dtype, long long n_events, long long row_stride,
                        double* out_host) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    if (n_events <= 0) return n_events == 0 ? DP_OK : fail(DP_ERR_INVALID, "negative n_events");
    if (!traces_host || !out_host) return fail(DP_ERR_INVALID, "null buffer");
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (row_stride < p->N || (row_stride & 1)) return fail(DP_ERR_INVALID, "row_stride must be even and >= nb_samples");
    DP_ON_DEVICE(p->device);
    const size_t esz = in_dtype == DP_IN_F64 ? 8 : (in_dtype == DP_IN_F32 ? 4 : 2);
    const size_t ev_bytes = (size_t)p->n_chan * (size_t)row_stride * esz;
    // chunk: <= 256 MiB of traces per stage and about eight stages per call (the copy of one overlaps the kernel of
    // the previous one; the first copy and the last kernel are exposed), a multiple of the persistent grid with at
    // least four events per CTA
    long long chunk = std::max<long long>(1, (256LL << 20) / (long long)ev_bytes);
    {
        const long long g = std::max(1, p->grid_max);
        int stages = 8;
        if (const char* e = std::getenv("DP_HOST_STAGES")) stages = std::max(1, std::atoi(e));  // development switch
        long long want = std::max<long long>(4 * g, (n_events + stages - 1) / stages);
        want = (want + g - 1) / g * g;
        chunk = std::min(chunk, want);
    }
    chunk = std::min(chunk, n_events);
    if (p->stage_events < chunk || p->stage_dtype != in_dtype || p->stage_stride != row_stride) {
        for (int i = 0; i < 2; ++i) {
            if (p->stage_dev[i]) cudaFree(p->stage_dev[i]);
            if (p->stage_out[i]) cudaFree(p->stage_out[i]);
            p->stage_dev[i] = nullptr;
            p->stage_out[i] = nullptr;
            DP_CUDA(cudaMalloc(&p->stage_dev[i], ev_bytes * (size_t)chunk));
            DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->stage_out[i]), sizeof(double) * (size_t)p->n_out * (size_t)chunk));
            if (!p->streams[i]) DP_CUDA(cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking));
        }
        p->stage_events = chunk;
        p->stage_dtype = in_dtype;
        p->stage_stride = row_stride;
    }
    if (p->stage_out_host_rows < n_events) {
        if (p->stage_out_host) cudaFreeHost(p->stage_out_host);
        p->stage_out_host = nullptr;
        p->stage_out_host_rows = 0;
        DP_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p->stage_out_host), sizeof(double) * (size_t)p->n_out * (size_t)n_events));
        p->stage_out_host_rows = n_events;
    }
    const unsigned char* src = reinterpret_cast<const unsigned char*>(traces_host);
    int k = 0;
    for (long long e0 = 0; e0 < n_events; e0 += chunk, k ^= 1) {
        const long long ne = std::min(chunk, n_events - e0);
        cudaStream_t st = p->streams[k];
        DP_CUDA(cudaMemcpyAsync(p->stage_dev[k], src + (size_t)e0 * ev_bytes, ev_bytes * (size_t)ne, cudaMemcpyHostToDevice, st));
        int rc = of_dispatch(p, p->stage_dev[k], in_dtype, ne, default_layout(p->n_chan, row_stride), p->stage_out[k], st, false);
        if (rc) return rc;
        DP_CUDA(cudaMemcpyAsync(p->stage_out_host + (size_t)e0 * p->n_out, p->stage_out[k], sizeof(double) * (size_t)p->n_out * (size_t)ne,
                                cudaMemcpyDeviceToHost, st));
    }
    DP_CUDA(cudaStreamSynchronize(p->streams[0]));
    DP_CUDA(cudaStreamSynchronize(p->streams[1]));
    std::memcpy(out_host, p->stage_out_host, sizeof(double) * (size_t)p->n_out * (size_t)n_events);
    return DP_OK;
}

}  // extern "C"

// ======================================================================= reduce plan
struct dp_reduce_plan {
    dpred::Plan plan;
    int n_chan = 0;
    bool finalized = false;
    int device = 0;
    std::vector<void*> owned;
    const DpRedChan* d_chans = nullptr;
    const DpLeaf* d_leaves = nullptr;
    const DpNode* d_nodes = nullptr;
    const int* d_level_off = nullptr;
    const DpRedFeat* d_feats = nullptr;
    std::vector<double> adc;  // [n_chan][2] gain, offset of int16 traces
    const double* d_adc = nullptr;
    long long* d_chan_off = nullptr;       // [n_chan] channel offsets of the last dp_window_reduce_batch_ex layout
    std::vector<long long> chan_off_host;
    int grid_max = 0;
    size_t smem = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

extern "C" {

int dp_channel_combine(const void* base_dev, int in_dtype, long long n_events, long long event_stride, int nb_samples, int n_out,
                       const int* n_terms, const long long* offsets, const double* weights, const int* weighted,
                       const double* adc_gain, const double* adc_offset, double* out_dev, void* stream) {
    if (!base_dev || !out_dev || !n_terms || !offsets || !weights) return fail(DP_ERR_INVALID, "null buffer");
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (n_events < 0 || nb_samples <= 0 || event_stride < 0) return fail(DP_ERR_INVALID, "negative size");
    if (n_out < 1 || n_out > DP_COMBINE_MAX_OUT) return fail(DP_ERR_UNSUPPORTED, "1..8 combined channels per call");
    if (n_events == 0) return DP_OK;
    DpCombineParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.base = base_dev;
    prm.event_stride = event_stride;
    prm.n_events = n_events;
    prm.nb_samples = nb_samples;
    prm.n_out = n_out;
    prm.out = out_dev;
    for (int j = 0; j < n_out; ++j) {
        if (n_terms[j] < 1 || n_terms[j] > DP_COMBINE_MAX_TERMS) return fail(DP_ERR_UNSUPPORTED, "1..4 terms per combined channel");
        prm.n_terms[j] = n_terms[j];
        prm.weighted[j] = weighted ? weighted[j] : 1;
        for (int t = 0; t < n_terms[j]; ++t) {
            const int k = j * DP_COMBINE_MAX_TERMS + t;
            if (offsets[k] < 0) return fail(DP_ERR_INVALID, "negative input offset");
            prm.off[j][t] = offsets[k];
            prm.w[j][t] = weights[k];
            prm.gain[j][t] = adc_gain ? adc_gain[k] : 1.0;
            prm.offs[j][t] = adc_offset ? adc_offset[k] : 0.0;
        }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long total = n_events * (long long)n_out * nb_samples;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
    if (in_dtype == DP_IN_F64)
        dp_combine_kernel<0><<<grid, 256, 0, st>>>(prm);
    else if (in_dtype == DP_IN_F32)
        dp_combine_kernel<1><<<grid, 256, 0, st>>>(prm);
    else
        dp_combine_kernel<2><<<grid, 256, 0, st>>>(prm);
    DP_CUDA(cudaGetLastError());
    return DP_OK;
}

int dp_reduce_plan_create(dp_reduce_plan** plan, int nb_samples, double sample_rate, int n_chan) {
    if (!plan) return fail(DP_ERR_INVALID, "null plan pointer");
    if (nb_samples < 1 || n_chan < 1 || !(sample_rate > 0)) return fail(DP_ERR_INVALID, "bad plan arguments");
    auto p = std::make_unique<dp_reduce_plan>();
    p->plan.nb_samples = nb_samples;
    p->plan.fs = sample_rate;
    p->plan.chan_feats.resize(n_chan);
    p->n_chan = n_chan;
    *plan = p.release();
    return DP_OK;
}
void dp_reduce_plan_destroy(dp_reduce_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}
int dp_reduce_plan_add(dp_reduce_plan* p, int chan, int op, int window_lo, int window_hi, int* feat_index) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    if (chan < 0 || chan >= p->n_chan) return fail(DP_ERR_INVALID, "channel index out of range");
    if (op < DP_OP_BASELINE || op > DP_OP_MINIMUM) return fail(DP_ERR_INVALID, "unknown op");
    // python slice semantics of trace[a:b]
    const int n = p->plan.nb_samples;
    auto clampi = [n](int v) { return v < 0 ? std::max(0, v + n) : std::min(v, n); };
    int lo = clampi(window_lo), hi = clampi(window_hi);
    if (hi < lo) hi = lo;
    if ((op == DP_OP_MAXIMUM || op == DP_OP_MINIMUM) && hi == lo)
        return fail(DP_ERR_INVALID, "zero-size array to reduction operation maximum/minimum which has no identity");
    p->plan.chan_feats[chan].push_back(dpred::Feat{op, lo, hi});
    if (feat_index) *feat_index = (int)p->plan.chan_feats[chan].size() - 1;
    return DP_OK;
}
int dp_reduce_plan_column(const dp_reduce_plan* p, int chan, int feat_index, int* column) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (chan < 0 || chan >= p->n_chan) return fail(DP_ERR_INVALID, "channel index out of range");
    if (feat_index < 0 || feat_index >= (int)p->plan.chan_feats[chan].size()) return fail(DP_ERR_INVALID, "feature index out of range");
    int col = 0;
    for (int c = 0; c < chan; ++c) col += (int)p->plan.chan_feats[c].size();
    *column = col + feat_index;
    return DP_OK;
}
int dp_reduce_plan_finalize(dp_reduce_plan* p, int device) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    try {
        dpred::finalize(p->plan);
    } catch (const std::exception& e) {
        return fail(DP_ERR_INVALID, e.what());
    }
    p->device = device;
    DP_ON_DEVICE(device);
    int rc;
    if ((rc = upload(p->owned, p->plan.chans, &p->d_chans))) return rc;
    if ((rc = upload(p->owned, p->plan.leaves, &p->d_leaves))) return rc;
    if ((rc = upload(p->owned, p->plan.nodes, &p->d_nodes))) return rc;
    if ((rc = upload(p->owned, p->plan.level_off, &p->d_level_off))) return rc;
    if ((rc = upload(p->owned, p->plan.feats, &p->d_feats))) return rc;
    if (p->adc.empty()) {
        p->adc.assign(2 * (size_t)p->n_chan, 0.0);
        for (int c = 0; c < p->n_chan; ++c) p->adc[2 * c] = 1.0;
    }
    if ((rc = upload(p->owned, p->adc, &p->d_adc))) return rc;
    p->smem = sizeof(double) * (size_t)(p->plan.max_nodes + 96);
    int occ = 0, sms = 0;
    DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dp_reduce_kernel<256>, 256, p->smem));
    DP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    p->grid_max = sms * std::max(occ, 1);
    DP_CUDA(cudaEventCreate(&p->ev0));
    DP_CUDA(cudaEventCreate(&p->ev1));
    p->finalized = true;
    return DP_OK;
}
int dp_reduce_plan_n_out(const dp_reduce_plan* p, int* n_out) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    *n_out = p->plan.n_out;
    return DP_OK;
}
int dp_reduce_plan_set_adc_conversion(dp_reduce_plan* p, int chan, double gain, double offset) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    if (chan < 0 || chan >= p->n_chan) return fail(DP_ERR_INVALID, "channel index out of range");
    if (!(gain > 0) || !std::isfinite(gain) || !std::isfinite(offset)) return fail(DP_ERR_INVALID, "adc gain must be finite and > 0");
    if (p->adc.empty()) {
        p->adc.assign(2 * (size_t)p->n_chan, 0.0);
        for (int c = 0; c < p->n_chan; ++c) p->adc[2 * c] = 1.0;
    }
    p->adc[2 * chan] = gain;
    p->adc[2 * chan + 1] = offset;
    return DP_OK;
}
int dp_window_reduce_batch(dp_reduce_plan* p, const double* traces_dev, long long n_events, long long row_stride,
                           double* out_dev, void* stream) {
    return dp_window_reduce_batch_raw(p, traces_dev, DP_IN_F64, n_events, row_stride, out_dev, stream);
}
int dp_window_reduce_batch_raw(dp_reduce_plan* p, const void* traces_dev, int in_dtype, long long n_events, long long row_stride,
                               double* out_dev, void* stream) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    if (row_stride < p->plan.nb_samples) return fail(DP_ERR_INVALID, "row_stride < nb_samples");
    return dp_window_reduce_batch_ex(p, traces_dev, in_dtype, n_events, (long long)p->n_chan * row_stride, nullptr, row_stride, nullptr, 0,
                                     out_dev, stream);
}
int dp_window_reduce_batch_ex(dp_reduce_plan* p, const void* base_dev, int in_dtype, long long n_events, long long event_stride,
                              const long long* chan_offsets, long long chan_stride, const long long* start_index_dev,
                              long long n_stream_samples, double* out_dev, void* stream) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    DP_ON_DEVICE(p->device);
    if (in_dtype != DP_IN_F64 && in_dtype != DP_IN_I16) return fail(DP_ERR_UNSUPPORTED, "window reductions take float64 or int16 traces");
    if (n_events < 0 || n_stream_samples < 0) return fail(DP_ERR_INVALID, "negative size");
    if (n_events == 0 || p->plan.n_out == 0) return DP_OK;
    if (!base_dev || !out_dev) return fail(DP_ERR_INVALID, "null buffer");
    if (n_events * p->n_chan > 2000000000LL) return fail(DP_ERR_INVALID, "batch too large; split it");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DpReduceParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.traces = base_dev;
    prm.adc = p->d_adc;
    prm.event_stride = event_stride;
    prm.chan_stride = chan_stride;
    prm.row_start = start_index_dev;
    prm.stream_len = n_stream_samples;
    prm.nb_samples = p->plan.nb_samples;
    if (chan_offsets != nullptr) {
        for (int c = 0; c < p->n_chan; ++c)
            if (chan_offsets[c] < 0) return fail(DP_ERR_INVALID, "negative channel offset");
        if (p->d_chan_off == nullptr) {
            DP_CUDA(cudaMalloc(&p->d_chan_off, sizeof(long long) * (size_t)p->n_chan));
            p->owned.push_back(p->d_chan_off);
        }
        if (p->chan_off_host.size() != (size_t)p->n_chan || !std::equal(p->chan_off_host.begin(), p->chan_off_host.end(), chan_offsets)) {
            DP_CUDA(cudaStreamSynchronize(st));
            DP_CUDA(cudaMemcpy(p->d_chan_off, chan_offsets, sizeof(long long) * (size_t)p->n_chan, cudaMemcpyHostToDevice));
            p->chan_off_host.assign(chan_offsets, chan_offsets + p->n_chan);
        }
        prm.chan_offset = p->d_chan_off;
    }
    prm.n_rows = (int)(n_events * p->n_chan);
    prm.n_chan = p->n_chan;
    prm.chans = p->d_chans;
    prm.leaves = p->d_leaves;
    prm.nodes = p->d_nodes;
    prm.level_off = p->d_level_off;
    prm.feats = p->d_feats;
    prm.out = out_dev;
    prm.n_out = p->plan.n_out;
    prm.fs = p->plan.fs;
    prm.max_nodes = p->plan.max_nodes;
    const int grid = (int)std::min<long long>(prm.n_rows, p->grid_max);
    DP_CUDA(cudaEventRecord(p->ev0, st));
    if (in_dtype == DP_IN_I16)
        dp_reduce_kernel<256, 2><<<grid, 256, p->smem, st>>>(prm);
    else
        dp_reduce_kernel<256, 0><<<grid, 256, p->smem, st>>>(prm);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaEventRecord(p->ev1, st));
    p->timed = true;
    return DP_OK;
}
int dp_reduce_plan_last_kernel_ms(dp_reduce_plan* p, float* ms) {
    if (!p || !p->finalized || !p->timed) return fail(DP_ERR_STATE, "no timed launch");
    DP_CUDA(cudaEventSynchronize(p->ev1));
    DP_CUDA(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return DP_OK;
}

}  // extern "C"

// ========================================================================== PSD plan
struct dp_psd_plan {
    int v2_r1 = 0;  // != 0: v2 FFT core (dp_psd2_kernel.cuh)
    const void *tw3 = nullptr, *groups = nullptr, *chunk3 = nullptr;
    void* park1 = nullptr;          // first-pass outputs that do not fit into TMEM (fp64, 65536 samples)
    long long park1_per_cta = 0;
    int N = 0;
    double fs = 0;
    int precision = DP_PREC_F64;
    int device = 0;
    dpplan::Geometry geom;
    std::vector<void*> owned;
    const void *tw1 = nullptr, *tw2 = nullptr, *twn = nullptr, *twp = nullptr;
    void* scratch = nullptr;
    long long scratch_per_cta = 0;
    double* partial = nullptr;
    long long partial_per_cta = 0;
    unsigned long long* count = nullptr;
    unsigned long long* count_out = nullptr;
    const int* loc = nullptr;
    int grid_max = 0;
    size_t smem = 0;
    double scale = 1.0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

namespace {

template <class T> int psd_finalize(dp_psd_plan* p) {
    // twiddle bases are shared with the OF kernel: build them through an empty table set
    std::vector<dpplan::Channel> none;
    dpplan::DeviceTables<T> dt;
    try {
        dt = dpplan::build_tables<T>(p->geom, p->fs, none, 0.0, 1.0);
    } catch (const std::exception& e) {
        return fail(DP_ERR_INVALID, e.what());
    }
    int rc;
    const cx<T>* d;
    if ((rc = upload(p->owned, dt.tw1, &d))) return rc;
    p->tw1 = d;
    if ((rc = upload(p->owned, dt.tw2, &d))) return rc;
    p->tw2 = d;
    if ((rc = upload(p->owned, dt.twn, &d))) return rc;
    p->twn = d;
    if ((rc = upload(p->owned, dt.twp, &d))) return rc;
    p->twp = d;
    const int prec = sizeof(T) == 8 ? 0 : 1;
    const int src = prec == 0 ? dp_psd_setup_p0_0(p->geom.R1, p->geom.P, p->device, &p->smem, &p->grid_max)
                              : dp_psd_setup_p1_0(p->geom.R1, p->geom.P, p->device, &p->smem, &p->grid_max);
    if (src == -1) return fail(DP_ERR_UNSUPPORTED, "unsupported trace length for this precision");
    if (src != 0) return fail(DP_ERR_CUDA, "PSD kernel setup failed");
    const int NT = p->geom.NT, P = p->geom.P, M = p->N / 2;
    p->scratch_per_cta = 32LL * NT;
    DP_CUDA(cudaMalloc(&p->scratch, sizeof(cx<T>) * (size_t)p->scratch_per_cta * (size_t)p->grid_max));
    p->owned.push_back(p->scratch);
    p->partial_per_cta = 32LL * P * NT + 17 * 2 * P;
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->partial), sizeof(double) * (size_t)p->partial_per_cta * (size_t)p->grid_max));
    p->owned.push_back(p->partial);
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->count), sizeof(unsigned long long) * (size_t)(p->grid_max + 1)));
    p->owned.push_back(p->count);
    p->count_out = p->count + p->grid_max;
    // natural bin k -> slot in one CTA's partial array (thread 0's placeholders are skipped)
    std::vector<int> loc(M + 1, -1);
    for (int t = 1; t < NT; ++t)
        for (int e = 0; e < 32 * P; ++e) loc[dpplan::k_of(p->geom, t, e)] = e * NT + t;
    for (int l = 0; l < 17; ++l) {
        int bins[4];
        bool dup[4];
        dpplan::self_bins(p->geom, l, bins, dup);
        for (int j = 0; j < 2 * P; ++j)
            if (!dup[j]) loc[bins[j]] = 32 * P * NT + l * 2 * P + j;
    }
    for (int k = 0; k <= M; ++k)
        if (loc[k] < 0) return fail(DP_ERR_STATE, "internal: PSD bin map incomplete");
    if ((rc = upload(p->owned, loc, &p->loc))) return rc;
    return DP_OK;
}

template <class T, int R1> int psd2_tables(dp_psd_plan* p) {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    std::vector<dpplan::Channel> none;
    dpplan2::Tables2<T> dt;
    try {
        dt = dpplan2::build_tables2<T, R1>(p->fs, none, 0.0, 1.0);
    } catch (const std::exception& e) {
        return fail(DP_ERR_STATE, e.what());
    }
    int rc;
    const cx<T>* d;
    if ((rc = upload(p->owned, dt.tw1, &d))) return rc;
    p->tw1 = d;
    if ((rc = upload(p->owned, dt.tw2, &d))) return rc;
    p->tw2 = d;
    if ((rc = upload(p->owned, dt.tw3, &d))) return rc;
    p->tw3 = d;
    const cx<S>* ds;
    if ((rc = upload(p->owned, dt.twn, &ds))) return rc;
    p->twn = ds;
    const int2* dg;
    if ((rc = upload(p->owned, dt.groups, &dg))) return rc;
    p->groups = dg;
    const int* dc3;
    if ((rc = upload(p->owned, dt.chunk3, &dc3))) return rc;
    p->chunk3 = dc3;
    p->park1_per_cta = Dp2Core<T, R1, 0>::CAN_PARK ? Dp2Core<T, R1, 0>::PARK1_V : 0;
    // natural bin k -> slot in one CTA's partial array
    const std::vector<int> loc = dpplan2::partial_slot_of_bin<G>();
    if ((rc = upload(p->owned, loc, &p->loc))) return rc;
    return DP_OK;
}

template <class T> int psd2_finalize(dp_psd_plan* p) {
    int rc;
    switch (p->v2_r1) {
        case 2: rc = psd2_tables<T, 2>(p); break;
        case 4: rc = psd2_tables<T, 4>(p); break;
        default: rc = psd2_tables<T, 8>(p); break;
    }
    if (rc) return rc;
    const int prec = sizeof(typename Dp2Traits<T>::S) == 8 ? 0 : 1;
    for (int in = 0; in < 3; ++in) {
        int gm = 0;
        const int src = dp_psd2_setup_table[prec][in](p->v2_r1, p->device, &p->smem, &gm, &p->partial_per_cta);
        if (src != 0) return fail(DP_ERR_CUDA, "PSD kernel setup failed");
        p->grid_max = in == 0 ? gm : std::min(p->grid_max, gm);
    }
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->partial), sizeof(double) * (size_t)p->partial_per_cta * (size_t)p->grid_max));
    p->owned.push_back(p->partial);
    if (p->park1_per_cta > 0) {
        DP_CUDA(cudaMalloc(&p->park1, 16 * (size_t)p->park1_per_cta * (size_t)p->grid_max));
        p->owned.push_back(p->park1);
    }
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->count), sizeof(unsigned long long) * (size_t)(p->grid_max + 1)));
    p->owned.push_back(p->count);
    p->count_out = p->count + p->grid_max;
    return DP_OK;
}

}  // namespace

extern "C" {

int dp_psd_plan_create(dp_psd_plan** plan, int nb_samples, double sample_rate, int precision, int device) {
    if (!plan) return fail(DP_ERR_INVALID, "null plan pointer");
    if (!(sample_rate > 0)) return fail(DP_ERR_INVALID, "sample_rate must be > 0");
    if (precision != DP_PREC_F64 && precision != DP_PREC_F32) return fail(DP_ERR_INVALID, "unknown precision");
    auto p = std::make_unique<dp_psd_plan>();
    const char* gen = std::getenv("DP_OF_KERNEL");
    p->v2_r1 = (gen && std::string(gen) == "v1") ? 0 : dpplan2::r1_of(nb_samples);
    if (!p->v2_r1) {
        try {
            p->geom = dpplan::pick_geometry(nb_samples, precision == DP_PREC_F64);
        } catch (const std::exception& e) {
            return fail(DP_ERR_UNSUPPORTED, e.what());
        }
    }
    p->N = nb_samples;
    p->fs = sample_rate;
    p->precision = precision;
    p->device = device;
    DP_ON_DEVICE(device);
    int rc;
    if (p->v2_r1)
        rc = precision == DP_PREC_F32 ? psd2_finalize<f2>(p.get()) : psd2_finalize<double>(p.get());
    else
        rc = precision == DP_PREC_F32 ? psd_finalize<float>(p.get()) : psd_finalize<double>(p.get());
    if (rc) {
        for (void* d : p->owned) cudaFree(d);
        return rc;
    }
    DP_CUDA(cudaEventCreate(&p->ev0));
    DP_CUDA(cudaEventCreate(&p->ev1));
    DP_CUDA(cudaMemset(p->partial, 0, sizeof(double) * (size_t)p->partial_per_cta * (size_t)p->grid_max));
    DP_CUDA(cudaMemset(p->count, 0, sizeof(unsigned long long) * (size_t)(p->grid_max + 1)));
    *plan = p.release();
    return DP_OK;
}
void dp_psd_plan_destroy(dp_psd_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}
int dp_psd_plan_set_scale(dp_psd_plan* p, double typical_rms) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (!(typical_rms > 0)) return fail(DP_ERR_INVALID, "typical_rms must be > 0");
    p->scale = p->precision == DP_PREC_F32 ? std::exp2(-std::round(std::log2(typical_rms))) : 1.0;
    return DP_OK;
}
int dp_psd_reset(dp_psd_plan* p, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    DP_ON_DEVICE(p->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DP_CUDA(cudaMemsetAsync(p->partial, 0, sizeof(double) * (size_t)p->partial_per_cta * (size_t)p->grid_max, st));
    DP_CUDA(cudaMemsetAsync(p->count, 0, sizeof(unsigned long long) * (size_t)(p->grid_max + 1), st));
    return DP_OK;
}
int dp_psd_accumulate(dp_psd_plan* p, const void* traces_dev, int in_dtype, long long n_traces, long long row_stride,
                      const unsigned char* mask_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    DP_ON_DEVICE(p->device);
    if (n_traces < 0) return fail(DP_ERR_INVALID, "negative n_traces");
    if (n_traces == 0) return DP_OK;
    if (!traces_dev) return fail(DP_ERR_INVALID, "null buffer");
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (in_dtype != DP_IN_F64 && !p->v2_r1) return fail(DP_ERR_UNSUPPORTED, "float32 / int16 traces need nb_samples 16384, 32768 or 65536");
    {
        const size_t esz = in_dtype == DP_IN_F64 ? 8 : (in_dtype == DP_IN_F32 ? 4 : 2);
        if ((reinterpret_cast<uintptr_t>(traces_dev) % (2 * esz)) != 0) return fail(DP_ERR_INVALID, "trace buffer misaligned");
    }
    if (row_stride < p->N || (row_stride & 1)) return fail(DP_ERR_INVALID, "row_stride must be even and >= nb_samples");
    if (n_traces > 2000000000LL) return fail(DP_ERR_INVALID, "batch too large; split it");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (int)std::min<long long>(n_traces, p->grid_max);
    auto fill = [&](auto& prm) {
        std::memset(&prm, 0, sizeof(prm));
        prm.traces = traces_dev;
        prm.row_stride = row_stride;
        prm.n_rows = (int)n_traces;
        prm.mask = mask_dev;
        prm.scratch_per_cta = p->scratch_per_cta;
        prm.partial = p->partial;
        prm.partial_per_cta = p->partial_per_cta;
        prm.count = p->count;
        prm.scale = p->scale;
        prm.subtract_first = p->precision == DP_PREC_F32 ? 1 : 0;
    };
    DP_CUDA(cudaEventRecord(p->ev0, st));
    int rc;
    if (p->v2_r1) {
        auto fill2 = [&](auto& prm) {
            std::memset(&prm, 0, sizeof(prm));
            prm.traces = traces_dev;
            prm.row_stride = row_stride;
            prm.n_rows = (int)n_traces;
            prm.mask = mask_dev;
            prm.groups = reinterpret_cast<const int2*>(p->groups);
            prm.chunk3 = reinterpret_cast<const int*>(p->chunk3);
            prm.partial = p->partial;
            prm.park1 = reinterpret_cast<decltype(prm.park1)>(p->park1);
            prm.partial_per_cta = p->partial_per_cta;
            prm.count = p->count;
            prm.scale = p->scale;
            prm.subtract_first = p->precision == DP_PREC_F32 ? 1 : 0;
        };
        if (p->precision == DP_PREC_F32) {
            DpPsd2Params<f2> prm;
            fill2(prm);
            prm.tw1 = (const cx<f2>*)p->tw1; prm.tw2 = (const cx<f2>*)p->tw2; prm.tw3 = (const cx<f2>*)p->tw3;
            prm.twn = (const cx<float>*)p->twn;
            rc = dp_psd2_launch_table[1][in_dtype](p->v2_r1, &prm, grid, p->smem, st);
        } else {
            DpPsd2Params<double> prm;
            fill2(prm);
            prm.tw1 = (const cx<double>*)p->tw1; prm.tw2 = (const cx<double>*)p->tw2; prm.tw3 = (const cx<double>*)p->tw3;
            prm.twn = (const cx<double>*)p->twn;
            rc = dp_psd2_launch_table[0][in_dtype](p->v2_r1, &prm, grid, p->smem, st);
        }
    } else if (p->precision == DP_PREC_F32) {
        DpPsdParams<float> prm;
        fill(prm);
        prm.tw1 = (const cx<float>*)p->tw1; prm.tw2 = (const cx<float>*)p->tw2;
        prm.twn = (const cx<float>*)p->twn; prm.twp = (const cx<float>*)p->twp;
        prm.scratch = (cx<float>*)p->scratch;
        rc = dp_psd_launch_p1_0(p->geom.R1, p->geom.P, &prm, grid, p->smem, st);
    } else {
        DpPsdParams<double> prm;
        fill(prm);
        prm.tw1 = (const cx<double>*)p->tw1; prm.tw2 = (const cx<double>*)p->tw2;
        prm.twn = (const cx<double>*)p->twn; prm.twp = (const cx<double>*)p->twp;
        prm.scratch = (cx<double>*)p->scratch;
        rc = dp_psd_launch_p0_0(p->geom.R1, p->geom.P, &prm, grid, p->smem, st);
    }
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("PSD kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    DP_CUDA(cudaEventRecord(p->ev1, st));
    p->timed = true;
    return DP_OK;
}
int dp_psd_get_sums(dp_psd_plan* p, double* sums_dev, unsigned long long* count_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    DP_ON_DEVICE(p->device);
    if (!sums_dev || !count_dev) return fail(DP_ERR_INVALID, "null buffer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nbins = p->N / 2 + 1;
    DP_CUDA(cudaMemsetAsync(sums_dev, 0, sizeof(double) * (size_t)nbins, st));
    DP_CUDA(cudaMemsetAsync(p->count_out, 0, sizeof(unsigned long long), st));
    DpPsdReduceParams prm;
    prm.partial = p->partial;
    prm.partial_per_cta = p->partial_per_cta;
    prm.grid = p->grid_max;
    prm.loc = p->loc;
    prm.nbins = nbins;
    prm.sum_out = sums_dev;
    prm.count = p->count;
    prm.count_out = p->count_out;
    const int rc = dp_psd_reduce_launch(&prm, st);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("PSD reduce launch: ") + cudaGetErrorString((cudaError_t)rc));
    DP_CUDA(cudaMemcpyAsync(count_dev, p->count_out, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    return DP_OK;
}
int dp_psd_plan_last_kernel_ms(dp_psd_plan* p, float* ms) {
    if (!p || !p->timed) return fail(DP_ERR_STATE, "no timed launch");
    DP_CUDA(cudaEventSynchronize(p->ev1));
    DP_CUDA(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return DP_OK;
}

}  // extern "C"

// ====================================================================== trigger plan
struct dp_trigger_plan {
    int nb_filter = 0;      // Nt taps
    int precision = DP_PREC_F64;
    int device = 0;
    int r1 = 0, F = 0;      // FFT size F = r1 * 8192
    int hop = 0, lead = 0;  // outputs per chunk, samples loaded ahead of the chunk's first output
    long long max_samples = 0;
    int max_chunks = 0;
    double iw = 1.0, w = 1.0, scale = 1.0;
    std::vector<double> phi_td;
    bool tables_ok = false;
    std::vector<void*> owned;
    const void *tw1 = nullptr, *tw2 = nullptr, *tw3 = nullptr, *twn = nullptr, *groups = nullptr, *chunk3 = nullptr;
    void* phi = nullptr;       // re-uploaded when the scale changes
    void* phi_self = nullptr;
    size_t phi_bytes = 0, phi_self_bytes = 0;
    void* scratch = nullptr;
    long long scratch_per_cta = 0;
    int* cand_idx = nullptr;
    double* cand_amp = nullptr;
    int* cand_count = nullptr;
    long long* chunk_offset = nullptr;
    void* park1 = nullptr;        // first-pass outputs that do not fit into TMEM (fp64, F = 65536)
    double* taps_dev = nullptr;   // phi_td on the device (dp_trigger_filtered_at)
    double* cand_val = nullptr;   // residual delta chi2 of the candidates (allocated by the first residual pass)
    int last_chunks = 0;          // chunk count of the last dp_trigger_run (the candidate list the follow-up calls use)
    bool resid_list = false;      // the candidate list holds residual survivors (cand_val is valid)
    // parallel grouping workspace
    int* tile_heads = nullptr;
    unsigned long long *best_key = nullptr, *best_g = nullptr;
    int best_cap = 0;
    int n_sm = 0;
    bool group_serial = false;  // DP_TRIG_GROUP=serial: the single-CTA grouping kernel (development cross-check)
    int grid_max = 0;
    size_t smem = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    bool timed = false;
};

namespace {

// filter spectrum on the one-sided bins, with the output alignment, iW and all scalings folded in
std::vector<dpplan::cplx> trig_filter_onesided(const dp_trigger_plan* p) {
    const int F = p->F, M = F / 2, Nt = p->nb_filter;
    std::vector<dpplan::cplx> a(F, dpplan::cplx(0, 0));
    for (int j = 0; j < Nt; ++j) a[j] = dpplan::cplx(p->phi_td[j], 0.0);
    dpplan::fft_pow2(a);
    const int c = (Nt - 1) / 2;
    const long long sh = (long long)p->lead + c;  // circular-convolution index of output r = 0
    std::vector<dpplan::cplx> pe(M + 1);
    for (int k = 0; k <= M; ++k) {
        const dpplan::cplx ramp = std::conj(dpplan::unit_root(((long long)k * sh) % F, F));  // e^{+2 pi i k sh / F}
        pe[k] = a[k] * ramp * (p->iw / ((double)F * 2.0 * p->scale));
    }
    pe[0] = dpplan::cplx(pe[0].real(), 0.0);
    pe[M] = dpplan::cplx(pe[M].real(), 0.0);
    return pe;
}

template <class T, int R1> int trig_tables(dp_trigger_plan* p, bool filter_only) {
    using S = typename Dp2Traits<T>::S;
    int rc;
    if (!filter_only) {
        std::vector<dpplan::Channel> none;
        dpplan2::Tables2<T> dt;
        try {
            dt = dpplan2::build_tables2<T, R1>(1.0, none, 0.0, 1.0);
        } catch (const std::exception& e) {
            return fail(DP_ERR_STATE, e.what());
        }
        const cx<T>* d;
        if ((rc = upload(p->owned, dt.tw1, &d))) return rc;
        p->tw1 = d;
        if ((rc = upload(p->owned, dt.tw2, &d))) return rc;
        p->tw2 = d;
        if ((rc = upload(p->owned, dt.tw3, &d))) return rc;
        p->tw3 = d;
        const cx<S>* ds;
        if ((rc = upload(p->owned, dt.twn, &ds))) return rc;
        p->twn = ds;
        const int2* dg;
        if ((rc = upload(p->owned, dt.groups, &dg))) return rc;
        p->groups = dg;
        const int* dc3;
        if ((rc = upload(p->owned, dt.chunk3, &dc3))) return rc;
        p->chunk3 = dc3;
    }
    std::vector<cx<T>> phi;
    std::vector<cx<S>> phi_self;
    dpplan2::pack_onesided<T, R1>(trig_filter_onesided(p), phi, phi_self);
    if (!p->phi) {
        p->phi_bytes = sizeof(cx<T>) * phi.size();
        p->phi_self_bytes = sizeof(cx<S>) * phi_self.size();
        DP_CUDA(cudaMalloc(&p->phi, p->phi_bytes));
        p->owned.push_back(p->phi);
        DP_CUDA(cudaMalloc(&p->phi_self, p->phi_self_bytes));
        p->owned.push_back(p->phi_self);
    }
    DP_CUDA(cudaMemcpy(p->phi, phi.data(), p->phi_bytes, cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(p->phi_self, phi_self.data(), p->phi_self_bytes, cudaMemcpyHostToDevice));
    return DP_OK;
}
template <class T> int trig_tables_r1(dp_trigger_plan* p, bool filter_only) {
    switch (p->r1) {
        case 2: return trig_tables<T, 2>(p, filter_only);
        case 4: return trig_tables<T, 4>(p, filter_only);
        default: return trig_tables<T, 8>(p, filter_only);
    }
}

}  // namespace

extern "C" {

int dp_trigger_plan_create(dp_trigger_plan** plan, const double* phi_td, int nb_filter, double iw, double w, int precision,
                           long long max_samples, int device) {
    if (!plan || !phi_td) return fail(DP_ERR_INVALID, "null pointer");
    if (nb_filter < 2 || nb_filter > 32768) return fail(DP_ERR_UNSUPPORTED, "filter length must be in [2, 32768] samples");
    if (precision != DP_PREC_F64 && precision != DP_PREC_F32) return fail(DP_ERR_INVALID, "unknown precision");
    if (max_samples < 2 * (long long)nb_filter) return fail(DP_ERR_INVALID, "max_samples must be >= 2 * filter length");
    if (!(w > 0) || !(iw > 0)) return fail(DP_ERR_INVALID, "weights must be > 0");
    auto p = std::make_unique<dp_trigger_plan>();
    p->nb_filter = nb_filter;
    p->precision = precision;
    p->device = device;
    p->iw = iw;
    p->w = w;
    p->phi_td.assign(phi_td, phi_td + nb_filter);
    p->F = nb_filter <= 8192 ? 16384 : (nb_filter <= 16384 ? 32768 : 65536);
    p->r1 = p->F / 8192;
    const int c = (nb_filter - 1) / 2, d = nb_filter - 1 - c;
    p->lead = d + (d & 1);
    p->hop = (p->F - p->lead - c) & ~1;
    p->max_samples = max_samples;
    p->max_chunks = (int)((max_samples + p->hop - 1) / p->hop);
    DP_ON_DEVICE(device);
    int rc = precision == DP_PREC_F32 ? trig_tables_r1<f2>(p.get(), false) : trig_tables_r1<double>(p.get(), false);
    if (!rc) {
        const int src = precision == DP_PREC_F32 ? dp_trig_setup_p1(p->r1, device, &p->smem, &p->grid_max, &p->scratch_per_cta)
                                                 : dp_trig_setup_p0(p->r1, device, &p->smem, &p->grid_max, &p->scratch_per_cta);
        if (src != 0) rc = fail(DP_ERR_CUDA, "trigger kernel setup failed");
    }
    auto alloc = [&](void** ptr, size_t bytes) {
        if (rc) return;
        if (cudaMalloc(ptr, std::max<size_t>(bytes, 16)) != cudaSuccess) {
            rc = fail(DP_ERR_CUDA, "cudaMalloc failed (trigger plan)");
            return;
        }
        p->owned.push_back(*ptr);
    };
    alloc(&p->scratch, 16 * (size_t)std::max<long long>(p->scratch_per_cta, 1) * (size_t)std::max(p->grid_max, 1));
    if (precision == DP_PREC_F64 && p->r1 == 8)   // Dp2Core<double, 8>::PARK1_V: one phase of two blocks
        alloc(&p->park1, 16 * (size_t)Dp2Core<double, 8, 0>::PARK1_V * (size_t)std::max(p->grid_max, 1));
    alloc(reinterpret_cast<void**>(&p->cand_idx), sizeof(int) * (size_t)p->max_chunks * (size_t)p->hop);
    alloc(reinterpret_cast<void**>(&p->cand_amp), sizeof(double) * (size_t)p->max_chunks * (size_t)p->hop);
    alloc(reinterpret_cast<void**>(&p->cand_count), sizeof(int) * (size_t)p->max_chunks);
    alloc(reinterpret_cast<void**>(&p->chunk_offset), sizeof(long long) * (size_t)(p->max_chunks + 1));
    alloc(reinterpret_cast<void**>(&p->tile_heads), sizeof(int) * ((size_t)p->max_chunks * (size_t)p->hop / 1024 + 2));
    {
        const char* e = std::getenv("DP_TRIG_GROUP");
        p->group_serial = e && std::string(e) == "serial";
        cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, p->device);
    }
    if (rc) {
        for (void* dptr : p->owned) cudaFree(dptr);
        return rc;
    }
    DP_CUDA(cudaEventCreate(&p->ev0));
    DP_CUDA(cudaEventCreate(&p->ev1));
    DP_CUDA(cudaEventCreate(&p->ev2));
    *plan = p.release();
    return DP_OK;
}

void dp_trigger_plan_destroy(dp_trigger_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->ev2) cudaEventDestroy(p->ev2);
    delete p;
}

int dp_trigger_plan_set_scale(dp_trigger_plan* p, double typical_rms) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (!(typical_rms > 0)) return fail(DP_ERR_INVALID, "typical_rms must be > 0");
    if (p->precision != DP_PREC_F32) return DP_OK;
    p->scale = std::exp2(-std::round(std::log2(typical_rms)));
    DP_ON_DEVICE(p->device);
    return trig_tables_r1<f2>(p, true);
}

int dp_trigger_plan_geometry(const dp_trigger_plan* p, int* fft_size, int* hop) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (fft_size) *fft_size = p->F;
    if (hop) *hop = p->hop;
    return DP_OK;
}

int dp_trigger_run(dp_trigger_plan* p, const double* trace_dev, long long n_samples, double chi2_threshold,
                   long long pileup_window_samples, long long index_shift, int padding, long long* trig_index_dev,
                   double* trig_amp_dev, double* trig_dchi2_dev, int max_triggers, int* n_triggers_dev, void* stream) {
    return dp_trigger_run_raw(p, trace_dev, DP_IN_F64, n_samples, chi2_threshold, pileup_window_samples, index_shift, padding,
                              trig_index_dev, trig_amp_dev, trig_dchi2_dev, max_triggers, n_triggers_dev, stream);
}

int dp_trigger_run_raw(dp_trigger_plan* p, const void* trace_dev, int in_dtype, long long n_samples, double chi2_threshold,
                       long long pileup_window_samples, long long index_shift, int padding, long long* trig_index_dev,
                       double* trig_amp_dev, double* trig_dchi2_dev, int max_triggers, int* n_triggers_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    DP_ON_DEVICE(p->device);
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (!trace_dev || !trig_index_dev || !trig_amp_dev || !trig_dchi2_dev || !n_triggers_dev) return fail(DP_ERR_INVALID, "null buffer");
    if (n_samples < 2 || n_samples > p->max_samples) return fail(DP_ERR_INVALID, "n_samples out of the plan's range");
    if ((reinterpret_cast<uintptr_t>(trace_dev) & 15) != 0) return fail(DP_ERR_INVALID, "trace buffer misaligned");
    if (max_triggers < 0 || pileup_window_samples < 0) return fail(DP_ERR_INVALID, "negative argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n_chunks = (int)((n_samples + p->hop - 1) / p->hop);
    const int Nt = p->nb_filter;
    long long vlo = 0, vhi = n_samples;
    if (padding) {  // oftrigger.py:676-679
        vlo = Nt;
        vhi = n_samples - Nt + ((Nt + 1) % 2);
    }
    auto fill = [&](auto& prm) {
        std::memset(&prm, 0, sizeof(prm));
        prm.trace = trace_dev;
        prm.n_samples = n_samples;
        prm.n_chunks = n_chunks;
        prm.hop = p->hop;
        prm.lead = p->lead;
        prm.valid_lo = vlo;
        prm.valid_hi = vhi;
        prm.groups = reinterpret_cast<const int2*>(p->groups);
        prm.chunk3 = reinterpret_cast<const int*>(p->chunk3);
        prm.park1 = reinterpret_cast<decltype(prm.park1)>(p->park1);
        prm.scratch_per_cta = p->scratch_per_cta;
        prm.w = p->w;
        prm.thr = chi2_threshold;
        prm.scale = p->scale;
        prm.subtract_first = p->precision == DP_PREC_F32 ? 1 : 0;
        prm.cand_idx = p->cand_idx;
        prm.cand_amp = p->cand_amp;
        prm.cand_count = p->cand_count;
    };
    const int grid = std::min(n_chunks, p->grid_max);
    DP_CUDA(cudaEventRecord(p->ev0, st));
    int rc;
    if (p->precision == DP_PREC_F32) {
        DpTrigParams<f2> prm;
        fill(prm);
        prm.tw1 = (const cx<f2>*)p->tw1; prm.tw2 = (const cx<f2>*)p->tw2; prm.tw3 = (const cx<f2>*)p->tw3;
        prm.twn = (const cx<float>*)p->twn;
        prm.phi = (const cx<f2>*)p->phi; prm.phi_self = (const cx<float>*)p->phi_self;
        prm.scratch = (cx<f2>*)p->scratch;
        rc = dp_trig_launch_p1(p->r1, in_dtype, &prm, grid, p->smem, st);
    } else {
        DpTrigParams<double> prm;
        fill(prm);
        prm.tw1 = (const cx<double>*)p->tw1; prm.tw2 = (const cx<double>*)p->tw2; prm.tw3 = (const cx<double>*)p->tw3;
        prm.twn = (const cx<double>*)p->twn;
        prm.phi = (const cx<double>*)p->phi; prm.phi_self = (const cx<double>*)p->phi_self;
        prm.scratch = (cx<double>*)p->scratch;
        rc = dp_trig_launch_p0(p->r1, in_dtype, &prm, grid, p->smem, st);
    }
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("trigger filter launch: ") + cudaGetErrorString((cudaError_t)rc));
    DP_CUDA(cudaEventRecord(p->ev1, st));
    DpTrigGroupParams gp;
    gp.cand_idx = p->cand_idx;
    gp.cand_amp = p->cand_amp;
    gp.cand_count = p->cand_count;
    gp.n_chunks = n_chunks;
    gp.hop = p->hop;
    gp.pileup_window = pileup_window_samples;
    gp.index_shift = index_shift;
    gp.w = p->w;
    gp.trig_index = trig_index_dev;
    gp.trig_amp = trig_amp_dev;
    gp.trig_dchi2 = trig_dchi2_dev;
    gp.max_triggers = max_triggers;
    gp.n_triggers = n_triggers_dev;
    gp.chunk_offset = p->chunk_offset;
    gp.tile_heads = p->tile_heads;
    gp.cand_val = nullptr;
    p->last_chunks = n_chunks;
    p->resid_list = false;
    if (!p->group_serial && p->best_cap < max_triggers) {
        void* a = nullptr;
        void* b = nullptr;
        DP_CUDA(cudaMalloc(&a, sizeof(unsigned long long) * (size_t)max_triggers));
        DP_CUDA(cudaMalloc(&b, sizeof(unsigned long long) * (size_t)max_triggers));
        p->owned.push_back(a);
        p->owned.push_back(b);
        p->best_key = reinterpret_cast<unsigned long long*>(a);
        p->best_g = reinterpret_cast<unsigned long long*>(b);
        p->best_cap = max_triggers;
    }
    gp.best_key = p->best_key;
    gp.best_g = p->best_g;
    rc = p->group_serial ? dp_trig_group_launch(&gp, st) : dp_trig_group_par_launch(&gp, std::max(1, 2 * p->n_sm), st);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("trigger group launch: ") + cudaGetErrorString((cudaError_t)rc));
    DP_CUDA(cudaEventRecord(p->ev2, st));
    p->timed = true;
    return DP_OK;
}

int dp_trigger_plan_last_kernel_ms(dp_trigger_plan* p, float* filter_ms, float* group_ms) {
    if (!p || !p->timed) return fail(DP_ERR_STATE, "no timed launch");
    DP_CUDA(cudaEventSynchronize(p->ev2));
    if (filter_ms) DP_CUDA(cudaEventElapsedTime(filter_ms, p->ev0, p->ev1));
    if (group_ms) DP_CUDA(cudaEventElapsedTime(group_ms, p->ev1, p->ev2));
    return DP_OK;
}

}  // extern "C"


namespace {
int trig_best_buffers(dp_trigger_plan* p, int max_triggers) {
    if (p->best_cap >= max_triggers) return DP_OK;
    void* a = nullptr;
    void* b = nullptr;
    DP_CUDA(cudaMalloc(&a, sizeof(unsigned long long) * (size_t)std::max(max_triggers, 1)));
    p->owned.push_back(a);
    DP_CUDA(cudaMalloc(&b, sizeof(unsigned long long) * (size_t)std::max(max_triggers, 1)));
    p->owned.push_back(b);
    p->best_key = reinterpret_cast<unsigned long long*>(a);
    p->best_g = reinterpret_cast<unsigned long long*>(b);
    p->best_cap = max_triggers;
    return DP_OK;
}
void trig_group_params(const dp_trigger_plan* p, DpTrigGroupParams& gp) {
    std::memset(&gp, 0, sizeof(gp));
    gp.cand_idx = p->cand_idx;
    gp.cand_amp = p->cand_amp;
    gp.cand_count = p->cand_count;
    gp.n_chunks = p->last_chunks;
    gp.hop = p->hop;
    gp.w = p->w;
    gp.chunk_offset = p->chunk_offset;
    gp.tile_heads = p->tile_heads;
    gp.best_key = p->best_key;
    gp.best_g = p->best_g;
    gp.cand_val = p->resid_list ? p->cand_val : nullptr;
}
}  // namespace

extern "C" {

int dp_trigger_candidates(dp_trigger_plan* p, long long* idx_dev, double* amp_dev, double* dchi2_dev, long long max_candidates,
                          long long* n_candidates_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->last_chunks <= 0) return fail(DP_ERR_STATE, "dp_trigger_run has not been called on this plan");
    if (!idx_dev || !amp_dev || !n_candidates_dev || max_candidates < 0) return fail(DP_ERR_INVALID, "bad argument");
    DP_ON_DEVICE(p->device);
    DpTrigGroupParams gp;
    trig_group_params(p, gp);
    DpTrigFlattenParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.cand_idx = p->cand_idx;
    fp.cand_amp = p->cand_amp;
    fp.cand_val = p->resid_list ? p->cand_val : nullptr;
    fp.cand_count = p->cand_count;
    fp.chunk_offset = p->chunk_offset;
    fp.n_chunks = p->last_chunks;
    fp.hop = p->hop;
    fp.w = p->w;
    fp.out_idx = idx_dev;
    fp.out_amp = amp_dev;
    fp.out_val = dchi2_dev;
    fp.max_out = max_candidates;
    fp.n_out = n_candidates_dev;
    const int rc = dp_trig_flatten_launch(&gp, &fp, std::max(1, std::min(p->last_chunks, 4 * std::max(p->n_sm, 1))), stream);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("trigger candidate list: ") + cudaGetErrorString((cudaError_t)rc));
    return DP_OK;
}

int dp_trigger_filtered_at(dp_trigger_plan* p, const void* trace_dev, int in_dtype, long long n_samples, const long long* idx_dev,
                           int n_idx, double* filtered_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (!trace_dev || !idx_dev || !filtered_dev || n_idx < 0 || n_samples < 1) return fail(DP_ERR_INVALID, "bad argument");
    if (n_idx == 0) return DP_OK;
    DP_ON_DEVICE(p->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!p->taps_dev) {  // the taps go to the device with the first call
        void* taps = nullptr;
        DP_CUDA(cudaMalloc(&taps, sizeof(double) * (size_t)p->nb_filter));
        p->owned.push_back(taps);
        DP_CUDA(cudaMemcpyAsync(taps, p->phi_td.data(), sizeof(double) * (size_t)p->nb_filter, cudaMemcpyHostToDevice, st));
        p->taps_dev = reinterpret_cast<double*>(taps);
    }
    DpTrigAtParams ap;
    std::memset(&ap, 0, sizeof(ap));
    ap.trace = trace_dev;
    ap.in_dtype = in_dtype;
    ap.n_samples = n_samples;
    ap.phi_td = p->taps_dev;
    ap.nt = p->nb_filter;
    ap.iw = p->iw;
    ap.idx = idx_dev;
    ap.n_idx = n_idx;
    ap.out = filtered_dev;
    const int rc = dp_trig_filtered_at_launch(&ap, stream);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("trigger filtered_at launch: ") + cudaGetErrorString((cudaError_t)rc));
    return DP_OK;
}

int dp_trigger_residual_run(dp_trigger_plan* p, const long long* pulse_start_dev, const double* pulse_amp2_dev, int n_pulses,
                            const double* shape_dev, int n_shape, double chi2_threshold, long long pileup_window_samples,
                            long long index_shift, long long* trig_index_dev, double* trig_amp_dev, double* trig_dchi2_dev,
                            int max_triggers, int* n_triggers_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->last_chunks <= 0) return fail(DP_ERR_STATE, "dp_trigger_run has not been called on this plan");
    if (n_pulses < 0 || n_shape < 1 || max_triggers < 0 || pileup_window_samples < 0) return fail(DP_ERR_INVALID, "bad argument");
    if (n_pulses > 0 && (!pulse_start_dev || !pulse_amp2_dev)) return fail(DP_ERR_INVALID, "null pulse list");
    if (!shape_dev || !trig_index_dev || !trig_amp_dev || !trig_dchi2_dev || !n_triggers_dev) return fail(DP_ERR_INVALID, "null buffer");
    DP_ON_DEVICE(p->device);
    if (!p->cand_val) {
        void* v = nullptr;
        DP_CUDA(cudaMalloc(&v, sizeof(double) * (size_t)p->max_chunks * (size_t)p->hop));
        p->owned.push_back(v);
        p->cand_val = reinterpret_cast<double*>(v);
    }
    int rc = trig_best_buffers(p, max_triggers);
    if (rc) return rc;
    DpTrigResidParams rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.cand_idx = p->cand_idx;
    rp.cand_amp = p->cand_amp;
    rp.cand_val = p->cand_val;
    rp.cand_count = p->cand_count;
    rp.n_chunks = p->last_chunks;
    rp.hop = p->hop;
    rp.w = p->w;
    rp.thr = chi2_threshold;
    rp.pulse_start = pulse_start_dev;
    rp.pulse_a2 = pulse_amp2_dev;
    rp.n_pulses = n_pulses;
    rp.shape = shape_dev;
    rp.n_shape = n_shape;
    // an empty pulse list on a list that already holds residual survivors only regroups it (the caller's retry with larger
    // output buffers): running the kernel again would replace the stored residuals by amp^2 w
    if (!(n_pulses == 0 && p->resid_list)) {
        rc = dp_trig_residual_launch(&rp, std::max(1, std::min(p->last_chunks, 2 * std::max(p->n_sm, 1))), stream);
        if (rc != 0) return fail(DP_ERR_CUDA, std::string("trigger residual launch: ") + cudaGetErrorString((cudaError_t)rc));
    }
    p->resid_list = true;
    DpTrigGroupParams gp;
    trig_group_params(p, gp);
    gp.pileup_window = pileup_window_samples;
    gp.index_shift = index_shift;
    gp.trig_index = trig_index_dev;
    gp.trig_amp = trig_amp_dev;
    gp.trig_dchi2 = trig_dchi2_dev;
    gp.max_triggers = max_triggers;
    gp.n_triggers = n_triggers_dev;
    rc = dp_trig_group_par_launch(&gp, std::max(1, 2 * p->n_sm), stream);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("trigger group launch: ") + cudaGetErrorString((cudaError_t)rc));
    return DP_OK;
}

}  // extern "C"

// ======================================================================= band amplitudes (psd_amp)
struct dp_band_plan {
    int N = 0, n_bands = 0, device = 0;
    double fs = 0;
    int* bin_lo = nullptr;
    int* bin_hi = nullptr;
    double2* roots = nullptr;
    int n_sm = 0;
    std::vector<void*> owned;
};

extern "C" {

int dp_band_plan_create(dp_band_plan** plan, int nb_samples, double sample_rate, const int* bin_lo, const int* bin_hi, int n_bands,
                        int device) {
    if (!plan || !bin_lo || !bin_hi) return fail(DP_ERR_INVALID, "null pointer");
    if (nb_samples < 2 || !(sample_rate > 0) || n_bands < 1) return fail(DP_ERR_INVALID, "bad argument");
    for (int b = 0; b < n_bands; ++b)
        if (bin_lo[b] < 0 || bin_hi[b] <= bin_lo[b] || bin_hi[b] > nb_samples / 2 + 1) return fail(DP_ERR_INVALID, "bin range outside the one-sided spectrum");
    auto p = std::make_unique<dp_band_plan>();
    p->N = nb_samples;
    p->fs = sample_rate;
    p->n_bands = n_bands;
    p->device = device;
    DP_ON_DEVICE(device);
    std::vector<double2> roots((size_t)nb_samples);
    for (int j = 0; j < nb_samples; ++j) {
        const dpplan::cplx w = dpplan::unit_root(j, nb_samples);   // exp(-2 pi i j / N), exact octant symmetry
        roots[(size_t)j] = make_double2(w.real(), w.imag());
    }
    int rc;
    const double2* dr;
    if ((rc = upload(p->owned, roots, &dr))) return rc;
    p->roots = const_cast<double2*>(dr);
    const std::vector<int> lo(bin_lo, bin_lo + n_bands), hi(bin_hi, bin_hi + n_bands);
    const int* di;
    if ((rc = upload(p->owned, lo, &di))) return rc;
    p->bin_lo = const_cast<int*>(di);
    if ((rc = upload(p->owned, hi, &di))) return rc;
    p->bin_hi = const_cast<int*>(di);
    cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, device);
    *plan = p.release();
    return DP_OK;
}

void dp_band_plan_destroy(dp_band_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    delete p;
}

int dp_band_amplitudes(dp_band_plan* p, const void* base_dev, int in_dtype, long long n_events, long long event_stride, double adc_gain,
                       double adc_offset, double* out_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (in_dtype < DP_IN_F64 || in_dtype > DP_IN_I16) return fail(DP_ERR_INVALID, "unknown in_dtype");
    if (!base_dev || !out_dev || n_events < 0 || event_stride < p->N) return fail(DP_ERR_INVALID, "bad argument");
    if (n_events == 0) return DP_OK;
    DP_ON_DEVICE(p->device);
    DpBandParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.base = base_dev;
    prm.in_dtype = in_dtype;
    prm.n_events = n_events;
    prm.event_stride = event_stride;
    prm.N = p->N;
    prm.gain = adc_gain;
    prm.offset = adc_offset;
    prm.bin_lo = p->bin_lo;
    prm.bin_hi = p->bin_hi;
    prm.n_bands = p->n_bands;
    prm.roots = p->roots;
    prm.norm = (double)p->N / (p->fs * p->fs * p->fs);
    prm.out = out_dev;
    const int grid = (int)std::min<long long>(n_events, 8LL * std::max(p->n_sm, 1));
    const int rc = dp_band_launch(&prm, grid, stream);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("band amplitude launch: ") + cudaGetErrorString((cudaError_t)rc));
    return DP_OK;
}

}  // extern "C"

// ======================================================================= NxM plan
struct dp_nxm_plan {
    int N = 0, n = 0, m = 0, precision = DP_PREC_F64;
    double fs = 0;
    int r1 = 0;
    dpnxm::Setup setup;
    bool have_filter = false, finalized = false;
    double rms = 1.0;
    bool ac = true;
    int lo = 0, hi = 0, outside = 0;
    int device = 0;
    double scale = 1.0;
    int subtract_first = 0;
    std::vector<void*> owned;
    DpNxmParams<double> prm64;
    DpNxmParams<f2> prm32;
    int grid_max = 0, threads = 0;
    size_t smem = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

namespace {
template <class T> int nxm_finalize(dp_nxm_plan* p, DpNxmParams<T>& prm) {
    using S = typename Dp2Traits<T>::S;
    dpnxm::Tables<T> dt;
    try {
        switch (p->r1) {
            case 2: dt = dpnxm::build_tables<T, 2>(p->setup, p->scale); break;
            case 4: dt = dpnxm::build_tables<T, 4>(p->setup, p->scale); break;
            default: dt = dpnxm::build_tables<T, 8>(p->setup, p->scale); break;
        }
    } catch (const std::exception& e) {
        return fail(DP_ERR_INVALID, e.what());
    }
    std::memset(&prm, 0, sizeof(prm));
    int rc;
    if ((rc = upload(p->owned, dt.tw1, &prm.tw1))) return rc;
    if ((rc = upload(p->owned, dt.tw2, &prm.tw2))) return rc;
    if ((rc = upload(p->owned, dt.tw3, &prm.tw3))) return rc;
    if ((rc = upload(p->owned, dt.twn, &prm.twn))) return rc;
    if ((rc = upload(p->owned, dt.groups, &prm.groups))) return rc;
    if ((rc = upload(p->owned, dt.chunk3, &prm.chunk3))) return rc;
    if ((rc = upload(p->owned, dt.g, &prm.g))) return rc;
    if ((rc = upload(p->owned, dt.g_self, &prm.g_self))) return rc;
    for (int a = 0; a < p->n; ++a) {
        if ((rc = upload(p->owned, dt.wd[a], &prm.wd[a]))) return rc;
        if ((rc = upload(p->owned, dt.wd_self[a], &prm.wd_self[a]))) return rc;
    }
    for (size_t k = 0; k < dt.wo.size(); ++k) {
        if ((rc = upload(p->owned, dt.wo[k], &prm.wo[k]))) return rc;
        if ((rc = upload(p->owned, dt.wo_self[k], &prm.wo_self[k]))) return rc;
    }
    std::memcpy(prm.cmat, dt.cmat, sizeof(prm.cmat));
    std::memcpy(prm.amat, dt.amat, sizeof(prm.amat));
    const int prec = sizeof(S) == 8 ? 0 : 1;
    const int src = dp_nxm_setup_table[prec][p->n - 1](p->r1, p->device, &p->smem, &p->grid_max, &p->threads);
    if (src == -1) return fail(DP_ERR_UNSUPPORTED, "unsupported trace length");
    if (src != 0) return fail(DP_ERR_CUDA, std::string("NxM kernel setup: ") + cudaGetErrorString((cudaError_t)src));
    prm.scratch_per_cta = dp_nxm_scratch_table[prec][p->n - 1](p->r1, p->n, p->m);
    void* scr = nullptr;
    DP_CUDA(cudaMalloc(&scr, sizeof(cx<T>) * (size_t)prm.scratch_per_cta * (size_t)p->grid_max));
    p->owned.push_back(scr);
    prm.scratch = reinterpret_cast<cx<T>*>(scr);
    prm.n_chan = p->n;
    prm.n_templ = p->m;
    prm.pretrigger = p->setup.pretrigger;
    prm.n_out = 4 + 2 * p->m;
    prm.scale = p->scale;
    prm.subtract_first = p->subtract_first;
    {
        const char* e = std::getenv("DP_NXM_PREFETCH");  // development switch, default on
        prm.prefetch = (e && std::string(e) == "0") ? 0 : 1;
    }
    return DP_OK;
}
template <class T> int nxm_run(dp_nxm_plan* p, DpNxmParams<T> prm, const double* traces, long long n_events, long long ev_stride,
                               long long chan_stride, double* out, cudaStream_t st) {
    prm.traces = traces;
    prm.ev_stride = ev_stride;
    prm.chan_stride = chan_stride;
    prm.n_events = (int)n_events;
    prm.lo = p->lo;
    prm.hi = p->hi;
    prm.outside = p->outside;
    prm.out = out;
    const int grid = (int)std::min<long long>(n_events, p->grid_max);
    DP_CUDA(cudaEventRecord(p->ev0, st));
    const int prec = sizeof(typename Dp2Traits<T>::S) == 8 ? 0 : 1;
    const int rc = dp_nxm_launch_table[prec][p->n - 1](p->r1, &prm, grid, p->smem, st);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("NxM kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    DP_CUDA(cudaEventRecord(p->ev1, st));
    p->timed = true;
    return DP_OK;
}
}  // namespace

extern "C" {

int dp_nxm_plan_create(dp_nxm_plan** plan, int nb_samples, double sample_rate, int n_chan, int n_templ, int precision) {
    if (!plan) return fail(DP_ERR_INVALID, "null plan pointer");
    if (!dpplan2::r1_of(nb_samples)) return fail(DP_ERR_UNSUPPORTED, "the NxM filter needs nb_samples 16384, 32768 or 65536");
    if (n_chan < 1 || n_chan > DP_NXM_MAX_CHAN) return fail(DP_ERR_INVALID, "n_chan must be 1.." + std::to_string(DP_NXM_MAX_CHAN));
    if (n_templ < 1 || n_templ > DP_NXM_MAX_TEMPL) return fail(DP_ERR_INVALID, "n_templ must be 1.." + std::to_string(DP_NXM_MAX_TEMPL));
    if (!(sample_rate > 0)) return fail(DP_ERR_INVALID, "sample_rate must be > 0");
    if (precision != DP_PREC_F64 && precision != DP_PREC_F32) return fail(DP_ERR_INVALID, "unknown precision");
    auto p = std::make_unique<dp_nxm_plan>();
    p->N = nb_samples;
    p->fs = sample_rate;
    p->n = n_chan;
    p->m = n_templ;
    p->precision = precision;
    p->r1 = dpplan2::r1_of(nb_samples);
    p->hi = nb_samples;
    *plan = p.release();
    return DP_OK;
}

void dp_nxm_plan_destroy(dp_nxm_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}

int dp_nxm_plan_set_filter(dp_nxm_plan* p, const double* templates, const double* csd, int pretrigger_samples, int coupling_ac) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    if (!templates || !csd) return fail(DP_ERR_INVALID, "null templates / csd");
    if (pretrigger_samples < 0 || pretrigger_samples >= p->N) return fail(DP_ERR_INVALID, "pretrigger_samples out of range");
    try {
        p->setup = dpnxm::make_setup(p->N, p->fs, p->n, p->m, templates, csd, pretrigger_samples, coupling_ac != 0);
        p->rms = dpnxm::typical_rms(p->setup, csd);
    } catch (const std::exception& e) {
        return fail(DP_ERR_INVALID, e.what());
    }
    p->ac = coupling_ac != 0;
    p->have_filter = true;
    return DP_OK;
}

int dp_nxm_plan_set_window(dp_nxm_plan* p, int window_lo, int window_hi, int outside) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (window_lo < 0 || window_hi > p->N || window_lo > window_hi) return fail(DP_ERR_INVALID, "bad delay window");
    p->lo = window_lo;
    p->hi = window_hi;
    p->outside = outside ? 1 : 0;
    return DP_OK;
}

int dp_nxm_plan_finalize(dp_nxm_plan* p, int device) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (p->finalized) return fail(DP_ERR_STATE, "plan already finalized");
    if (!p->have_filter) return fail(DP_ERR_STATE, "set the templates and the csd first");
    p->device = device;
    DP_ON_DEVICE(device);
    if (p->precision == DP_PREC_F32) {
        p->scale = std::exp2(-std::round(std::log2(p->rms)));
        p->subtract_first = p->ac ? 1 : 0;
    }
    const int rc = p->precision == DP_PREC_F32 ? nxm_finalize<f2>(p, p->prm32) : nxm_finalize<double>(p, p->prm64);
    if (rc) return rc;
    DP_CUDA(cudaEventCreate(&p->ev0));
    DP_CUDA(cudaEventCreate(&p->ev1));
    p->finalized = true;
    return DP_OK;
}

int dp_nxm_plan_n_out(const dp_nxm_plan* p, int* n_out) {
    if (!p || !n_out) return fail(DP_ERR_INVALID, "null argument");
    *n_out = 4 + 2 * p->m;
    return DP_OK;
}

int dp_nxm_plan_get_p_matrix(const dp_nxm_plan* p, double* p_matrix, double* p_inverse) {
    if (!p || !p->have_filter) return fail(DP_ERR_STATE, "no filter set");
    for (int i = 0; i < p->m * p->m; ++i) {
        if (p_matrix) p_matrix[i] = p->setup.P[i];
        if (p_inverse) p_inverse[i] = p->setup.Pinv[i];
    }
    return DP_OK;
}

int dp_ofnxm_batch(dp_nxm_plan* p, const double* traces_dev, long long n_events, long long event_stride, long long chan_stride,
                   double* out_dev, void* stream) {
    if (!p || !p->finalized) return fail(DP_ERR_STATE, "plan not finalized");
    if (n_events < 0) return fail(DP_ERR_INVALID, "negative n_events");
    if (n_events == 0) return DP_OK;
    if (!traces_dev || !out_dev) return fail(DP_ERR_INVALID, "null buffer");
    if (chan_stride < p->N || (chan_stride & 1) || (event_stride & 1) || event_stride < chan_stride * (p->n - 1) + p->N)
        return fail(DP_ERR_INVALID, "strides must be even, chan_stride >= nb_samples, event_stride >= the channels of an event");
    if ((reinterpret_cast<uintptr_t>(traces_dev) & 15) != 0) return fail(DP_ERR_INVALID, "trace buffer misaligned");
    if (n_events > 2000000000LL) return fail(DP_ERR_INVALID, "batch too large; split it");
    DP_ON_DEVICE(p->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p->precision == DP_PREC_F32) return nxm_run<f2>(p, p->prm32, traces_dev, n_events, event_stride, chan_stride, out_dev, st);
    return nxm_run<double>(p, p->prm64, traces_dev, n_events, event_stride, chan_stride, out_dev, st);
}

int dp_nxm_plan_last_kernel_ms(dp_nxm_plan* p, float* ms) {
    if (!p || !p->finalized || !p->timed) return fail(DP_ERR_STATE, "no timed launch");
    DP_CUDA(cudaEventSynchronize(p->ev1));
    DP_CUDA(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return DP_OK;
}

}  // extern "C"

// ======================================================================= CSD plan
struct dp_csd_plan {
    int N = 0, n = 0, precision = DP_PREC_F64, r1 = 0, device = 0;
    double fs = 0, scale = 1.0;
    std::vector<void*> owned;
    const void *tw1 = nullptr, *tw2 = nullptr, *tw3 = nullptr, *twn = nullptr, *groups = nullptr, *chunk3 = nullptr;
    const int* loc = nullptr;
    void* scratch = nullptr;
    long long scratch_per_cta = 0;
    double* partial = nullptr;
    long long partial_per_comp = 0, partial_per_cta = 0;
    int ncp = 0;  // components per partial-sum slot (n*n padded to even)
    unsigned long long *count = nullptr, *count_out = nullptr;
    int grid_max = 0;
    size_t smem = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

namespace {
template <class T, int R1> int csd_tables(dp_csd_plan* p) {
    using G = Dp2Geom<T, R1>;
    using S = typename G::S;
    std::vector<dpplan::Channel> none;
    dpplan2::Tables2<T> dt;
    try {
        dt = dpplan2::build_tables2<T, R1>(p->fs, none, 0.0, 1.0);
    } catch (const std::exception& e) {
        return fail(DP_ERR_STATE, e.what());
    }
    int rc;
    const cx<T>* d;
    if ((rc = upload(p->owned, dt.tw1, &d))) return rc;
    p->tw1 = d;
    if ((rc = upload(p->owned, dt.tw2, &d))) return rc;
    p->tw2 = d;
    if ((rc = upload(p->owned, dt.tw3, &d))) return rc;
    p->tw3 = d;
    const cx<S>* ds;
    if ((rc = upload(p->owned, dt.twn, &ds))) return rc;
    p->twn = ds;
    const int2* dg;
    if ((rc = upload(p->owned, dt.groups, &dg))) return rc;
    p->groups = dg;
    const int* dc3;
    if ((rc = upload(p->owned, dt.chunk3, &dc3))) return rc;
    p->chunk3 = dc3;
    // natural bin k -> slot of one component in a CTA's partial array (same map as the PSD plan)
    const std::vector<int> loc = dpplan2::partial_slot_of_bin<G>();
    if ((rc = upload(p->owned, loc, &p->loc))) return rc;
    return DP_OK;
}
template <class T> int csd_finalize(dp_csd_plan* p) {
    int rc;
    switch (p->r1) {
        case 2: rc = csd_tables<T, 2>(p); break;
        case 4: rc = csd_tables<T, 4>(p); break;
        default: rc = csd_tables<T, 8>(p); break;
    }
    if (rc) return rc;
    const int prec = sizeof(typename Dp2Traits<T>::S) == 8 ? 0 : 1;
    const int src = dp_csd_setup_table[prec][p->n - 2](p->r1, p->device, &p->smem, &p->grid_max, &p->partial_per_comp, &p->scratch_per_cta, &p->ncp);
    if (src != 0) return fail(DP_ERR_CUDA, "CSD kernel setup failed");
    p->partial_per_cta = p->partial_per_comp * p->ncp;
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->partial), sizeof(double) * (size_t)p->partial_per_cta * (size_t)p->grid_max));
    p->owned.push_back(p->partial);
    DP_CUDA(cudaMalloc(&p->scratch, sizeof(cx<T>) * (size_t)p->scratch_per_cta * (size_t)p->grid_max));
    p->owned.push_back(p->scratch);
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->count), sizeof(unsigned long long) * (size_t)(p->grid_max + 1)));
    p->owned.push_back(p->count);
    p->count_out = p->count + p->grid_max;
    return DP_OK;
}
template <class T> int csd_launch(dp_csd_plan* p, const double* traces, long long n_events, long long ev_stride, long long chan_stride,
                                  const unsigned char* mask, cudaStream_t st) {
    using S = typename Dp2Traits<T>::S;
    DpCsdParams<T> prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.traces = traces;
    prm.ev_stride = ev_stride;
    prm.chan_stride = chan_stride;
    prm.n_events = (int)n_events;
    prm.mask = mask;
    prm.tw1 = (const cx<T>*)p->tw1;
    prm.tw2 = (const cx<T>*)p->tw2;
    prm.tw3 = (const cx<T>*)p->tw3;
    prm.twn = (const cx<S>*)p->twn;
    prm.groups = (const int2*)p->groups;
    prm.chunk3 = (const int*)p->chunk3;
    prm.scratch = (cx<T>*)p->scratch;
    prm.scratch_per_cta = p->scratch_per_cta;
    prm.partial = p->partial;
    prm.partial_per_cta = p->partial_per_cta;
    prm.count = p->count;
    prm.scale = p->scale;
    prm.subtract_first = p->precision == DP_PREC_F32 ? 1 : 0;
    const int grid = (int)std::min<long long>(n_events, p->grid_max);
    const int prec = sizeof(S) == 8 ? 0 : 1;
    return dp_csd_launch_table[prec][p->n - 2](p->r1, &prm, grid, p->smem, st);
}
}  // namespace

extern "C" {

int dp_csd_plan_create(dp_csd_plan** plan, int nb_samples, double sample_rate, int n_chan, int precision, int device) {
    if (!plan) return fail(DP_ERR_INVALID, "null plan pointer");
    if (!(sample_rate > 0)) return fail(DP_ERR_INVALID, "sample_rate must be > 0");
    if (precision != DP_PREC_F64 && precision != DP_PREC_F32) return fail(DP_ERR_INVALID, "unknown precision");
    if (!dpplan2::r1_of(nb_samples)) return fail(DP_ERR_UNSUPPORTED, "the CSD estimator needs nb_samples 16384, 32768 or 65536");
    if (n_chan < 2 || n_chan > DP_CSD_MAX_CHAN) return fail(DP_ERR_INVALID, "n_chan must be 2.." + std::to_string(DP_CSD_MAX_CHAN));
    auto p = std::make_unique<dp_csd_plan>();
    p->N = nb_samples;
    p->fs = sample_rate;
    p->n = n_chan;
    p->precision = precision;
    p->device = device;
    p->r1 = dpplan2::r1_of(nb_samples);
    DP_ON_DEVICE(device);
    const int rc = precision == DP_PREC_F32 ? csd_finalize<f2>(p.get()) : csd_finalize<double>(p.get());
    if (rc) {
        for (void* d : p->owned) cudaFree(d);
        return rc;
    }
    DP_CUDA(cudaEventCreate(&p->ev0));
    DP_CUDA(cudaEventCreate(&p->ev1));
    DP_CUDA(cudaMemset(p->partial, 0, sizeof(double) * (size_t)p->partial_per_cta * (size_t)p->grid_max));
    DP_CUDA(cudaMemset(p->count, 0, sizeof(unsigned long long) * (size_t)(p->grid_max + 1)));
    *plan = p.release();
    return DP_OK;
}
void dp_csd_plan_destroy(dp_csd_plan* p) {
    if (!p) return;
    for (void* d : p->owned) cudaFree(d);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    delete p;
}
int dp_csd_plan_set_scale(dp_csd_plan* p, double typical_rms) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (!(typical_rms > 0)) return fail(DP_ERR_INVALID, "typical_rms must be > 0");
    p->scale = p->precision == DP_PREC_F32 ? std::exp2(-std::round(std::log2(typical_rms))) : 1.0;
    return DP_OK;
}
int dp_csd_reset(dp_csd_plan* p, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    DP_ON_DEVICE(p->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DP_CUDA(cudaMemsetAsync(p->partial, 0, sizeof(double) * (size_t)p->partial_per_cta * (size_t)p->grid_max, st));
    DP_CUDA(cudaMemsetAsync(p->count, 0, sizeof(unsigned long long) * (size_t)(p->grid_max + 1), st));
    return DP_OK;
}
int dp_csd_accumulate(dp_csd_plan* p, const double* traces_dev, long long n_events, long long event_stride, long long chan_stride,
                      const unsigned char* mask_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    if (n_events < 0) return fail(DP_ERR_INVALID, "negative n_events");
    if (n_events == 0) return DP_OK;
    if (!traces_dev) return fail(DP_ERR_INVALID, "null buffer");
    if (chan_stride < p->N || (chan_stride & 1) || (event_stride & 1) || event_stride < chan_stride * (p->n - 1) + p->N)
        return fail(DP_ERR_INVALID, "strides must be even, chan_stride >= nb_samples, event_stride >= the channels of an event");
    if ((reinterpret_cast<uintptr_t>(traces_dev) & 15) != 0) return fail(DP_ERR_INVALID, "trace buffer misaligned");
    if (n_events > 2000000000LL) return fail(DP_ERR_INVALID, "batch too large; split it");
    DP_ON_DEVICE(p->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DP_CUDA(cudaEventRecord(p->ev0, st));
    const int rc = p->precision == DP_PREC_F32 ? csd_launch<f2>(p, traces_dev, n_events, event_stride, chan_stride, mask_dev, st)
                                               : csd_launch<double>(p, traces_dev, n_events, event_stride, chan_stride, mask_dev, st);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("CSD kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
    DP_CUDA(cudaEventRecord(p->ev1, st));
    p->timed = true;
    return DP_OK;
}
int dp_csd_get_sums(dp_csd_plan* p, double* sums_dev, unsigned long long* count_dev, void* stream) {
    if (!p) return fail(DP_ERR_INVALID, "null plan");
    DP_ON_DEVICE(p->device);
    if (!sums_dev || !count_dev) return fail(DP_ERR_INVALID, "null buffer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nbins = p->N / 2 + 1, ncomp = p->n * p->n;
    DP_CUDA(cudaMemsetAsync(sums_dev, 0, sizeof(double) * (size_t)nbins * ncomp, st));
    DP_CUDA(cudaMemsetAsync(p->count_out, 0, sizeof(unsigned long long), st));
    DpCsdReduceParams prm;
    prm.partial = p->partial;
    prm.partial_per_cta = p->partial_per_cta;
    prm.ncp = p->ncp;
    prm.grid = p->grid_max;
    prm.loc = p->loc;
    prm.nbins = nbins;
    prm.ncomp = ncomp;
    prm.sum_out = sums_dev;
    prm.count = p->count;
    prm.count_out = p->count_out;
    const int rc = dp_csd_reduce_launch(&prm, st);
    if (rc != 0) return fail(DP_ERR_CUDA, std::string("CSD reduce launch: ") + cudaGetErrorString((cudaError_t)rc));
    DP_CUDA(cudaMemcpyAsync(count_dev, p->count_out, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    return DP_OK;
}
int dp_csd_plan_last_kernel_ms(dp_csd_plan* p, float* ms) {
    if (!p || !p->timed) return fail(DP_ERR_STATE, "no timed launch");
    DP_CUDA(cudaEventSynchronize(p->ev1));
    DP_CUDA(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return DP_OK;
}

}  // extern "C"
