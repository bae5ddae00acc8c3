// One translation unit per (precision, input type) of the v2 fused OF kernel
// (nb_samples 16384 / 32768 / 65536), so they compile in parallel.  Build with
//   -DDP_INST_PREC=0|1 (double | packed float)  -DDP_INST_IN=0|1|2 (f64 | f32 | i16)
#ifndef DP_INST_PREC
#error "DP_INST_PREC must be defined"
#endif
#ifndef DP_INST_IN
#error "DP_INST_IN must be defined"
#endif
#include <cuda_runtime.h>

#include "dp_of2_kernel.cuh"
#include "dp_psd2_kernel.cuh"
#include "dp_of2_launch.hpp"

#if DP_INST_PREC == 0
using InstT = double;
#else
using InstT = f2;
#endif

#define DP_CAT_(a, b, c) a##b##_##c
#define DP_CAT(a, b, c) DP_CAT_(a, b, c)

namespace {
template <int R1> int setup_one(int device, size_t* smem, int* grid_max, int* occ_out, int* threads) {
    using K = Dp2OfKernel<InstT, R1, DP_INST_IN>;
    auto kern = dp_of2_kernel<InstT, R1, DP_INST_IN, false, false>;
    auto kern_m = dp_of2_kernel<InstT, R1, DP_INST_IN, true, false>;
    auto kern_n = dp_of2_kernel<InstT, R1, DP_INST_IN, false, true>;
    auto kern_mn = dp_of2_kernel<InstT, R1, DP_INST_IN, true, true>;
    *smem = K::SMEM_BYTES;
    *threads = K::NT;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kern_m, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kern_n, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kern_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int occ = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_m, K::NT, K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    *occ_out = occ;
    *grid_max = sms * occ;
    return 0;
}
// persist_bytes > 0: the launch carries an access-policy window that makes the scratch area (X column,
// parked block results) persisting in L2 and everything else streaming -- the set-aside was reserved by
// the plan (cudaLimitPersistingL2CacheSize)
template <int R1> int launch_one(const Dp2Params<InstT>& prm, int multi, int grid, size_t smem, cudaStream_t st, size_t persist_bytes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Dp2Geom<InstT, R1>::NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (persist_bytes > 0) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = prm.scratch;
        attr[0].val.accessPolicyWindow.num_bytes = persist_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaError_t e;
    // bit 0: some channel has more than one template; bit 1: every template's delay windows are narrow (column-wise scan)
    switch (multi & 3) {
        case 3: e = cudaLaunchKernelEx(&cfg, dp_of2_kernel<InstT, R1, DP_INST_IN, true, true>, prm); break;
        case 2: e = cudaLaunchKernelEx(&cfg, dp_of2_kernel<InstT, R1, DP_INST_IN, false, true>, prm); break;
        case 1: e = cudaLaunchKernelEx(&cfg, dp_of2_kernel<InstT, R1, DP_INST_IN, true, false>, prm); break;
        default: e = cudaLaunchKernelEx(&cfg, dp_of2_kernel<InstT, R1, DP_INST_IN, false, false>, prm); break;
    }
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}
}  // namespace

int DP_CAT(dp_of2_setup_p, DP_INST_PREC, DP_INST_IN)(int R1, int device, size_t* smem, int* grid_max, int* occ, int* threads) {
    switch (R1) {
        case 2: return setup_one<2>(device, smem, grid_max, occ, threads);
        case 4: return setup_one<4>(device, smem, grid_max, occ, threads);
        case 8: return setup_one<8>(device, smem, grid_max, occ, threads);
        default: return -1;
    }
}
int DP_CAT(dp_of2_launch_p, DP_INST_PREC, DP_INST_IN)(int R1, int multi, const void* prm_v, int grid, size_t smem, void* st_v, size_t persist_bytes) {
    const Dp2Params<InstT>& prm = *reinterpret_cast<const Dp2Params<InstT>*>(prm_v);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(st_v);
    switch (R1) {
        case 2: return launch_one<2>(prm, multi, grid, smem, st, persist_bytes);
        case 4: return launch_one<4>(prm, multi, grid, smem, st, persist_bytes);
        case 8: return launch_one<8>(prm, multi, grid, smem, st, persist_bytes);
        default: return -1;
    }
}

// ------------------------------------------------------------------------ PSD kernels
// (float64 / float32 / int16 traces; not for the 8-byte-aligned window instantiation)
#if DP_INST_IN <= 2
namespace {
template <int R1> int psd_setup_one(int device, size_t* smem, int* grid_max, long long* partial_per_cta) {
    using K = DpPsd2Kernel<InstT, R1, DP_INST_IN>;
    auto kern = dp_psd2_kernel<InstT, R1, DP_INST_IN>;
    *smem = K::SMEM_BYTES;
    *partial_per_cta = K::PARTIAL;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int occ = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K::NT, K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return -2;
    *grid_max = sms * occ;
    return 0;
}
template <int R1> int psd_launch_one(const DpPsd2Params<InstT>& prm, int grid, size_t smem, cudaStream_t st) {
    dp_psd2_kernel<InstT, R1, DP_INST_IN><<<grid, Dp2Geom<InstT, R1>::NT, smem, st>>>(prm);
    return (int)cudaGetLastError();
}
}  // namespace

int DP_CAT(dp_psd2_setup_p, DP_INST_PREC, DP_INST_IN)(int R1, int device, size_t* smem, int* grid_max, long long* partial_per_cta) {
    switch (R1) {
        case 2: return psd_setup_one<2>(device, smem, grid_max, partial_per_cta);
        case 4: return psd_setup_one<4>(device, smem, grid_max, partial_per_cta);
        case 8: return psd_setup_one<8>(device, smem, grid_max, partial_per_cta);
        default: return -1;
    }
}
int DP_CAT(dp_psd2_launch_p, DP_INST_PREC, DP_INST_IN)(int R1, const void* prm_v, int grid, size_t smem, void* st_v) {
    const DpPsd2Params<InstT>& prm = *reinterpret_cast<const DpPsd2Params<InstT>*>(prm_v);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(st_v);
    switch (R1) {
        case 2: return psd_launch_one<2>(prm, grid, smem, st);
        case 4: return psd_launch_one<4>(prm, grid, smem, st);
        case 8: return psd_launch_one<8>(prm, grid, smem, st);
        default: return -1;
    }
}
#endif
