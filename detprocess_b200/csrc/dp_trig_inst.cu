// Translation unit of the continuous-stream trigger kernels (dp_trig_kernel.cuh), one per
// precision:  -DDP_INST_PREC=0|1 (double | packed float).  Streams: float64, float32 or int16.
#ifndef DP_INST_PREC
#error "DP_INST_PREC must be defined"
#endif
#include <cuda_runtime.h>

#if DP_INST_PREC == 0
#define DP_TRIG_DEFINE_GROUP_KERNEL 1
#endif
#include "dp_trig_kernel.cuh"
#include "dp_trig_launch.hpp"

#if DP_INST_PREC == 0
using InstT = double;
#else
using InstT = f2;
#endif

#define DP_CAT_(a, b) a##b
#define DP_CAT(a, b) DP_CAT_(a, b)

namespace {
template <int R1, int IN> cudaError_t prep_one(int* occ) {
    using K = DpTrigKernel<InstT, R1, IN>;
    auto kern = dp_trig_filter_kernel<InstT, R1, IN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, K::NT, K::SMEM_BYTES);
    if (o < *occ) *occ = o;
    return e;
}
template <int R1> int setup_one(int device, size_t* smem, int* grid_max, long long* scratch_per_cta) {
    using K = DpTrigKernel<InstT, R1, 0>;
    *smem = K::SMEM_BYTES;
    *scratch_per_cta = K::SCR_PARK;
    int occ = 1 << 20, sms = 0;
    cudaError_t e = prep_one<R1, 0>(&occ);
    if (e != cudaSuccess) return (int)e;
    e = prep_one<R1, 1>(&occ);
    if (e != cudaSuccess) return (int)e;
    e = prep_one<R1, 2>(&occ);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return -2;
    *grid_max = sms * occ;
    return 0;
}
template <int R1> int launch_one(const DpTrigParams<InstT>& prm, int in_dtype, int grid, size_t smem, cudaStream_t st) {
    switch (in_dtype) {
        case 0: dp_trig_filter_kernel<InstT, R1, 0><<<grid, Dp2Geom<InstT, R1>::NT, smem, st>>>(prm); break;
        case 1: dp_trig_filter_kernel<InstT, R1, 1><<<grid, Dp2Geom<InstT, R1>::NT, smem, st>>>(prm); break;
        case 2: dp_trig_filter_kernel<InstT, R1, 2><<<grid, Dp2Geom<InstT, R1>::NT, smem, st>>>(prm); break;
        default: return -1;
    }
    return (int)cudaGetLastError();
}
}  // namespace

int DP_CAT(dp_trig_setup_p, DP_INST_PREC)(int R1, int device, size_t* smem, int* grid_max, long long* scratch_per_cta) {
    switch (R1) {
        case 2: return setup_one<2>(device, smem, grid_max, scratch_per_cta);
        case 4: return setup_one<4>(device, smem, grid_max, scratch_per_cta);
        case 8: return setup_one<8>(device, smem, grid_max, scratch_per_cta);
        default: return -1;
    }
}
int DP_CAT(dp_trig_launch_p, DP_INST_PREC)(int R1, int in_dtype, const void* prm_v, int grid, size_t smem, void* st_v) {
    const DpTrigParams<InstT>& prm = *reinterpret_cast<const DpTrigParams<InstT>*>(prm_v);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(st_v);
    switch (R1) {
        case 2: return launch_one<2>(prm, in_dtype, grid, smem, st);
        case 4: return launch_one<4>(prm, in_dtype, grid, smem, st);
        case 8: return launch_one<8>(prm, in_dtype, grid, smem, st);
        default: return -1;
    }
}
#if DP_INST_PREC == 0
int dp_trig_group_launch(const void* prm_v, void* st_v) {
    const DpTrigGroupParams& prm = *reinterpret_cast<const DpTrigGroupParams*>(prm_v);
    dp_trig_group_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(st_v)>>>(prm);
    return (int)cudaGetLastError();
}
// the same grouping from multi-CTA kernels (grid: CTAs for the per-candidate kernels)
int dp_trig_group_par_launch(const void* prm_v, int grid, void* st_v) {
    const DpTrigGroupParams& prm = *reinterpret_cast<const DpTrigGroupParams*>(prm_v);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(st_v);
    cudaError_t e = cudaMemsetAsync(prm.best_key, 0, sizeof(unsigned long long) * (size_t)prm.max_triggers, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(prm.best_g, 0xff, sizeof(unsigned long long) * (size_t)prm.max_triggers, st);
    if (e != cudaSuccess) return (int)e;
    dp_trig_par_offsets_kernel<<<1, 1024, 0, st>>>(prm);
    dp_trig_par_heads_kernel<<<grid, 1024, 0, st>>>(prm);
    dp_trig_par_scan_kernel<<<1, 1024, 0, st>>>(prm);
    dp_trig_par_best_kernel<0><<<grid, 1024, 0, st>>>(prm);
    dp_trig_par_best_kernel<1><<<grid, 1024, 0, st>>>(prm);
    dp_trig_par_emit_kernel<<<64, 256, 0, st>>>(prm);
    return (int)cudaGetLastError();
}
// residual pass: residuals + in-place compaction of the candidate list (then dp_trig_group_par_launch with cand_val)
int dp_trig_residual_launch(const void* prm_v, int grid, void* st_v) {
    const DpTrigResidParams& prm = *reinterpret_cast<const DpTrigResidParams*>(prm_v);
    dp_trig_residual_kernel<<<grid, 1024, 0, reinterpret_cast<cudaStream_t>(st_v)>>>(prm);
    return (int)cudaGetLastError();
}
// chunk offsets of the current candidate counts, then the flat ordered list
int dp_trig_flatten_launch(const void* group_prm_v, const void* prm_v, int grid, void* st_v) {
    const DpTrigGroupParams& gp = *reinterpret_cast<const DpTrigGroupParams*>(group_prm_v);
    const DpTrigFlattenParams& prm = *reinterpret_cast<const DpTrigFlattenParams*>(prm_v);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(st_v);
    dp_trig_par_offsets_kernel<<<1, 1024, 0, st>>>(gp);
    dp_trig_flatten_kernel<<<grid, 256, 0, st>>>(prm);
    return (int)cudaGetLastError();
}
int dp_trig_filtered_at_launch(const void* prm_v, void* st_v) {
    const DpTrigAtParams& prm = *reinterpret_cast<const DpTrigAtParams*>(prm_v);
    dp_trig_filtered_at_kernel<<<prm.n_idx < 1024 ? prm.n_idx : 1024, 256, 0, reinterpret_cast<cudaStream_t>(st_v)>>>(prm);
    return (int)cudaGetLastError();
}
#endif
