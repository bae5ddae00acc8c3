// One translation unit per (precision, channel count) of the fused NxM optimal-filter kernel (nb_samples 16384 / 32768 /
// 65536).  Build with -DDP_INST_PREC=0|1 (double | packed float) -DDP_INST_NCH=1..4.
#ifndef DP_INST_PREC
#error "DP_INST_PREC must be defined"
#endif
#ifndef DP_INST_NCH
#error "DP_INST_NCH must be defined"
#endif
#include <cuda_runtime.h>

#include "dp_nxm_kernel.cuh"
#include "dp_nxm_launch.hpp"

#if DP_INST_PREC == 0
using InstT = double;
#else
using InstT = f2;
#endif

#define DP_CAT_(a, b, c) a##b##_##c
#define DP_CAT(a, b, c) DP_CAT_(a, b, c)

namespace {
template <int R1> int setup_one(int device, size_t* smem, int* grid_max, int* threads) {
    using K = DpNxmKernel<InstT, R1, DP_INST_NCH>;
    auto kern = dp_nxm_kernel<InstT, R1, DP_INST_NCH>;
    *smem = K::SMEM_BYTES;
    *threads = K::NT;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    int occ = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K::NT, K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return (int)cudaErrorLaunchOutOfResources;
    *grid_max = sms * occ;
    return 0;
}
template <int R1> int launch_one(const DpNxmParams<InstT>& prm, int grid, size_t smem, cudaStream_t st) {
    dp_nxm_kernel<InstT, R1, DP_INST_NCH><<<grid, Dp2Geom<InstT, R1>::NT, smem, st>>>(prm);
    return (int)cudaGetLastError();
}
}  // namespace

int DP_CAT(dp_nxm_setup_p, DP_INST_PREC, DP_INST_NCH)(int R1, int device, size_t* smem, int* grid_max, int* threads) {
    switch (R1) {
        case 2: return setup_one<2>(device, smem, grid_max, threads);
        case 4: return setup_one<4>(device, smem, grid_max, threads);
        case 8: return setup_one<8>(device, smem, grid_max, threads);
        default: return -1;
    }
}
int DP_CAT(dp_nxm_launch_p, DP_INST_PREC, DP_INST_NCH)(int R1, const void* prm, int grid, size_t smem, void* stream) {
    const auto& p = *reinterpret_cast<const DpNxmParams<InstT>*>(prm);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (R1) {
        case 2: return launch_one<2>(p, grid, smem, st);
        case 4: return launch_one<4>(p, grid, smem, st);
        case 8: return launch_one<8>(p, grid, smem, st);
        default: return -1;
    }
}
long long DP_CAT(dp_nxm_scratch_p, DP_INST_PREC, DP_INST_NCH)(int R1, int n_chan, int n_templ) {
    switch (R1) {
        case 2: return DpNxmKernel<InstT, 2, DP_INST_NCH>::scratch_v(n_chan, n_templ);
        case 4: return DpNxmKernel<InstT, 4, DP_INST_NCH>::scratch_v(n_chan, n_templ);
        case 8: return DpNxmKernel<InstT, 8, DP_INST_NCH>::scratch_v(n_chan, n_templ);
        default: return -1;
    }
}
