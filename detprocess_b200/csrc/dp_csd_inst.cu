// One translation unit per (precision, channel count) of the noise-CSD kernel (nb_samples 16384 / 32768 / 65536).
// Build with -DDP_INST_PREC=0|1 (double | packed float) -DDP_INST_NCH=2..4.
#ifndef DP_INST_PREC
#error "DP_INST_PREC must be defined"
#endif
#ifndef DP_INST_NCH
#error "DP_INST_NCH must be defined"
#endif
#include <cuda_runtime.h>

#include "dp_csd_kernel.cuh"
#include "dp_csd_launch.hpp"

#if DP_INST_PREC == 0
using InstT = double;
#else
using InstT = f2;
#endif

#define DP_CAT_(a, b, c) a##b##_##c
#define DP_CAT(a, b, c) DP_CAT_(a, b, c)

namespace {
template <int R1> int setup_one(int device, size_t* smem, int* grid_max, long long* partial_per_comp, long long* scratch_per_cta, int* ncp) {
    using K = DpCsdKernel<InstT, R1, DP_INST_NCH>;
    auto kern = dp_csd_kernel<InstT, R1, DP_INST_NCH>;
    *smem = K::SMEM_BYTES;
    *partial_per_comp = K::PARTIAL;  // slots; the plan multiplies by the padded component count
    *ncp = K::NCP;
    *scratch_per_cta = K::scratch_v();
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
template <int R1> int launch_one(const DpCsdParams<InstT>& prm, int grid, size_t smem, cudaStream_t st) {
    dp_csd_kernel<InstT, R1, DP_INST_NCH><<<grid, Dp2Geom<InstT, R1>::NT, smem, st>>>(prm);
    return (int)cudaGetLastError();
}
}  // namespace

int DP_CAT(dp_csd_setup_p, DP_INST_PREC, DP_INST_NCH)(int R1, int device, size_t* smem, int* grid_max, long long* partial_per_comp,
                                                      long long* scratch_per_cta, int* ncp) {
    switch (R1) {
        case 2: return setup_one<2>(device, smem, grid_max, partial_per_comp, scratch_per_cta, ncp);
        case 4: return setup_one<4>(device, smem, grid_max, partial_per_comp, scratch_per_cta, ncp);
        case 8: return setup_one<8>(device, smem, grid_max, partial_per_comp, scratch_per_cta, ncp);
        default: return -1;
    }
}
int DP_CAT(dp_csd_launch_p, DP_INST_PREC, DP_INST_NCH)(int R1, const void* prm, int grid, size_t smem, void* stream) {
    const auto& p = *reinterpret_cast<const DpCsdParams<InstT>*>(prm);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (R1) {
        case 2: return launch_one<2>(p, grid, smem, st);
        case 4: return launch_one<4>(p, grid, smem, st);
        case 8: return launch_one<8>(p, grid, smem, st);
        default: return -1;
    }
}

#if DP_INST_PREC == 0 && DP_INST_NCH == 2
// the (precision independent) reduction kernel lives in exactly one translation unit
__global__ void dp_csd_reduce_kernel(const DpCsdReduceParams prm) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int comp = blockIdx.y;
    if (k < prm.nbins) {
        const long long l = (long long)prm.loc[k] * prm.ncp + comp;
        double s = 0.0;
        for (int c = 0; c < prm.grid; ++c) s += prm.partial[(long long)c * prm.partial_per_cta + l];
        prm.sum_out[(long long)comp * prm.nbins + k] += s;
    }
    if (k == 0 && comp == 0) {
        unsigned long long n = 0;
        for (int c = 0; c < prm.grid; ++c) n += prm.count[c];
        *prm.count_out += n;
    }
}
int dp_csd_reduce_launch(const void* prm_v, void* st_v) {
    const DpCsdReduceParams& prm = *reinterpret_cast<const DpCsdReduceParams*>(prm_v);
    dp_csd_reduce_kernel<<<dim3((prm.nbins + 255) / 256, prm.ncomp), 256, 0, reinterpret_cast<cudaStream_t>(st_v)>>>(prm);
    return (int)cudaGetLastError();
}
#endif
