// Translation unit of the band-amplitude kernel (dp_band_kernel.cuh).
#include <cuda_runtime.h>
#define DP_BAND_DEFINE_KERNEL 1
#include "dp_band_kernel.cuh"
#include "dp_band_launch.hpp"

int dp_band_launch(const void* prm_v, int grid, void* st_v) {
    const DpBandParams& prm = *reinterpret_cast<const DpBandParams*>(prm_v);
    dp_band_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(st_v)>>>(prm);
    return (int)cudaGetLastError();
}
