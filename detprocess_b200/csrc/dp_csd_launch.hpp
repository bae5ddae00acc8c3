// Launch entry points of the noise-CSD kernel instantiation units (dp_csd_inst.cu).
// p0 = float64, p1 = packed float32; second index = channel count 2..4.
#pragma once
#include <cstddef>

#define DP_CSD_DECL(P, C)                                                                                              \
    int dp_csd_setup_p##P##_##C(int R1, int device, size_t* smem, int* grid_max, long long* partial_per_comp,            \
                                long long* scratch_per_cta, int* ncp);                                                 \
    int dp_csd_launch_p##P##_##C(int R1, const void* prm, int grid, size_t smem, void* stream);
DP_CSD_DECL(0, 2) DP_CSD_DECL(0, 3) DP_CSD_DECL(0, 4) DP_CSD_DECL(1, 2) DP_CSD_DECL(1, 3) DP_CSD_DECL(1, 4)
#undef DP_CSD_DECL
int dp_csd_reduce_launch(const void* prm, void* stream);

typedef int (*dp_csd_setup_fn)(int, int, size_t*, int*, long long*, long long*, int*);
typedef int (*dp_csd_launch_fn)(int, const void*, int, size_t, void*);
static const dp_csd_setup_fn dp_csd_setup_table[2][3] = {{dp_csd_setup_p0_2, dp_csd_setup_p0_3, dp_csd_setup_p0_4},
                                                         {dp_csd_setup_p1_2, dp_csd_setup_p1_3, dp_csd_setup_p1_4}};
static const dp_csd_launch_fn dp_csd_launch_table[2][3] = {{dp_csd_launch_p0_2, dp_csd_launch_p0_3, dp_csd_launch_p0_4},
                                                           {dp_csd_launch_p1_2, dp_csd_launch_p1_3, dp_csd_launch_p1_4}};
