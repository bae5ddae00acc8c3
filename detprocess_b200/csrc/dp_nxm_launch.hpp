// Launch entry points of the NxM optimal-filter kernel instantiation units (dp_nxm_inst.cu).
// p0 = float64, p1 = packed float32; second index = channel count 1..4.
#pragma once
#include <cstddef>

#define DP_NXM_DECL(P, C)                                                                                   \
    int dp_nxm_setup_p##P##_##C(int R1, int device, size_t* smem, int* grid_max, int* threads);               \
    int dp_nxm_launch_p##P##_##C(int R1, const void* prm, int grid, size_t smem, void* stream);               \
    long long dp_nxm_scratch_p##P##_##C(int R1, int n_chan, int n_templ);
DP_NXM_DECL(0, 1) DP_NXM_DECL(0, 2) DP_NXM_DECL(0, 3) DP_NXM_DECL(0, 4)
DP_NXM_DECL(1, 1) DP_NXM_DECL(1, 2) DP_NXM_DECL(1, 3) DP_NXM_DECL(1, 4)
#undef DP_NXM_DECL

typedef int (*dp_nxm_setup_fn)(int, int, size_t*, int*, int*);
typedef int (*dp_nxm_launch_fn)(int, const void*, int, size_t, void*);
typedef long long (*dp_nxm_scratch_fn)(int, int, int);
static const dp_nxm_setup_fn dp_nxm_setup_table[2][4] = {{dp_nxm_setup_p0_1, dp_nxm_setup_p0_2, dp_nxm_setup_p0_3, dp_nxm_setup_p0_4},
                                                         {dp_nxm_setup_p1_1, dp_nxm_setup_p1_2, dp_nxm_setup_p1_3, dp_nxm_setup_p1_4}};
static const dp_nxm_launch_fn dp_nxm_launch_table[2][4] = {{dp_nxm_launch_p0_1, dp_nxm_launch_p0_2, dp_nxm_launch_p0_3, dp_nxm_launch_p0_4},
                                                           {dp_nxm_launch_p1_1, dp_nxm_launch_p1_2, dp_nxm_launch_p1_3, dp_nxm_launch_p1_4}};
static const dp_nxm_scratch_fn dp_nxm_scratch_table[2][4] = {
    {dp_nxm_scratch_p0_1, dp_nxm_scratch_p0_2, dp_nxm_scratch_p0_3, dp_nxm_scratch_p0_4},
    {dp_nxm_scratch_p1_1, dp_nxm_scratch_p1_2, dp_nxm_scratch_p1_3, dp_nxm_scratch_p1_4}};
