// Launch entry points of the trigger kernels (dp_trig_inst.cu).  p0 = float64, p1 = packed float32.
#pragma once
#include <cstddef>
int dp_trig_setup_p0(int R1, int device, size_t* smem, int* grid_max, long long* scratch_per_cta);
int dp_trig_setup_p1(int R1, int device, size_t* smem, int* grid_max, long long* scratch_per_cta);
int dp_trig_launch_p0(int R1, int in_dtype, const void* prm, int grid, size_t smem, void* stream);
int dp_trig_launch_p1(int R1, int in_dtype, const void* prm, int grid, size_t smem, void* stream);
int dp_trig_group_launch(const void* prm, void* stream);
int dp_trig_group_par_launch(const void* prm, int grid, void* stream);
int dp_trig_residual_launch(const void* prm, int grid, void* stream);
int dp_trig_flatten_launch(const void* group_prm, const void* prm, int grid, void* stream);
int dp_trig_filtered_at_launch(const void* prm, void* stream);
