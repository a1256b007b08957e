// Host-visible launchers of the device code (implemented in reach_kernels.cu / constraint_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "armour_types.cuh"

namespace armour {
typedef unsigned long long u64;
struct FlatPZ { int n, dim, cap; u64* keys; double* coef; double* center; double* ind; };
struct FlatOut { int n, dim; };

cudaError_t upload_robot_model(const RobotModel& rm);
size_t arena_bytes(int mcap, int ncap, int groups);
size_t reach_gmem_bytes(int ncap);
int reach_max_ctas_per_sm(int nt, int minb, int groups, int scap, int tcap);
cudaError_t launch_reach_build(const Tables& tb, char* arena, size_t arena_stride, int mcap, int ncap, int scap, int tcap, int n_work, int grid, int nt, int minb, int groups, cudaStream_t stream);
cudaError_t launch_pz_binary(int op, const FlatPZ& a, const FlatPZ& b, const FlatPZ& r, FlatOut* out, char* gmem, int ncap, int scap, int tcap, double thr, int* err, cudaStream_t stream);
cudaError_t launch_hyperplanes(const Tables& tb, cudaStream_t stream);
int eval_max_obstacles();
cudaError_t launch_constraint_eval(const Tables& tb, int prob, const double* x_host, double* g, double* jac, double* link_center, int what, unsigned* done_counter,
                                   unsigned long long* done_flag, unsigned long long seq, int blocks_per_sm, cudaStream_t stream);
cudaError_t launch_constraint_eval_batch(const Tables& tb, int first, int count, const double* x_dev, double* g, double* jac, double* link_center, int what, cudaStream_t stream);
double measure_fp64_tflops(int sm_count);
void read_phase_cycles(unsigned long long* cycles, unsigned long long* calls, bool reset);
}  // namespace armour
