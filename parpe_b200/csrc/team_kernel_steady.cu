// The large-model kernel with steady-state support (simulate_team_kernel<512, true>: pre-equilibration, t = inf
// output times): instantiated here, once, for every network model.
#include "team_kernels.cuh"
namespace pb200 {
int team_launch_kernel_steady(const KernelArgs& a, const TeamState& st, int grid, int use_sparse, cudaStream_t stream) {
    return team_launch_kernel<true>(a, st, grid, use_sparse, stream);
}
}  // namespace pb200
