// The large-model kernel without steady-state support (simulate_team_kernel<512, false>): instantiated here, once,
// for every network model (the kernel is generic over the tables).
#include "team_kernels.cuh"
namespace pb200 {
int team_launch_kernel_plain(const KernelArgs& a, const TeamState& st, int grid, int use_sparse, cudaStream_t stream) {
    return team_launch_kernel<false>(a, st, grid, use_sparse, stream);
}
}  // namespace pb200
