// Kernel instantiations for connect_four.
#include "batch_kernels.cuh"
#include "rules_connect_four.cuh"
namespace b2s {
GameOps* make_ops_connect_four() { return new GameOpsT<ConnectFourRules>(); }
}  // namespace b2s
