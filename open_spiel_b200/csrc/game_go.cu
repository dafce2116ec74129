// Kernel instantiations for go.
#include "batch_kernels.cuh"
#include "rules_go.cuh"
namespace b2s {
GameOps* make_ops_go() { return new GameOpsT<GoRules>(); }
}  // namespace b2s
