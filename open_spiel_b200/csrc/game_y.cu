// Kernel instantiations for y.
#include "batch_kernels.cuh"
#include "rules_y.cuh"
namespace b2s {
GameOps* make_ops_y() { return new GameOpsT<YRules>(); }
}  // namespace b2s
