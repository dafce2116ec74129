// Kernel instantiations for breakthrough.
#include "batch_kernels.cuh"
#include "rules_breakthrough.cuh"
namespace b2s {
GameOps* make_ops_breakthrough() { return new GameOpsT<BreakthroughRules>(); }
}  // namespace b2s
