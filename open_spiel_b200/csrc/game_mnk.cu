// Kernel instantiations for mnk.
#include "batch_kernels.cuh"
#include "rules_mnk.cuh"
namespace b2s {
GameOps* make_ops_mnk() { return new GameOpsT<MnkRules>(); }
}  // namespace b2s
