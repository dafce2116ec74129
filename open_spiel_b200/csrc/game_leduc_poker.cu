// Kernel instantiations for leduc_poker.
#include "batch_kernels.cuh"
#include "rules_leduc_poker.cuh"
namespace b2s {
GameOps* make_ops_leduc_poker() { return new GameOpsT<LeducRules>(); }
}  // namespace b2s
