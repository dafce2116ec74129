// Kernel instantiations for kuhn_poker.
#include "batch_kernels.cuh"
#include "rules_kuhn_poker.cuh"
namespace b2s {
GameOps* make_ops_kuhn_poker() { return new GameOpsT<KuhnRules>(); }
}  // namespace b2s
