// Kernel instantiations for havannah.
#include "batch_kernels.cuh"
#include "rules_havannah.cuh"
namespace b2s {
GameOps* make_ops_havannah() { return new GameOpsT<HavannahRules>(); }
}  // namespace b2s
