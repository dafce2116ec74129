// Kernel instantiations for hex.
#include "batch_kernels.cuh"
#include "rules_hex.cuh"
namespace b2s {
GameOps* make_ops_hex() { return new GameOpsT<HexRules>(); }
}  // namespace b2s
