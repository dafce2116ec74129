// Kernel instantiations for leduc_poker: the 8-byte two-player core and the 16-byte 3..4-player core.
#include "batch_kernels.cuh"
#include "rules_leduc_poker.cuh"
#include "rules_leduc_poker_n.cuh"
namespace b2s {
GameOps* make_ops_leduc_poker() { return new GameOpsT<LeducRules>(); }
GameOps* make_ops_leduc_poker_n() { return new GameOpsT<LeducNRules>(); }
}  // namespace b2s
