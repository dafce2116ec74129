// Kernel instantiations for othello.
#include "batch_kernels.cuh"
#include "rules_othello.cuh"
namespace b2s {
GameOps* make_ops_othello() { return new GameOpsT<OthelloRules>(); }
}  // namespace b2s
