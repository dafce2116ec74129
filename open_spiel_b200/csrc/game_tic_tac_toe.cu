// Kernel instantiations for tic_tac_toe.
#include "batch_kernels.cuh"
#include "rules_tic_tac_toe.cuh"
namespace b2s {
GameOps* make_ops_tic_tac_toe() { return new GameOpsT<TicTacToeRules>(); }
}  // namespace b2s
