// Games whose device rule cores are not built yet return no ops table (b2s_batch_create fails loudly).
#include "batch_kernels.cuh"
namespace b2s {
GameOps* make_ops_go() { return nullptr; }
}  // namespace b2s
