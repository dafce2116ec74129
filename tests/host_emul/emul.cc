// TEST INFRASTRUCTURE ONLY — never part of the product (libb2s.so has no CPU path and fails without a GPU).
//
// Compiles the PRODUCT's per-game rule cores (open_spiel_b200/csrc/rules_*.cuh — the exact source the CUDA kernels are
// built from) for the host with g++, so that their bit-twiddling (multiply-gathers, flood fills, Zobrist / superko,
// packed poker histories, observation packing) is unit-tested by the CPU suite lock-step against the oracle
// (tests/test_rule_cores_host.py) before any GPU time is spent.  CUDA's host_defines.h makes __device__ /
// __forceinline__ harmless under a plain host compiler; the few device intrinsics the cores use get host definitions
// below.  The loops here restate what the generic kernels of batch_kernels.cuh do per lane (k_reset, k_apply,
// k_legal_mask, k_status, k_obs); launch geometry, coalescing and shared-memory staging are NOT exercised here — that is
// what the -m gpu tests are for.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <type_traits>
#include <vector>

#include "../../open_spiel_b200/csrc/host_compat.h"   // host definitions of the device intrinsics (product header)
// one "thread" at a time: the kernels index with blockIdx.x * blockDim.x + threadIdx.x
static struct { unsigned x, y, z; } blockIdx, blockDim = {1, 1, 1}, threadIdx;

#include "../../open_spiel_b200/csrc/common.cuh"
#include "../../open_spiel_b200/csrc/rules_tic_tac_toe.cuh"
#include "../../open_spiel_b200/csrc/rules_connect_four.cuh"
#include "../../open_spiel_b200/csrc/rules_breakthrough.cuh"
#include "../../open_spiel_b200/csrc/rules_hex.cuh"
#include "../../open_spiel_b200/csrc/rules_go.cuh"
#include "../../open_spiel_b200/csrc/rules_kuhn_poker.cuh"
#include "../../open_spiel_b200/csrc/rules_leduc_poker.cuh"
#include "../../open_spiel_b200/csrc/rules_leduc_poker_n.cuh"
#include "../../open_spiel_b200/csrc/rules_mnk.cuh"
#include "../../open_spiel_b200/csrc/rules_othello.cuh"
#include "../../open_spiel_b200/csrc/rules_y.cuh"
#include "../../open_spiel_b200/csrc/rules_havannah.cuh"
#include "../../open_spiel_b200/csrc/mcts.cuh"

namespace {
using namespace b2s;

template <class R> auto call_init(int) -> decltype(R::device_init(), void()) { R::device_init(); }
template <class R> void call_init(long) {}

struct Emu {
  virtual ~Emu() {}
  virtual void reset(long long n) = 0;
  virtual void apply(const int* a, long long n) = 0;
  virtual void legal(u32* out, long long n) = 0;
  virtual void status(signed char* cur, unsigned char* term, float* rets, long long n) = 0;
  virtual int obs(int player, int which, float* out, long long n) = 0;
  virtual void rollout(u64 seed, long long lane_offset, float* rets, int* plies, long long n) = 0;
  virtual int mcts(long long n, const b2s_mcts_config& mc, int* visits, double* reward, float* outcome, int* best, int* sims_run) = 0;
  b2s_game_info info;
  ErrBuf err;
};

template <class R>
struct EmuT : Emu {
  typename R::Cfg cfg;
  std::vector<char> planes;
  std::vector<u64> hist;
  long long cap = 0;
  Ctx ctx() { Ctx c; c.planes = planes.data(); c.cap = cap; c.hist = hist.empty() ? nullptr : hist.data(); c.err = &err; return c; }
  const char* configure(const b2s_params& p, long long capacity) {
    memset(&info, 0, sizeof info);
    const char* e = R::make_cfg(p, cfg, info);
    if (e) return e;
    int width = info.num_distinct_actions > info.max_chance_outcomes ? info.num_distinct_actions : info.max_chance_outcomes;
    info.mask_words = (width + 31) / 32;                    // as GameOpsT<R>::configure (batch_kernels.cuh)
    if (info.mask_words > R::kMaskWords) return "action space too large for the device path";
    info.state_bytes = (int)(sizeof(typename R::Chunk) * R::kChunks);
    info.game_id = R::kGameId;
    cap = capacity;
    planes.assign(sizeof(typename R::Chunk) * R::kChunks * (size_t)cap, 0);
    if (info.history_bytes) hist.assign((size_t)info.history_bytes / sizeof(u64) * (size_t)cap, 0);
    call_init<R>(0);
    return nullptr;
  }
  void reset(long long n) override {
    err.count = 0; err.first = 0x7fffffffffffffffLL;
    Ctx c = ctx();
    for (long long i = 0; i < n; ++i) { typename R::S s; R::init(s, cfg, c, i); R::store(s, c, i); }
  }
  void apply(const int* a, long long n) override {                      // k_apply
    Ctx c = ctx();
    for (long long i = 0; i < n; ++i) {
      if (a[i] == -1) continue;
      typename R::S s;
      R::load(s, c, i);
      if (R::terminal(s, cfg) || !R::apply(s, a[i], cfg, c, i)) { flag_error(&err, i); continue; }
      R::store(s, c, i);
    }
  }
  void legal(u32* out, long long n) override {                           // k_legal_mask
    Ctx c = ctx();
    for (long long i = 0; i < n; ++i) {
      typename R::S s;
      R::load(s, c, i);
      u32 m[R::kMaskWords];
      R::legal(s, cfg, m);
      for (int w = 0; w < info.mask_words; ++w) out[i * info.mask_words + w] = m[w];
    }
  }
  void status(signed char* cur, unsigned char* term, float* rets, long long n) override {   // k_status
    Ctx c = ctx();
    for (long long i = 0; i < n; ++i) {
      typename R::S s;
      R::load(s, c, i);
      int cp = R::cur_player(s, cfg);
      cur[i] = (signed char)cp;
      term[i] = cp == kTerminalPlayerId ? 1 : 0;
      float r[R::kPlayers];
      R::returns(s, cfg, r);
      for (int p = 0; p < info.num_players; ++p) rets[i * info.num_players + p] = r[p];
    }
  }
  int obs(int player, int which, float* out, long long n) override {     // k_obs: obs_pack, then obs_elem per element
    int size = which == 0 ? info.observation_tensor_size : info.information_state_tensor_size;
    if (size <= 0 || (which == 1 && !R::kHasInfoState)) return 1;
    Ctx c = ctx();
    for (long long i = 0; i < n; ++i) {
      typename R::S s;
      R::load(s, c, i);
      int pl = player;
      if (pl < 0) { pl = R::cur_player(s, cfg); if (pl < 0) pl = 0; }
      typename R::ObsPack pk;
      R::obs_pack(s, cfg, pl, which, pk);
      for (int e = 0; e < size; ++e) out[i * size + e] = R::obs_elem(pk, cfg, e);
    }
    return 0;
  }
  void rollout(u64 seed, long long lane_offset, float* rets, int* plies, long long n) override {   // k_rollout
    Ctx c = ctx();
    for (long long i = 0; i < n; ++i) {
      typename R::S s;
      R::load(s, c, i);
      int ply = 0;
      while (!R::terminal(s, cfg) && ply < info.max_game_length + 4) {
        auto draw = [&](u32 b, u32 m) { return philox_uniform(seed, (u64)(i + lane_offset), b, m); };
        playout_step<R>(s, cfg, c, i, info.mask_words, draw, (u32)ply);
        ++ply;
      }
      R::store(s, c, i);
      plies[i] = ply;
      float r[R::kPlayers];
      R::returns(s, cfg, r);
      for (int p = 0; p < info.num_players; ++p) rets[i * info.num_players + p] = r[p];
    }
  }
  // b2s_mcts_search: the argument block is filled as api.cu / GameOpsT<R>::mcts do, then the KERNEL BODY of mcts.cuh is
  // executed once per tree with blockIdx.x = tree (one thread = one tree on the device as well).
  int mcts(long long n, const b2s_mcts_config& mc, int* visits, double* reward, float* outcome, int* best, int* sims_run) override {
    return mcts_impl(n, mc, visits, reward, outcome, best, sims_run, std::integral_constant<bool, (R::kMaxPath > 0)>());
  }
  int mcts_impl(long long, const b2s_mcts_config&, int*, double*, float*, int*, int*, std::false_type) { return 1; }
  int mcts_impl(long long n, const b2s_mcts_config& mc, int* visits, double* reward, float* outcome, int* best, int* sims_run,
                std::true_type) {
    if (info.max_game_length + 2 > R::kMaxPath) return 2;
    std::vector<char> wplanes(sizeof(typename R::Chunk) * R::kChunks * (size_t)n, 0);
    std::vector<u64> whist(info.history_bytes ? (size_t)info.history_bytes / sizeof(u64) * (size_t)n : 0, 0);
    Ctx work;
    work.planes = wplanes.data(); work.cap = n; work.hist = whist.empty() ? nullptr : whist.data(); work.err = &err;
    {                                                       // api.cu: B->ops->copy(work, 0, roots, 0, n) — k_copy
      Ctx src = ctx();
      for (long long i = 0; i < n; ++i) {
        typename R::S s;
        R::load(s, src, i);
        R::store(s, work, i);
        R::copy_history(work, i, src, i, s, cfg);
      }
    }
    std::vector<double> logt((size_t)mc.max_simulations + 2, 0.0);
    for (size_t k = 1; k < logt.size(); ++k) logt[k] = std::log((double)k);
    // arena sizing as api.cu (b2s_mcts_search)
    const unsigned long long A = (unsigned long long)info.num_distinct_actions;
    const int compact = (mc.n_rollouts & (mc.n_rollouts - 1)) == 0 && (long long)mc.max_simulations * mc.n_rollouts < (1ll << 30);
    unsigned long long per_tree;
    if (mc.max_nodes_total > 0) per_tree = (unsigned long long)mc.max_nodes_total / (unsigned long long)n;
    else {
      per_tree = 2ull + (unsigned long long)mc.max_simulations * A;
      if (mc.max_nodes_per_tree > 1) {
        unsigned long long want = 2ull * (unsigned long long)mc.max_nodes_per_tree + 8 * A + 64;
        if (want < per_tree) per_tree = want;
      }
    }
    const size_t node_bytes = compact ? sizeof(MctsNodeC) : sizeof(MctsNodeW);
    std::vector<char> pool((size_t)per_tree * (size_t)n * node_bytes + 16);
    unsigned long long used = 0;
    std::vector<int> gc(n, 0);
    MctsArgs a;
    memset(&a, 0, sizeof a);
    a.sims = mc.max_simulations; a.n_rollouts = mc.n_rollouts; a.solve = mc.solve; a.uct_c = mc.uct_c;
    a.puct = mc.child_selection_policy == B2S_MCTS_PUCT;
    a.max_nodes = (int)mc.max_nodes_per_tree; a.max_seconds = 0;
    a.seed = mc.seed; a.tree_offset = mc.tree_index_offset; a.log_table = logt.data();
    a.pool = (void*)(((uintptr_t)pool.data() + 15) & ~(uintptr_t)15); a.nodes_per_tree = per_tree; a.nodes_used = &used; a.compact = compact;
    a.visits_out = visits; a.reward_out = reward; a.outcome_out = outcome; a.best_out = best; a.sims_out = sims_run;
    a.gc_out = mc.gc_runs_d ? mc.gc_runs_d : gc.data(); a.err = &err;
    a.num_actions = info.num_distinct_actions; a.mask_words = info.mask_words;
    a.max_plies = info.max_game_length + 4; a.max_utility = info.max_utility;
    blockDim.x = 1; threadIdx.x = 0;
    for (long long t = 0; t < n; ++t) {
      blockIdx.x = (unsigned)t;
      if (compact) k_mcts<R, StatsC, R::kMaxPath, 4>(ctx(), work, cfg, a, n);
      else k_mcts<R, StatsW, R::kMaxPath, 4>(ctx(), work, cfg, a, n);
    }
    return 0;
  }
};

std::string g_err;
template <class R>
Emu* make(const b2s_params& p, long long cap) {
  auto* e = new EmuT<R>();
  const char* msg = e->configure(p, cap);
  if (msg) { g_err = msg; delete e; return nullptr; }
  return e;
}
}  // namespace

extern "C" {
const char* emu_last_error() { return g_err.c_str(); }
void emu_params_default(b2s_params* p) {                     // "unset" = -1 / NaN, as b2s_params_default
  memset(p, 0xff, sizeof *p);
  p->komi = __builtin_nan("");
  for (double& d : p->reserved_d) d = __builtin_nan("");
}
void* emu_create(int game_id, const b2s_params* p, long long cap) {
  switch (game_id) {
    case B2S_TIC_TAC_TOE: return make<TicTacToeRules>(*p, cap);
    case B2S_CONNECT_FOUR: return make<ConnectFourRules>(*p, cap);
    case B2S_BREAKTHROUGH: return make<BreakthroughRules>(*p, cap);
    case B2S_HEX: return make<HexRules>(*p, cap);
    case B2S_GO: return make<GoRules>(*p, cap);
    case B2S_KUHN_POKER: return make<KuhnRules>(*p, cap);
    case B2S_MNK: return make<MnkRules>(*p, cap);
    case B2S_OTHELLO: return make<OthelloRules>(*p, cap);
    case B2S_Y: return make<YRules>(*p, cap);
    case B2S_HAVANNAH: return make<HavannahRules>(*p, cap);
    case B2S_LEDUC_POKER: return p->players > 2 ? make<LeducNRules>(*p, cap) : make<LeducRules>(*p, cap);
  }
  g_err = "unknown game id";
  return nullptr;
}
void emu_destroy(void* h) { delete (Emu*)h; }
void emu_info(void* h, b2s_game_info* out) { *out = ((Emu*)h)->info; }
void emu_reset(void* h, long long n) { ((Emu*)h)->reset(n); }
void emu_apply(void* h, const int* a, long long n) { ((Emu*)h)->apply(a, n); }
void emu_legal_mask(void* h, uint32_t* out, long long n) { ((Emu*)h)->legal(out, n); }
void emu_status(void* h, signed char* cur, unsigned char* term, float* rets, long long n) { ((Emu*)h)->status(cur, term, rets, n); }
int emu_observation(void* h, int player, int which, float* out, long long n) { return ((Emu*)h)->obs(player, which, out, n); }
long long emu_error_count(void* h) { return (long long)((Emu*)h)->err.count; }
void emu_rollout(void* h, unsigned long long seed, long long lane_offset, float* rets, int* plies, long long n) {
  ((Emu*)h)->rollout(seed, lane_offset, rets, plies, n);
}
int emu_mcts(void* h, long long n, const b2s_mcts_config* mc, int* visits, double* reward, float* outcome, int* best, int* sims_run) {
  return ((Emu*)h)->mcts(n, *mc, visits, reward, outcome, best, sims_run);
}
}
