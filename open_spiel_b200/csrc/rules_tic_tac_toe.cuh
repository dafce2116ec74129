// tic_tac_toe rule core.  Semantics: reference open_spiel/games/tic_tac_toe/tic_tac_toe.cc
// (DoApplyAction :128-136, LegalActions :138-148, BoardHasLine :108-120, IsTerminal :215-217,
// Returns :219-227, ObservationTensor :241-251).  Packed: one uint32 — bits 0-8 player 0 ("x", kCross),
// bits 9-17 player 1 ("o", kNought); everything else derived.
#pragma once
#include "common.cuh"

namespace b2s {

struct TicTacToeRules {
  static constexpr int kGameId = B2S_TIC_TAC_TOE;
  typedef u32 Chunk;
  static constexpr int kChunks = 1;
  static constexpr int kMaskWords = 1;
  static constexpr int kObsWords = 1;
  static constexpr int kPlayers = 2;
  static constexpr int kMaxPath = 16;   // MCTS path stack (>= max_game_length + 2); 0 = no device MCTS
  static constexpr int kMaxLegal = 9;   // most legal actions any state can have (MCTS children block size)
  static constexpr int kFilterWords = 0;   // no per-lane history filter (see rules_go.cuh)
  static constexpr int kIlp = 4;      // lanes per thread in the streaming kernels
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = false;
  struct Cfg { int dummy; };
  struct S { u32 b; };

  static __host__ const char* make_cfg(const b2s_params&, Cfg& c, b2s_game_info& gi) {
    c.dummy = 0;
    gi.num_players = 2;
    gi.num_distinct_actions = 9;       // tic_tac_toe.h:136
    gi.max_game_length = 9;            // tic_tac_toe.h:157
    gi.observation_tensor_size = 27;
    gi.obs_shape[0] = 3; gi.obs_shape[1] = 3; gi.obs_shape[2] = 3;
    gi.min_utility = -1; gi.max_utility = 1;
    return nullptr;
  }
  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) { s.b = reinterpret_cast<const u32*>(ctx.planes)[i]; }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) { reinterpret_cast<u32*>(ctx.planes)[i] = s.b; }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) { s.b = 0; }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ bool line(u32 m) {   // m: 9-bit board of one player
    // rows 0x007,0x038,0x1C0; cols 0x049,0x092,0x124; diagonals 0x111,0x054
    return ((m & 0x007) == 0x007) | ((m & 0x038) == 0x038) | ((m & 0x1C0) == 0x1C0) | ((m & 0x049) == 0x049) |
           ((m & 0x092) == 0x092) | ((m & 0x124) == 0x124) | ((m & 0x111) == 0x111) | ((m & 0x054) == 0x054);
  }
  __device__ static __forceinline__ u32 xs(const S& s) { return s.b & 0x1ff; }
  __device__ static __forceinline__ u32 os(const S& s) { return (s.b >> 9) & 0x1ff; }
  __device__ static __forceinline__ int mover(const S& s) { return __popc(s.b & 0x3ffff) & 1; }
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) {
    return line(xs(s)) || line(os(s)) || ((xs(s) | os(s)) == 0x1ff);
  }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg& c) { return terminal(s, c) ? kTerminalPlayerId : mover(s); }
  __device__ static __forceinline__ void returns(const S& s, const Cfg&, float* r) {
    if (line(xs(s))) { r[0] = 1.f; r[1] = -1.f; }
    else if (line(os(s))) { r[0] = -1.f; r[1] = 1.f; }
    else { r[0] = 0.f; r[1] = 0.f; }
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg&, u32* m) { m[0] = ~(xs(s) | os(s)) & 0x1ff; }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (terminal(s, c)) m[0] = 0; else legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg&, const Ctx&, long long) {
    if (a < 0 || a >= 9) return false;
    if (((xs(s) | os(s)) >> a) & 1u) return false;
    s.b |= 1u << (a + 9 * mover(s));
    return true;
  }
  static constexpr bool kObsBitPacked = true;   // ObsPack = the tensor as a flat bit string in output order
  struct ObsPack { u32 w; };
  // planes by CellState enum: 0 empty, 1 nought (player 1), 2 cross (player 0) — tic_tac_toe.h:51-55
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg&, int, int, ObsPack& p) {
    u32 x = xs(s), o = os(s);
    p.w = (~(x | o) & 0x1ff) | (o << 9) | (x << 18);
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) { return (float)((p.w >> e) & 1u); }
};

}  // namespace b2s
