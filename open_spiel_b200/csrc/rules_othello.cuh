// othello (8x8) rule core on two 64-bit bitboards.  Semantics: reference open_spiel/games/othello/othello.cc
// (DoApplyAction :177-207, CountSteps / CanCapture / Capture :124-163, LegalActions :219-224 — the pass move 64 is legal only
// when the mover has no capture —, NoValidActions :169-172, Returns :276-284, ObservationTensor :298-316, initial position
// :237-243).  bit = row * 8 + col; black = player 0 moves first.
// Move generation and flips are the classic directional fills: along each of the eight rays a set of the mover's discs is
// smeared through runs of opposing discs (six steps cover the longest run on an 8x8 board), with the wrap-around columns
// masked off for the six rays that change the column.
// 16 B per state, one 128-bit chunk {black, white}.  The player to move is not derivable from the discs (passes break the
// parity): white is stored complemented when player 1 is to move — the two sets are disjoint, so black & stored_white == 0
// exactly when player 0 is to move; a position where black has no discs is over whoever "moves", and is stored plain.
// The end of the game (neither side can move, othello.cc:169-172) IS derivable and is re-derived when a state is loaded;
// the unpacked state carries the mover's move set so that the kernels that keep a state in registers across plies (playouts,
// MCTS) generate moves once per ply.
#pragma once
#include "common.cuh"

namespace b2s {

struct OthelloRules {
  static constexpr int kGameId = B2S_OTHELLO;
  typedef uint4 Chunk;
  static constexpr int kChunks = 1;
  static constexpr int kMaskWords = 3;     // 65 actions: 64 cells + pass
  static constexpr int kPlayers = 2;
  static constexpr int kMaxPath = 132;     // MCTS path stack (>= max_game_length + 2)
  static constexpr int kMaxLegal = 33;     // most legal actions any state can have (no known position exceeds 33)
  static constexpr int kFilterWords = 0;
  static constexpr int kIlp = 2;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = false;
  static constexpr int kPass = 64;
  static constexpr u64 kNotColA = 0xfefefefefefefefeull, kNotColH = 0x7f7f7f7f7f7f7f7full;

  struct Cfg { int unused; };
  // moves = the mover's capturing moves; term = neither side has one
  struct S { u64 b, w, moves; int mover, term; };

  static __host__ const char* make_cfg(const b2s_params&, Cfg& c, b2s_game_info& gi) {
    c.unused = 0;
    gi.num_players = 2;
    gi.num_distinct_actions = 65;                  // othello.h:145
    gi.max_game_length = 128;                      // othello.h:158
    gi.observation_tensor_size = 192;              // othello.h:153-155
    gi.obs_shape[0] = 3; gi.obs_shape[1] = 8; gi.obs_shape[2] = 8;
    gi.min_utility = -1; gi.max_utility = 1;
    return nullptr;
  }

  // one step along ray d (0..7: E, W, S, N, SE, SW, NE, NW in row-major bit order), wrap-around columns removed
  template <int D>
  __device__ static __forceinline__ u64 step(u64 x) {
    switch (D) {
      case 0: return (x << 1) & kNotColA;
      case 1: return (x >> 1) & kNotColH;
      case 2: return x << 8;
      case 3: return x >> 8;
      case 4: return (x << 9) & kNotColA;
      case 5: return (x << 7) & kNotColH;
      case 6: return (x >> 7) & kNotColA;
      default: return (x >> 9) & kNotColH;
    }
  }
  // opposing discs reachable from `from` along ray D through an unbroken run of opposing discs
  template <int D>
  __device__ static __forceinline__ u64 run(u64 from, u64 opp) {
    u64 x = step<D>(from) & opp;
#pragma unroll
    for (int i = 0; i < 5; ++i) x |= step<D>(x) & opp;
    return x;
  }
  template <int D>
  __device__ static __forceinline__ u64 moves_dir(u64 mine, u64 opp, u64 empty) { return step<D>(run<D>(mine, opp)) & empty; }
  // cells where `mine` captures at least one opposing disc (CanCapture, othello.cc:138-148)
  __device__ static __forceinline__ u64 gen_moves(u64 mine, u64 opp) {
    const u64 empty = ~(mine | opp);
    return moves_dir<0>(mine, opp, empty) | moves_dir<1>(mine, opp, empty) | moves_dir<2>(mine, opp, empty) |
           moves_dir<3>(mine, opp, empty) | moves_dir<4>(mine, opp, empty) | moves_dir<5>(mine, opp, empty) |
           moves_dir<6>(mine, opp, empty) | moves_dir<7>(mine, opp, empty);
  }
  template <int D>
  __device__ static __forceinline__ u64 flips_dir(u64 sq, u64 mine, u64 opp) {
    const u64 x = run<D>(sq, opp);
    return (step<D>(x) & mine) ? x : 0ull;        // the run must end on one of the mover's discs (CountSteps :124-136)
  }
  __device__ static __forceinline__ u64 gen_flips(u64 sq, u64 mine, u64 opp) {
    return flips_dir<0>(sq, mine, opp) | flips_dir<1>(sq, mine, opp) | flips_dir<2>(sq, mine, opp) | flips_dir<3>(sq, mine, opp) |
           flips_dir<4>(sq, mine, opp) | flips_dir<5>(sq, mine, opp) | flips_dir<6>(sq, mine, opp) | flips_dir<7>(sq, mine, opp);
  }
  // the mover's move set and the end-of-game flag of a position
  __device__ static __forceinline__ void derive(S& s) {
    const u64 mine = s.mover == 0 ? s.b : s.w, opp = s.mover == 0 ? s.w : s.b;
    s.moves = gen_moves(mine, opp);
    s.term = (s.moves == 0 && gen_moves(opp, mine) == 0) ? 1 : 0;
  }

  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) {
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(ctx.planes)[i];
    s.b = v.x;
    s.mover = (v.x & v.y) != 0 ? 1 : 0;
    s.w = s.mover ? ~v.y : v.y;
    derive(s);
  }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) {
    const u64 w = (s.mover == 1 && s.b != 0) ? ~s.w : s.w;
    reinterpret_cast<ulonglong2*>(ctx.planes)[i] = make_ulonglong2(s.b, w);
  }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) {
    s.w = 1ull << 27 | 1ull << 36;                 // (3,3), (4,4) white; (3,4), (4,3) black
    s.b = 1ull << 28 | 1ull << 35;
    s.mover = 0;
    derive(s);
  }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) { return s.term != 0; }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg&) { return s.term ? kTerminalPlayerId : s.mover; }
  __device__ static __forceinline__ void returns(const S& s, const Cfg&, float* r) {
    r[0] = 0.f; r[1] = 0.f;
    if (!s.term) return;
    const int d = __popcll(s.b) - __popcll(s.w);   // outcome_ by disc count (othello.cc:192-201)
    r[0] = d > 0 ? 1.f : d < 0 ? -1.f : 0.f;
    r[1] = d < 0 ? 1.f : d > 0 ? -1.f : 0.f;
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg&, u32* m) {
    m[0] = (u32)s.moves; m[1] = (u32)(s.moves >> 32);
    m[2] = s.moves == 0 ? 1u : 0u;                 // pass only when there is nothing else (othello.cc:219-224)
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (s.term) { m[0] = 0; m[1] = 0; m[2] = 0; return; }
    legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg&, const Ctx&, long long) {
    if (a == kPass) {
      if (s.moves != 0) return false;
      s.mover ^= 1;                                // othello.cc:178-181; the other side can move, or the game would be over
      const u64 mine = s.mover == 0 ? s.b : s.w, opp = s.mover == 0 ? s.w : s.b;
      s.moves = gen_moves(mine, opp);
      return true;
    }
    if ((unsigned)a > 63u || !((s.moves >> a) & 1ull)) return false;
    const u64 sq = 1ull << a;
    u64 mine = s.mover == 0 ? s.b : s.w, opp = s.mover == 0 ? s.w : s.b;
    const u64 f = gen_flips(sq, mine, opp);
    mine |= sq | f;
    opp &= ~f;
    if (s.mover == 0) { s.b = mine; s.w = opp; } else { s.w = mine; s.b = opp; }
    s.mover ^= 1;                                  // othello.cc:203-205 (stays meaningless once the game is over)
    s.moves = gen_moves(opp, mine);
    s.term = (s.moves == 0 && gen_moves(mine, opp) == 0) ? 1 : 0;
    return true;
  }

  // planes (othello.cc:298-316): 0 empty, 1 the observing player's discs, 2 the opponent's; [plane][cell]
  static constexpr bool kObsBitPacked = true;
  struct ObsPack { u64 w[3]; };
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg&, int player, int, ObsPack& p) {
    p.w[0] = ~(s.b | s.w);
    p.w[1] = player == 0 ? s.b : s.w;
    p.w[2] = player == 0 ? s.w : s.b;
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) {
    return (float)((p.w[e >> 6] >> (e & 63)) & 1ull);
  }
};

}  // namespace b2s
