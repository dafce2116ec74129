// y (the connection game Y) rule core on 128-bit bitboards, board_size <= 11.  Semantics: reference open_spiel/games/y/y.cc
// (CalcXY y.h:57-64: cell = x + y * board_size, valid iff x + y < board_size; Move::Edge :103-108: edges x == 0, y == 0,
// x + y == board_size - 1; DoApplyAction :280-300: a stone joins its neighbours' groups and wins when its group touches all
// three edges; LegalActions :130-141; Returns :214-218; ObservationTensor :247-258; neighbour offsets :60-64).
// The reference keeps a union-find over cells; nothing of it is observable except the outcome, so the packed state is the
// two stone sets and the winner is decided when a stone is placed, by flooding that stone's group through the mover's
// stones (adjacency = six shifts with the two wrap-around columns masked) and OR-ing the edges it reaches.
// 32 B per state as two 16-byte SoA planes.  Boards have at most 121 cells, which leaves bits 121-127 of both planes:
//   plane 0 (player 1's stones) bits 121-127: last move (127 = none) — ToString brackets it (y.cc:190-198)
//   plane 1 (player 2's stones) bit 121: player to move, bits 122-123: outcome (0 running, 1 / 2 = player 0 / 1 has won)
#pragma once
#include "common.cuh"

namespace b2s {

struct YRules {
  static constexpr int kGameId = B2S_Y;
  typedef uint4 Chunk;
  static constexpr int kChunks = 2;
  static constexpr int kMaskWords = 4;     // up to 121 actions
  static constexpr int kPlayers = 2;
  static constexpr int kMaxPath = 72;      // MCTS path stack (>= max_game_length + 2 = 66 + 2)
  static constexpr int kMaxLegal = 66;     // 11 * 12 / 2 playable cells
  static constexpr int kFilterWords = 0;
  static constexpr int kIlp = 1;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = false;
  static constexpr int kNoMove = 127;

  struct Cfg {
    int n, cells;              // board_size, n * n (the cut-off corner included: those actions are never legal)
    B128 board, not_west, not_east, edge[3];
  };
  struct S { B128 p1, p2; int mover, outcome, last; };

  static __host__ const char* make_cfg(const b2s_params& p, Cfg& c, b2s_game_info& gi) {
    c.n = p.board_size >= 0 ? p.board_size : 19;          // y.h:39
    if (c.n < 1) return "y: board_size must be positive";
    if (c.n > 11) return "y: the packed layout holds board_size <= 11";
    c.cells = c.n * c.n;
    const B128 z = {0, 0};
    c.board = c.not_west = c.not_east = c.edge[0] = c.edge[1] = c.edge[2] = z;
    for (int y = 0; y < c.n; ++y)
      for (int x = 0; x + y < c.n; ++x) {
        const B128 b = b_bit(x + y * c.n);
        c.board = b_or(c.board, b);
        if (x != 0) c.not_west = b_or(c.not_west, b);
        if (x != c.n - 1) c.not_east = b_or(c.not_east, b);
        if (x == 0) c.edge[0] = b_or(c.edge[0], b);
        if (y == 0) c.edge[1] = b_or(c.edge[1], b);
        if (x + y == c.n - 1) c.edge[2] = b_or(c.edge[2], b);
      }
    gi.num_players = 2;
    gi.num_distinct_actions = c.cells;                  // y.h:172-176
    gi.max_game_length = c.n * (c.n + 1) / 2;           // y.h:188-193
    gi.observation_tensor_size = 3 * c.cells;           // y.h:185-187
    gi.obs_shape[0] = 3; gi.obs_shape[1] = c.n; gi.obs_shape[2] = c.n;
    gi.min_utility = -1; gi.max_utility = 1;
    return nullptr;
  }

  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) {
    const ulonglong2* pl = reinterpret_cast<const ulonglong2*>(ctx.planes);
    const ulonglong2 a = pl[i], b = pl[ctx.cap + i];
    s.p1 = {a.x, a.y & ((1ull << 57) - 1ull)};
    s.p2 = {b.x, b.y & ((1ull << 57) - 1ull)};
    s.last = (int)(a.y >> 57);
    s.mover = (int)((b.y >> 57) & 1ull);
    s.outcome = (int)((b.y >> 58) & 3ull);
  }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) {
    ulonglong2* pl = reinterpret_cast<ulonglong2*>(ctx.planes);
    pl[i] = make_ulonglong2(s.p1.lo, s.p1.hi | ((u64)s.last << 57));
    pl[ctx.cap + i] = make_ulonglong2(s.p2.lo, s.p2.hi | ((u64)s.mover << 57) | ((u64)s.outcome << 58));
  }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) {
    const B128 z = {0, 0};
    s.p1 = s.p2 = z; s.mover = 0; s.outcome = 0; s.last = kNoMove;
  }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  // cells adjacent to a cell of x: (0,-1), (1,-1), (1,0), (0,1), (-1,1), (-1,0) — y.cc:60-64
  __device__ static __forceinline__ B128 neighbours(B128 x, const Cfg& c) {
    const B128 xe = b_and(x, c.not_east), xw = b_and(x, c.not_west);
    B128 r = b_or(b_shl(xe, 1), b_shr(xw, 1));
    if (c.n > 1) {                                           // shift counts 1 <= n - 1 < n <= 11
      r = b_or(r, b_or(b_shr(x, c.n), b_shl(x, c.n)));
      r = b_or(r, b_or(b_shr(xe, c.n - 1), b_shl(xw, c.n - 1)));
    }
    return b_and(r, c.board);
  }
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) { return s.outcome != 0; }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg&) { return s.outcome ? kTerminalPlayerId : s.mover; }
  __device__ static __forceinline__ void returns(const S& s, const Cfg&, float* r) {
    r[0] = s.outcome == 1 ? 1.f : s.outcome == 2 ? -1.f : 0.f;
    r[1] = s.outcome == 2 ? 1.f : s.outcome == 1 ? -1.f : 0.f;
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg& c, u32* m) {
    const B128 e = b_andn(c.board, b_or(s.p1, s.p2));
    m[0] = (u32)e.lo; m[1] = (u32)(e.lo >> 32); m[2] = (u32)e.hi; m[3] = (u32)(e.hi >> 32);
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (s.outcome) { m[0] = m[1] = m[2] = m[3] = 0; return; }
    legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg& c, const Ctx&, long long) {
    if (a < 0 || a >= c.cells) return false;
    const B128 mb = b_bit(a);
    if (!b_any(b_and(mb, c.board)) || b_any(b_and(mb, b_or(s.p1, s.p2)))) return false;
    B128 own = b_or(s.mover == 0 ? s.p1 : s.p2, mb);
    if (s.mover == 0) s.p1 = own; else s.p2 = own;
    // the new stone's group and the edges it touches (JoinGroups + the edge test, y.cc:289-297)
    B128 group = mb, frontier = mb;
    for (;;) {
      const B128 g = b_andn(b_and(neighbours(frontier, c), own), group);
      if (!b_any(g)) break;
      group = b_or(group, g);
      frontier = g;
    }
    if (b_any(b_and(group, c.edge[0])) && b_any(b_and(group, c.edge[1])) && b_any(b_and(group, c.edge[2]))) s.outcome = s.mover + 1;
    s.last = a;
    s.mover ^= 1;
    return true;
  }

  // planes (y.cc:232-258): 0 the observing player's stones, 1 the other player's, 2 empty; the cut-off corner is all zero
  static constexpr bool kObsBitPacked = true;
  struct ObsPack { u64 w[6]; };             // 3 * 121 = 363 bits
  __device__ static __forceinline__ void put_flat(ObsPack& p, int off, B128 v) {
    const int i = off >> 6, sh = off & 63;
    p.w[i] |= v.lo << sh;
    const u64 c1 = sh ? (v.lo >> (64 - sh)) : 0ull;
    if (i + 1 < 6) p.w[i + 1] |= c1 | (v.hi << sh);
    if (i + 2 < 6 && sh) p.w[i + 2] |= v.hi >> (64 - sh);
  }
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg& c, int player, int, ObsPack& p) {
    for (int k = 0; k < 6; ++k) p.w[k] = 0;
    put_flat(p, 0, player == 0 ? s.p1 : s.p2);
    put_flat(p, c.cells, player == 0 ? s.p2 : s.p1);
    put_flat(p, 2 * c.cells, b_andn(c.board, b_or(s.p1, s.p2)));
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) {
    return (float)((p.w[e >> 6] >> (e & 63)) & 1ull);
  }
};

}  // namespace b2s
