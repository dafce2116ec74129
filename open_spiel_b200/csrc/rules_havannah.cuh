// havannah rule core on 256-bit bitboards, board_size <= 8 (diameter 15).  Semantics: reference
// open_spiel/games/havannah/havannah.cc (CalcXY havannah.h:58-66: cell = x + y * diameter, valid iff |x - y| < board_size;
// Move::Corner / Edge :128-158; neighbour offsets :74-78; DoApplyAction :324-359 incl. the swap move; JoinGroups :375-392;
// CheckRingDFS :394-409; LegalActions :186-201; Returns :281-286; ObservationTensor :310-322).
// The reference keeps a union-find over cells.  Observable through the outcome are (i) the connected group of the new stone
// — re-derived here by flooding through the mover's stones, its corners and edges OR-ed together (a win with >= 2 corners
// or >= 3 edges) —, and (ii) the ring test, which the reference runs only when `alreadyjoined`, i.e. when its neighbour loop
// (same-coloured neighbour found -> join, then SKIP the next direction) met a neighbour whose group had already been
// merged into the new stone's.  Both the loop with its skip rule and the bounded depth-first search for a ring (marks on
// the current path only, at most one 60-degree turn per step) are reproduced literally, because a differently defined ring
// test would not give the reference's answers on every position.
// 64 B per state as four 16-byte SoA planes: p1.lo, p1.hi, p2.lo, p2.hi.  225 cells leave bits 225-255 of each set:
//   p1 word 3, bits 56-63: last move (255 = none; ToString brackets it, and the swap move is "play on it")
//   p2 word 3, bit 63: player to move, bits 61-62: outcome (0 running, 1 / 2 = player 0 / 1 has won, 3 = draw)
#pragma once
#include "common.cuh"

namespace b2s {

struct HavannahRules {
  static constexpr int kGameId = B2S_HAVANNAH;
  typedef uint4 Chunk;
  static constexpr int kChunks = 4;
  static constexpr int kMaskWords = 8;     // up to 225 actions
  static constexpr int kPlayers = 2;
  static constexpr int kMaxPath = 176;     // MCTS path stack (>= max_game_length + 2 = 170 + 2)
  static constexpr int kMaxLegal = 170;    // 169 playable cells (+ the swap move never coincides with an empty cell)
  static constexpr int kFilterWords = 0;
  static constexpr int kIlp = 1;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = false;
  static constexpr int kNoMove = 255;
  static constexpr int kMaxStack = 128;    // ring search depth: a path of distinct stones of one colour (at most 85 of the 169 cells)

  struct Cfg {
    int size, d, cells, valid, swap;       // board_size, diameter, d * d, playable cells, swap rule
    B256 board, not_first_col, not_last_col, corner[6], edge[6];
  };
  struct S { B256 p1, p2; int mover, outcome, last; };

  static __host__ const char* make_cfg(const b2s_params& p, Cfg& c, b2s_game_info& gi) {
    c.size = p.board_size >= 0 ? p.board_size : 8;       // havannah.h:38
    c.swap = p.swap > 0 ? 1 : 0;
    if (c.size < 1) return "havannah: board_size must be positive";
    if (c.size > 8) return "havannah: the packed layout holds board_size <= 8";
    c.d = 2 * c.size - 1;
    c.cells = c.d * c.d;
    c.valid = c.cells - c.size * (c.size - 1);
    const B256 z = {{0, 0, 0, 0}};
    c.board = c.not_first_col = c.not_last_col = z;
    for (int k = 0; k < 6; ++k) c.corner[k] = c.edge[k] = z;
    const int m = c.size - 1, e = 2 * m;
    for (int y = 0; y < c.d; ++y)
      for (int x = 0; x < c.d; ++x) {
        if (!(y - x < c.size && x - y < c.size)) continue;
        const int i = x + y * c.d;
        q_set(c.board, i);
        if (x != 0) q_set(c.not_first_col, i);
        if (x != c.d - 1) q_set(c.not_last_col, i);
        // Move::Corner (havannah.cc:128-142): first match wins
        int corner = -1;
        if (x == 0 && y == 0) corner = 0;
        else if (x == m && y == 0) corner = 1;
        else if (x == e && y == m) corner = 2;
        else if (x == e && y == e) corner = 3;
        else if (x == m && y == e) corner = 4;
        else if (x == 0 && y == m) corner = 5;
        if (corner >= 0) q_set(c.corner[corner], i);
        // Move::Edge (havannah.cc:144-158): first match wins
        int edge = -1;
        if (y == 0 && x != 0 && x != m) edge = 0;
        else if (x - y == m && x != m && x != e) edge = 1;
        else if (x == e && y != m && y != e) edge = 2;
        else if (y == e && x != e && x != m) edge = 3;
        else if (y - x == m && x != m && x != 0) edge = 4;
        else if (x == 0 && y != m && y != 0) edge = 5;
        if (edge >= 0) q_set(c.edge[edge], i);
      }
    gi.num_players = 2;
    gi.num_distinct_actions = c.cells;                  // havannah.h:205-209
    gi.max_game_length = c.valid + c.swap;              // havannah.h:221-226
    gi.observation_tensor_size = 3 * c.cells;           // havannah.h:218-220
    gi.obs_shape[0] = 3; gi.obs_shape[1] = c.d; gi.obs_shape[2] = c.d;
    gi.min_utility = -1; gi.max_utility = 1;
    return nullptr;
  }

  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) {
    const ulonglong2* pl = reinterpret_cast<const ulonglong2*>(ctx.planes);
    const ulonglong2 a = pl[i], b = pl[ctx.cap + i], c = pl[2 * ctx.cap + i], d = pl[3 * ctx.cap + i];
    const u64 keep = (1ull << 56) - 1ull;
    s.p1 = {{a.x, a.y, b.x, b.y & keep}};
    s.p2 = {{c.x, c.y, d.x, d.y & keep}};
    s.last = (int)(b.y >> 56);
    s.mover = (int)(d.y >> 63);
    s.outcome = (int)((d.y >> 61) & 3ull);
  }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) {
    ulonglong2* pl = reinterpret_cast<ulonglong2*>(ctx.planes);
    pl[i] = make_ulonglong2(s.p1.w[0], s.p1.w[1]);
    pl[ctx.cap + i] = make_ulonglong2(s.p1.w[2], s.p1.w[3] | ((u64)s.last << 56));
    pl[2 * ctx.cap + i] = make_ulonglong2(s.p2.w[0], s.p2.w[1]);
    pl[3 * ctx.cap + i] = make_ulonglong2(s.p2.w[2], s.p2.w[3] | ((u64)s.mover << 63) | ((u64)s.outcome << 61));
  }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) {
    const B256 z = {{0, 0, 0, 0}};
    s.p1 = s.p2 = z; s.mover = 0; s.outcome = 0; s.last = kNoMove;
  }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  // the six neighbours (-1,-1), (0,-1), (1,0), (1,1), (0,1), (-1,0) (havannah.cc:74-78) of every cell of x
  __device__ static __forceinline__ B256 neighbours(const B256& x, const Cfg& c) {
    const B256 xl = q_and(x, c.not_last_col), xf = q_and(x, c.not_first_col);
    B256 r = q_or(q_shl(xl, 1), q_shr(xf, 1));
    if (c.d > 1) {                                           // shift counts 1 .. 16
      r = q_or(r, q_or(q_shl(x, c.d), q_shr(x, c.d)));
      r = q_or(r, q_or(q_shl(xl, c.d + 1), q_shr(xf, c.d + 1)));
    }
    return q_and(r, c.board);
  }
  // neighbour `dir` of one cell, -1 when it is off the board
  __device__ static __forceinline__ int neighbour(int cell, int dir, const Cfg& c) {
    const int y = cell / c.d, x = cell - y * c.d;
    const int dx = dir == 0 || dir == 5 ? -1 : (dir == 2 || dir == 3 ? 1 : 0);
    const int dy = dir == 0 || dir == 1 ? -1 : (dir == 3 || dir == 4 ? 1 : 0);
    const int nx = x + dx, ny = y + dy;
    if (nx < 0 || ny < 0 || nx >= c.d || ny >= c.d || !(ny - nx < c.size && nx - ny < c.size)) return -1;
    return nx + ny * c.d;
  }
  __device__ static __forceinline__ B256 flood(int from, const B256& within, const Cfg& c) {
    B256 group = {{0, 0, 0, 0}};
    q_set(group, from);
    B256 frontier = group;
    for (;;) {
      const B256 g = q_andn(q_and(neighbours(frontier, c), within), group);
      if (!q_any(g)) break;
      group = q_or(group, g);
      frontier = g;
    }
    return group;
  }
  // CheckRingDFS(move, 0, 3) (havannah.cc:394-409) with an explicit stack: a path of the mover's stones from `move`, at most
  // one 60-degree turn per step, that runs into a cell already on the path
  __device__ static __forceinline__ bool ring_from(int move, const B256& own, const Cfg& c) {
    unsigned char cell[kMaxStack];
    signed char next[kMaxStack], right[kMaxStack];
    B256 mark = {{0, 0, 0, 0}};
    int sp = 0;
    cell[0] = (unsigned char)move; next[0] = 0; right[0] = 3;
    q_set(mark, move);
    for (;;) {
      if (next[sp] > right[sp]) {                            // this cell's directions are exhausted: unmark, back to the parent
        q_clear(mark, cell[sp]);
        if (sp == 0) return false;
        --sp;
        continue;
      }
      const int i = next[sp]++;
      const int dir = (i + 6) % 6;
      const int nb = neighbour(cell[sp], dir, c);
      if (nb < 0 || !q_test(own, nb)) continue;
      if (q_test(mark, nb)) return true;                     // found a ring
      if (sp + 1 >= kMaxStack) return false;                 // cannot happen: a path holds distinct stones of one colour
      ++sp;
      cell[sp] = (unsigned char)nb; next[sp] = (signed char)(dir - 1); right[sp] = (signed char)(dir + 1);
      q_set(mark, nb);
    }
  }

  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) { return s.outcome != 0; }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg&) { return s.outcome ? kTerminalPlayerId : s.mover; }
  __device__ static __forceinline__ void returns(const S& s, const Cfg&, float* r) {
    r[0] = s.outcome == 1 ? 1.f : s.outcome == 2 ? -1.f : 0.f;
    r[1] = s.outcome == 2 ? 1.f : s.outcome == 1 ? -1.f : 0.f;
  }
  __device__ static __forceinline__ bool allow_swap(const S& s, const Cfg& c) {                // havannah.cc:208-210
    return c.swap && s.mover == 1 && q_popc(s.p1) + q_popc(s.p2) == 1;
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg& c, u32* m) {
    B256 e = q_andn(c.board, q_or(s.p1, s.p2));
    if (allow_swap(s, c)) q_set(e, s.last);                  // the second move may replace the first (havannah.cc:196-199)
    for (int i = 0; i < 4; ++i) { m[2 * i] = (u32)e.w[i]; m[2 * i + 1] = (u32)(e.w[i] >> 32); }
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (s.outcome) { for (int i = 0; i < kMaskWords; ++i) m[i] = 0; return; }
    legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg& c, const Ctx&, long long) {
    if (a < 0 || a >= c.cells || !q_test(c.board, a)) return false;
    B256& mine = s.mover == 0 ? s.p1 : s.p2;
    if (a == s.last && allow_swap(s, c)) {
      q_clear(s.p1, a);                                      // the stone changes colour; moves_made_ and last_move_ stay
    } else {
      if (q_test(s.p1, a) || q_test(s.p2, a)) return false;
      s.last = a;
    }
    const B256 before = mine;                                // the mover's stones without the new one
    q_set(mine, a);
    // the neighbour loop of DoApplyAction (:339-350): join, then skip the next direction; alreadyjoined = a neighbour whose
    // group had been merged into the new stone's by an earlier join
    bool alreadyjoined = false, skip = false;
    B256 joined = {{0, 0, 0, 0}};
    for (int dir = 0; dir < 6; ++dir) {
      if (skip) { skip = false; continue; }
      const int nb = neighbour(a, dir, c);
      if (nb < 0 || !q_test(before, nb)) continue;
      if (q_test(joined, nb)) alreadyjoined = true;
      else joined = q_or(joined, flood(nb, before, c));
      skip = true;
    }
    q_set(joined, a);                                        // = the new stone's connected group
    int corners = 0, edges = 0;
    for (int k = 0; k < 6; ++k) {
      corners += q_any(q_and(joined, c.corner[k])) ? 1 : 0;
      edges += q_any(q_and(joined, c.edge[k])) ? 1 : 0;
    }
    if (edges >= 3 || corners >= 2 || (alreadyjoined && ring_from(a, mine, c))) s.outcome = s.mover + 1;
    else if (q_popc(s.p1) + q_popc(s.p2) == c.valid) s.outcome = 3;                           // board full: draw (:355-357)
    s.mover ^= 1;
    return true;
  }

  // planes (havannah.cc:296-322): 0 the observing player's stones, 1 the other player's, 2 empty; cut-off corners all zero
  static constexpr bool kObsBitPacked = true;
  static constexpr int kObsWords = 11;      // 3 * 225 = 675 bits
  struct ObsPack { u64 w[kObsWords + 1]; };
  __device__ static __forceinline__ void put_flat(ObsPack& p, int off, const B256& v, int bits) {
    for (int k = 0; k < 4; ++k) {
      const int pos = off + 64 * k;
      if (64 * k >= bits) break;
      const int i = pos >> 6, sh = pos & 63;
      p.w[i] |= v.w[k] << sh;
      if (sh && i + 1 <= kObsWords) p.w[i + 1] |= v.w[k] >> (64 - sh);
    }
  }
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg& c, int player, int, ObsPack& p) {
    for (int k = 0; k <= kObsWords; ++k) p.w[k] = 0;
    put_flat(p, 0, player == 0 ? s.p1 : s.p2, c.cells);
    put_flat(p, c.cells, player == 0 ? s.p2 : s.p1, c.cells);
    put_flat(p, 2 * c.cells, q_andn(c.board, q_or(s.p1, s.p2)), c.cells);
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) {
    return (float)((p.w[e >> 6] >> (e & 63)) & 1ull);
  }
};

}  // namespace b2s
