// mnk (m,n,k-game: gomoku 15x15x5 by default) rule core on 256-bit bitboards.  Semantics: reference
// open_spiel/games/mnk/mnk.cc (DoApplyAction :117-126, BoardHasLine :92-115, LegalActions :148-160, IsTerminal :207-209,
// Returns :211-219, ObservationTensor :233-246; parameters m = columns, n = rows, k = stones in a row, mnk.h:34-36).
// A state is two 256-bit stone sets, bit = row * 16 + col (up to 15 columns with at least one guard column, so the four line
// directions — strides 1, 16, 17, 15 — are plain shifts that cannot wrap; up to 15 rows, which leaves bits 240-255 free).
// The reference caches `outcome_`; so do we, in the two top bits of the x plane: 0 = running, 1 = player 0 has a line,
// 2 = player 1 has a line, 3 = board full without a line — a step then costs one line test, for the stone just placed.
// 64 B per state as four 16-byte SoA planes: x.lo, x.hi, o.lo, o.hi (each a 128-bit half of a 256-bit set).
#pragma once
#include "common.cuh"

namespace b2s {

struct MnkRules {
  static constexpr int kGameId = B2S_MNK;
  typedef uint4 Chunk;
  static constexpr int kChunks = 4;
  static constexpr int kMaskWords = 8;     // up to 225 cells
  static constexpr int kPlayers = 2;
  static constexpr int kMaxPath = 232;     // MCTS path stack (>= max_game_length + 2)
  static constexpr int kMaxLegal = 225;
  static constexpr int kFilterWords = 0;
  static constexpr int kIlp = 1;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = false;
  static constexpr int kStride = 16;
  static constexpr int kObsWords = 11;     // 3 * 225 bits

  struct Cfg { int rows, cols, k, cells; };
  struct S { B256 x, o; int outcome; };     // outcome: 0 running, 1 / 2 = player 0 / 1 won, 3 = full board

  static __host__ const char* make_cfg(const b2s_params& p, Cfg& c, b2s_game_info& gi) {
    c.cols = p.columns >= 0 ? p.columns : 15;      // "m", mnk.h:35
    c.rows = p.rows >= 0 ? p.rows : 15;            // "n", mnk.h:34
    c.k = p.x_in_row >= 0 ? p.x_in_row : 5;        // "k", mnk.h:36
    if (c.rows < 1 || c.cols < 1 || c.k < 1) return "mnk: m, n, k must be positive";
    if (c.cols > 15 || c.rows > 15) return "mnk: the packed layout holds boards up to 15 x 15";
    c.cells = c.rows * c.cols;
    gi.num_players = 2;
    gi.num_distinct_actions = c.cells;             // mnk.h:106
    gi.max_game_length = c.cells;                  // mnk.h:118
    gi.observation_tensor_size = 3 * c.cells;      // mnk.h:115-117
    gi.obs_shape[0] = 3; gi.obs_shape[1] = c.rows; gi.obs_shape[2] = c.cols;
    gi.min_utility = -1; gi.max_utility = 1;
    return nullptr;
  }

  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) {
    const ulonglong2* pl = reinterpret_cast<const ulonglong2*>(ctx.planes);
    const ulonglong2 a = pl[i], b = pl[ctx.cap + i], c = pl[2 * ctx.cap + i], d = pl[3 * ctx.cap + i];
    s.x = {{a.x, a.y, b.x, b.y & 0x3fffffffffffffffull}};
    s.o = {{c.x, c.y, d.x, d.y}};
    s.outcome = (int)(b.y >> 62);
  }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) {
    ulonglong2* pl = reinterpret_cast<ulonglong2*>(ctx.planes);
    pl[i] = make_ulonglong2(s.x.w[0], s.x.w[1]);
    pl[ctx.cap + i] = make_ulonglong2(s.x.w[2], s.x.w[3] | ((u64)s.outcome << 62));
    pl[2 * ctx.cap + i] = make_ulonglong2(s.o.w[0], s.o.w[1]);
    pl[3 * ctx.cap + i] = make_ulonglong2(s.o.w[2], s.o.w[3]);
  }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) {
    s.x = {{0, 0, 0, 0}}; s.o = {{0, 0, 0, 0}}; s.outcome = 0;
  }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ int stones(const B256& b) { return __popcll(b.w[0]) + __popcll(b.w[1]) + __popcll(b.w[2]) + __popcll(b.w[3]); }
  __device__ static __forceinline__ int mover(const S& s) { return stones(s.x) > stones(s.o) ? 1 : 0; }      // x (player 0) moves first
  // k stones in a row in any of the four line directions (BoardHasLine :92-115 looks along all eight; lines are symmetric)
  __device__ static __forceinline__ bool has_line(const B256& b, const Cfg& c) {
    const int d[4] = {1, kStride, kStride + 1, kStride - 1};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if ((c.k - 1) * d[j] >= 256) continue;
      B256 m = b;
      for (int i = 1; i < c.k; ++i) m = q_and(m, q_shr(b, i * d[j]));
      if (q_any(m)) return true;
    }
    return false;
  }
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) { return s.outcome != 0; }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg& c) { return terminal(s, c) ? kTerminalPlayerId : mover(s); }
  __device__ static __forceinline__ void returns(const S& s, const Cfg&, float* r) {
    r[0] = s.outcome == 1 ? 1.f : s.outcome == 2 ? -1.f : 0.f;
    r[1] = s.outcome == 2 ? 1.f : s.outcome == 1 ? -1.f : 0.f;
  }
  // 256-bit board (stride 16) -> action-ordered bits (row * cols + col) appended at bit offset `off` of a word array
  __device__ static __forceinline__ void deposit_rows(const B256& b, const Cfg& c, u64* words, int off) {
    const u64 rowmask = (1ull << c.cols) - 1ull;
    for (int r = 0; r < c.rows; ++r) {
      const u64 row = (b.w[r >> 2] >> ((r & 3) * kStride)) & rowmask;
      const int pos = off + r * c.cols;
      words[pos >> 6] |= row << (pos & 63);
      if ((pos & 63) + c.cols > 64) words[(pos >> 6) + 1] |= row >> (64 - (pos & 63));
    }
  }
  __device__ static __forceinline__ B256 empties(const S& s, const Cfg& c) {
    const u64 rowmask = (1ull << c.cols) - 1ull;
    B256 e;
    for (int w = 0; w < 4; ++w) {
      u64 board = 0;
      for (int q = 0; q < 4; ++q) if (4 * w + q < c.rows) board |= rowmask << (q * kStride);
      e.w[w] = board & ~(s.x.w[w] | s.o.w[w]);
    }
    return e;
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg& c, u32* m) {
    u64 w[4] = {0, 0, 0, 0};
    deposit_rows(empties(s, c), c, w, 0);
    for (int i = 0; i < 4; ++i) { m[2 * i] = (u32)w[i]; m[2 * i + 1] = (u32)(w[i] >> 32); }
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (terminal(s, c)) { for (int i = 0; i < kMaskWords; ++i) m[i] = 0; return; }
    legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg& c, const Ctx&, long long) {
    if (a < 0 || a >= c.cells) return false;
    const int r = a / c.cols, col = a - r * c.cols, bit = r * kStride + col;
    if (q_test(s.x, bit) || q_test(s.o, bit)) return false;
    const int mv = mover(s);
    B256& mine = mv == 0 ? s.x : s.o;
    q_set(mine, bit);
    // outcome_ after the move (mnk.cc:120-122, 207-209): a line for the mover, else a full board
    s.outcome = has_line(mine, c) ? mv + 1 : (stones(s.x) + stones(s.o) == c.cells ? 3 : 0);
    return true;
  }

  // planes by CellState (mnk.h:38-42): 0 empty, 1 nought (player 1), 2 cross (player 0); [plane][row][col]
  static constexpr bool kObsBitPacked = true;
  struct ObsPack { u64 w[kObsWords + 1]; };
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg& c, int, int, ObsPack& p) {
    for (int i = 0; i <= kObsWords; ++i) p.w[i] = 0;
    deposit_rows(empties(s, c), c, p.w, 0);
    deposit_rows(s.o, c, p.w, c.cells);
    deposit_rows(s.x, c, p.w, 2 * c.cells);
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) {
    return (float)((p.w[e >> 6] >> (e & 63)) & 1ull);
  }
};

}  // namespace b2s
