// leduc_poker (2 players) rule core.  Semantics: reference open_spiel/games/leduc_poker/leduc_poker.cc
// (DoApplyAction :298-414, LegalActions :416-457, IsTerminal :498-500, Returns :502-514, NextPlayer :573-591,
// RankHand :593-626, ResolveWinner :628-678, ReadyForNextRound :680-683, observer tensors :92-192).
// Packed into one uint64 (field: bits):
//   priv0 0-2, priv1 3-5, pub 6-8 (7 = not dealt) | r1len 9-11, r1seq 12-19 | r2len 20-22, r2seq 23-30 |
//   cur 31-32 (0,1 player; 2 chance) | round2 33 | calls 34-35 | raises 36-37 | stakes 38-41 |
//   ante0 42-45, ante1 46-49 | folded0 50, folded1 51 | dealt 52-53
#pragma once
#include "common.cuh"

namespace b2s {

struct LeducRules {
  static constexpr int kGameId = B2S_LEDUC_POKER;
  typedef u64 Chunk;
  static constexpr int kChunks = 1;
  static constexpr int kMaskWords = 1;
  static constexpr int kPlayers = 2;
  static constexpr int kMaxPath = 0;   // MCTS path stack (>= max_game_length + 2); 0 = no device MCTS
  static constexpr int kMaxLegal = 6;   // most legal actions any state can have (MCTS children block size)
  static constexpr int kFilterWords = 0;   // no per-lane history filter (see rules_go.cuh)
  static constexpr int kIlp = 2;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = true;
  struct Cfg { int starting_player; };
  struct S {
    int priv0, priv1, pub, r1len, r1seq, r2len, r2seq, cur, round2, calls, raises, stakes, ante0, ante1, folded0, folded1, dealt;
  };
  // two-player fields are selected, never indexed, so the state stays in registers
  __device__ static __forceinline__ int priv_of(const S& s, int p) { return p ? s.priv1 : s.priv0; }
  __device__ static __forceinline__ int ante_of(const S& s, int p) { return p ? s.ante1 : s.ante0; }
  __device__ static __forceinline__ int folded_of(const S& s, int p) { return p ? s.folded1 : s.folded0; }
  __device__ static __forceinline__ void set_ante(S& s, int p, int v) { if (p) s.ante1 = v; else s.ante0 = v; }
  __device__ static __forceinline__ void set_folded(S& s, int p) { if (p) s.folded1 = 1; else s.folded0 = 1; }
  static constexpr int kNone = 7, kChance = 2;

  static __host__ const char* make_cfg(const b2s_params& p, Cfg& c, b2s_game_info& gi) {
    int n = p.players >= 0 ? p.players : 2;
    if (n != 2) return "leduc_poker: the device path supports players=2 only";
    c.starting_player = p.starting_player >= 0 ? p.starting_player : 0;
    if (c.starting_player > 1) return "leduc_poker: starting_player out of range";
    gi.num_players = 2;
    gi.num_distinct_actions = 3;
    gi.max_chance_outcomes = 6;
    gi.max_game_length = 8;                          // 2*(3n-2), leduc_poker.h:233-241
    gi.information_state_tensor_size = 30;           // leduc_poker.cc:811-820
    gi.observation_tensor_size = 16;                 // leduc_poker.cc:822-831
    gi.obs_shape[0] = 16;
    gi.min_utility = -13; gi.max_utility = 13;
    return nullptr;
  }
  __device__ static __forceinline__ u64 pack(const S& s) {
    return (u64)s.priv0 | (u64)s.priv1 << 3 | (u64)s.pub << 6 | (u64)s.r1len << 9 | (u64)s.r1seq << 12 |
           (u64)s.r2len << 20 | (u64)s.r2seq << 23 | (u64)s.cur << 31 | (u64)s.round2 << 33 | (u64)s.calls << 34 |
           (u64)s.raises << 36 | (u64)s.stakes << 38 | (u64)s.ante0 << 42 | (u64)s.ante1 << 46 |
           (u64)s.folded0 << 50 | (u64)s.folded1 << 51 | (u64)s.dealt << 52;
  }
  __device__ static __forceinline__ void unpack(S& s, u64 v) {
    s.priv0 = v & 7; s.priv1 = (v >> 3) & 7; s.pub = (v >> 6) & 7; s.r1len = (v >> 9) & 7; s.r1seq = (v >> 12) & 255;
    s.r2len = (v >> 20) & 7; s.r2seq = (v >> 23) & 255; s.cur = (v >> 31) & 3; s.round2 = (v >> 33) & 1;
    s.calls = (v >> 34) & 3; s.raises = (v >> 36) & 3; s.stakes = (v >> 38) & 15; s.ante0 = (v >> 42) & 15;
    s.ante1 = (v >> 46) & 15; s.folded0 = (v >> 50) & 1; s.folded1 = (v >> 51) & 1; s.dealt = (v >> 52) & 3;
  }
  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) { unpack(s, reinterpret_cast<const u64*>(ctx.planes)[i]); }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) { reinterpret_cast<u64*>(ctx.planes)[i] = pack(s); }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) {
    s.priv0 = s.priv1 = s.pub = kNone;
    s.r1len = s.r1seq = s.r2len = s.r2seq = 0;
    s.cur = kChance; s.round2 = 0; s.calls = 0; s.raises = 0; s.stakes = 1;
    s.ante0 = s.ante1 = 1; s.folded0 = s.folded1 = 0; s.dealt = 0;
  }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ int remaining(const S& s) { return 2 - s.folded0 - s.folded1; }
  __device__ static __forceinline__ bool ready_next(const S& s) {
    return (s.raises == 0 && s.calls == remaining(s)) || (s.raises > 0 && s.calls == remaining(s) - 1);
  }
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) {
    return remaining(s) == 1 || (s.round2 && s.cur != kChance && ready_next(s) && s.pub != kNone);
  }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg& c) {
    if (terminal(s, c)) return kTerminalPlayerId;
    return s.cur == kChance ? kChancePlayerId : s.cur;
  }
  __device__ static __forceinline__ int rank(const S& s, int p) {
    int lo = s.pub, hi = priv_of(s, p);
    if (lo > hi) { int t = lo; lo = hi; hi = t; }
    if ((lo & 1) == 0 && hi == lo + 1) return 36 + lo;
    return (hi >> 1) * 6 + (lo >> 1);
  }
  __device__ static __forceinline__ void returns(const S& s, const Cfg& c, float* r) {
    r[0] = 0.f; r[1] = 0.f;
    if (!terminal(s, c)) return;
    int pot = s.ante0 + s.ante1;
    int w;                                  // winner, or -1 for a split pot
    if (remaining(s) == 1) w = s.folded0 ? 1 : 0;
    else { int r0 = rank(s, 0), r1 = rank(s, 1); w = r0 == r1 ? -1 : (r0 > r1 ? 0 : 1); }
    if (w < 0) {                            // split pot: money += pot / 2.0 (leduc_poker.cc:670-676)
      r[0] = (float)pot * 0.5f - (float)s.ante0;
      r[1] = (float)pot * 0.5f - (float)s.ante1;
    } else {
      float win = (float)(pot - ante_of(s, w)), lose = (float)(-ante_of(s, 1 - w));
      r[0] = w == 0 ? win : lose;
      r[1] = w == 1 ? win : lose;
    }
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg&, u32* m) {
    if (s.cur == kChance) {
      u32 deck = 63u;
      if (s.priv0 != kNone) deck &= ~(1u << s.priv0);
      if (s.priv1 != kNone) deck &= ~(1u << s.priv1);
      if (s.pub != kNone) deck &= ~(1u << s.pub);
      m[0] = deck;
      return;
    }
    u32 v = 2u;                                             // call always
    if (s.stakes > ante_of(s, s.cur)) v |= 1u;                  // fold only under pressure
    if (s.raises < 2) v |= 4u;
    m[0] = v;
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (terminal(s, c)) m[0] = 0; else legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ int next_player(const S& s, const Cfg& c) {
    int from = s.cur == kChance ? ((c.starting_player + 1) & 1) : s.cur;
    int p = (from + 1) & 1;
    if (!folded_of(s, p)) return p;
    return from;
  }
  __device__ static __forceinline__ void append(S& s, int mv) {
    if (!s.round2) { s.r1seq |= mv << (2 * s.r1len); s.r1len++; }
    else { s.r2seq |= mv << (2 * s.r2len); s.r2len++; }
  }
  __device__ static __forceinline__ void after_move(S& s, const Cfg& c, bool may_advance) {
    if (terminal(s, c)) return;
    if (may_advance && ready_next(s)) { s.round2 = 1; s.raises = 0; s.calls = 0; s.cur = kChance; }
    else s.cur = next_player(s, c);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg& c, const Ctx&, long long) {
    if (s.cur == kChance) {
      u32 m; legal_nonterminal(s, c, &m);
      if (a < 0 || a > 5 || !((m >> a) & 1u)) return false;
      if (s.dealt < 2) {
        if (s.dealt) s.priv1 = a; else s.priv0 = a;
        s.dealt++;
        if (s.dealt == 2) s.cur = c.starting_player;
      } else {
        s.pub = a;
        s.cur = next_player(s, c);
      }
      return true;
    }
    int p = s.cur;
    if (a == 0) {
      if (!(s.stakes > ante_of(s, p))) return false;
      append(s, 0);
      set_folded(s, p);
      after_move(s, c, true);
    } else if (a == 1) {
      set_ante(s, p, s.stakes);
      s.calls++;
      append(s, 1);
      // terminal(): in round 2 the hand ends when betting is complete; in round 1 it moves to the public card
      if (s.round2 && ready_next(s)) return true;
      after_move(s, c, true);
    } else if (a == 2) {
      if (s.raises >= 2) return false;
      s.stakes += s.round2 ? 4 : 2;
      set_ante(s, p, s.stakes);
      s.raises++;
      s.calls = 0;
      append(s, 2);
      after_move(s, c, false);
    } else {
      return false;
    }
    return true;
  }
  // Tensors (LeducObserver::WriteTensor, leduc_poker.cc:92-192).  which = 0: observation {player(2),
  // private_card(6), community_card(6), pot_contribution(2)}; which = 1: information state {player(2),
  // private_card(6), community_card(6), betting(2x4x2)} with call = 10, raise = 01, fold = 00.
  struct ObsPack { u64 v; int player; int which; };
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg&, int player, int which, ObsPack& p) {
    p.v = pack(s); p.player = player; p.which = which;
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) {
    S s; unpack(s, p.v);
    if (e < 2) return e == p.player ? 1.f : 0.f;
    if (e < 8) return priv_of(s, p.player) == e - 2 ? 1.f : 0.f;
    if (e < 14) return s.pub == e - 8 ? 1.f : 0.f;
    if (p.which == 0) return (float)ante_of(s, e - 14);
    int k = e - 14, round = k >> 3, i = (k >> 1) & 3, bit = k & 1;
    int len = round == 0 ? s.r1len : s.r2len, seq = round == 0 ? s.r1seq : s.r2seq;
    if (i >= len) return 0.f;
    int mv = (seq >> (2 * i)) & 3;
    return (mv == 1 && bit == 0) || (mv == 2 && bit == 1) ? 1.f : 0.f;
  }
};

}  // namespace b2s
