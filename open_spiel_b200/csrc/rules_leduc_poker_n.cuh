// leduc_poker with 3 or 4 players.  Semantics: reference open_spiel/games/leduc_poker/leduc_poker.cc (DoApplyAction :298-414,
// LegalActions :416-457, IsTerminal :498-500, Returns :502-514, NextPlayer :573-591, RankHand :593-626, ResolveWinner
// :628-678, ReadyForNextRound :680-683, observer tensors :92-192) for num_players_ > 2; the two-player game keeps its own
// 8-byte core (rules_leduc_poker.cuh).  n players, 2 (n + 1) cards, at most 3n - 2 bets per round.
// Packed into one 128-bit chunk {lo, hi} (field: bits):
//   lo: private card of player p 4 bits at 4p (15 = not dealt) | public 16-19 | r1len 20-23 | r1seq 24-43 (2 bits per bet)
//       | cur 44-46 (player, 7 = chance) | round2 47 | calls 48-50 | raises 51-52 | stakes 53-56 | dealt 57-59
//   hi: r2len 0-3 | r2seq 4-23 | ante of player p 4 bits at 24 + 4p | folded 40-43
#pragma once
#include "common.cuh"

namespace b2s {

struct LeducNRules {
  static constexpr int kGameId = B2S_LEDUC_POKER;
  typedef uint4 Chunk;
  static constexpr int kChunks = 1;
  static constexpr int kMaskWords = 1;
  static constexpr int kPlayers = 4;      // most players the layout holds; the actual count is Cfg::n
  static constexpr int kMaxPath = 0;      // no device MCTS (chance nodes, imperfect information)
  static constexpr int kMaxLegal = 10;
  static constexpr int kFilterWords = 0;
  static constexpr int kIlp = 2;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = true;
  static constexpr int kNone = 15, kChance = 7;

  struct Cfg { int n, starting_player, cards, max_bets; };
  struct S {
    int priv[4], pub, r1len, r2len, cur, round2, calls, raises, stakes, dealt, folded;
    int ante[4];
    u32 r1seq, r2seq;
  };

  static __host__ const char* make_cfg(const b2s_params& p, Cfg& c, b2s_game_info& gi) {
    c.n = p.players >= 0 ? p.players : 2;
    if (c.n < 3 || c.n > kPlayers) return "leduc_poker: this rule core holds players = 3..4";
    c.starting_player = p.starting_player >= 0 ? p.starting_player : 0;
    if (c.starting_player >= c.n) return "leduc_poker: starting_player out of range";
    c.cards = 2 * (c.n + 1);
    c.max_bets = 3 * c.n - 2;
    gi.num_players = c.n;
    gi.num_distinct_actions = 3;
    gi.max_chance_outcomes = c.cards;
    gi.max_game_length = 2 * c.max_bets;                                       // leduc_poker.h:233-241
    gi.information_state_tensor_size = c.n + 2 * c.cards + 2 * c.max_bets * 2;   // leduc_poker.cc:811-820
    gi.observation_tensor_size = c.n + 2 * c.cards + c.n;                      // leduc_poker.cc:822-831
    gi.obs_shape[0] = gi.observation_tensor_size;
    gi.max_utility = (c.n - 1) * 13; gi.min_utility = -13;                     // leduc_poker.cc:833-841
    return nullptr;
  }
  __device__ static __forceinline__ int num_players(const Cfg& c) { return c.n; }

  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) {
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(ctx.planes)[i];
    const u64 lo = v.x, hi = v.y;
    for (int p = 0; p < 4; ++p) { s.priv[p] = (int)((lo >> (4 * p)) & 15); s.ante[p] = (int)((hi >> (24 + 4 * p)) & 15); }
    s.pub = (int)((lo >> 16) & 15); s.r1len = (int)((lo >> 20) & 15); s.r1seq = (u32)((lo >> 24) & 0xFFFFFu);
    s.cur = (int)((lo >> 44) & 7); s.round2 = (int)((lo >> 47) & 1); s.calls = (int)((lo >> 48) & 7);
    s.raises = (int)((lo >> 51) & 3); s.stakes = (int)((lo >> 53) & 15); s.dealt = (int)((lo >> 57) & 7);
    s.r2len = (int)(hi & 15); s.r2seq = (u32)((hi >> 4) & 0xFFFFFu); s.folded = (int)((hi >> 40) & 15);
  }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) {
    u64 lo = (u64)s.pub << 16 | (u64)s.r1len << 20 | (u64)s.r1seq << 24 | (u64)s.cur << 44 | (u64)s.round2 << 47 |
             (u64)s.calls << 48 | (u64)s.raises << 51 | (u64)s.stakes << 53 | (u64)s.dealt << 57;
    u64 hi = (u64)s.r2len | (u64)s.r2seq << 4 | (u64)s.folded << 40;
    for (int p = 0; p < 4; ++p) { lo |= (u64)s.priv[p] << (4 * p); hi |= (u64)s.ante[p] << (24 + 4 * p); }
    reinterpret_cast<ulonglong2*>(ctx.planes)[i] = make_ulonglong2(lo, hi);
  }
  __device__ static __forceinline__ void init(S& s, const Cfg& c, const Ctx&, long long) {
    for (int p = 0; p < 4; ++p) { s.priv[p] = kNone; s.ante[p] = p < c.n ? 1 : 0; }
    s.pub = kNone; s.r1len = s.r2len = 0; s.r1seq = s.r2seq = 0;
    s.cur = kChance; s.round2 = 0; s.calls = 0; s.raises = 0; s.stakes = 1; s.dealt = 0; s.folded = 0;
  }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ int remaining(const S& s, const Cfg& c) { return c.n - __popc((unsigned)s.folded); }
  __device__ static __forceinline__ bool ready_next(const S& s, const Cfg& c) {       // ReadyForNextRound :680-683
    const int rem = remaining(s, c);
    return (s.raises == 0 && s.calls == rem) || (s.raises > 0 && s.calls == rem - 1);
  }
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg& c) {
    return remaining(s, c) == 1 || (s.round2 && s.cur != kChance && ready_next(s, c) && s.pub != kNone);
  }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg& c) {
    if (terminal(s, c)) return kTerminalPlayerId;
    return s.cur == kChance ? kChancePlayerId : s.cur;
  }
  __device__ static __forceinline__ int rank(const S& s, const Cfg& c, int p) {        // RankHand :593-626
    int lo = s.pub, hi = s.priv[p];
    if (lo > hi) { int t = lo; lo = hi; hi = t; }
    if ((lo & 1) == 0 && hi == lo + 1) return c.cards * c.cards + lo;                  // a pair
    return (hi >> 1) * c.cards + (lo >> 1);
  }
  __device__ static __forceinline__ void returns(const S& s, const Cfg& c, float* r) {
    for (int p = 0; p < c.n; ++p) r[p] = 0.f;
    if (!terminal(s, c)) return;
    int pot = 0;
    for (int p = 0; p < c.n; ++p) pot += s.ante[p];
    unsigned winners = 0;
    if (remaining(s, c) == 1) {
      winners = ~(unsigned)s.folded & ((1u << c.n) - 1u);
    } else {                                                                            // ResolveWinner :628-678
      int best = -1;
      for (int p = 0; p < c.n; ++p) {
        if ((s.folded >> p) & 1) continue;
        int rk = rank(s, c, p);
        if (rk > best) { best = rk; winners = 1u << p; }
        else if (rk == best) winners |= 1u << p;
      }
    }
    const float share = (float)pot / (float)__popc(winners);      // at most two players tie (two cards per rank): exact
    for (int p = 0; p < c.n; ++p) r[p] = ((winners >> p) & 1 ? share : 0.f) - (float)s.ante[p];
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg& c, u32* m) {
    if (s.cur == kChance) {
      u32 deck = (1u << c.cards) - 1u;
      for (int p = 0; p < c.n; ++p) if (s.priv[p] != kNone) deck &= ~(1u << s.priv[p]);
      if (s.pub != kNone) deck &= ~(1u << s.pub);
      m[0] = deck;
      return;
    }
    u32 v = 2u;                                                   // call always
    if (s.stakes > s.ante[s.cur]) v |= 1u;                        // fold only under pressure
    if (s.raises < 2) v |= 4u;
    m[0] = v;
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (terminal(s, c)) m[0] = 0; else legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ int next_player(const S& s, const Cfg& c) {        // NextPlayer :573-591
    const int from = s.cur == kChance ? (c.starting_player + c.n - 1) % c.n : s.cur;
    for (int i = 1; i <= c.n; ++i) {
      int p = (from + i) % c.n;
      if (!((s.folded >> p) & 1)) return p;
    }
    return from;
  }
  __device__ static __forceinline__ void append(S& s, int mv) {
    if (!s.round2) { s.r1seq |= (u32)mv << (2 * s.r1len); s.r1len++; }
    else { s.r2seq |= (u32)mv << (2 * s.r2len); s.r2len++; }
  }
  __device__ static __forceinline__ void after_move(S& s, const Cfg& c, bool may_advance) {
    if (terminal_after(s, c)) return;
    if (may_advance && ready_next(s, c)) { s.round2 = 1; s.raises = 0; s.calls = 0; s.cur = kChance; }
    else s.cur = next_player(s, c);
  }
  // IsTerminal as the reference evaluates it inside DoApplyAction (round_ == 2 && ReadyForNextRound, no chance pending)
  __device__ static __forceinline__ bool terminal_after(const S& s, const Cfg& c) {
    return remaining(s, c) == 1 || (s.round2 && ready_next(s, c));
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg& c, const Ctx&, long long) {
    if (s.cur == kChance) {
      u32 m; legal_nonterminal(s, c, &m);
      if (a < 0 || a >= c.cards || !((m >> a) & 1u)) return false;
      if (s.dealt < c.n) {
        s.priv[s.dealt] = a;
        s.dealt++;
        if (s.dealt == c.n) s.cur = c.starting_player;
      } else {
        s.pub = a;
        s.cur = next_player(s, c);
      }
      return true;
    }
    const int p = s.cur;
    if (a == 0) {
      if (!(s.stakes > s.ante[p])) return false;
      append(s, 0);
      s.folded |= 1 << p;
      after_move(s, c, true);
    } else if (a == 1) {
      s.ante[p] = s.stakes;
      s.calls++;
      append(s, 1);
      after_move(s, c, true);
    } else if (a == 2) {
      if (s.raises >= 2) return false;
      s.stakes += s.round2 ? 4 : 2;
      s.ante[p] = s.stakes;
      s.raises++;
      s.calls = 0;
      append(s, 2);
      after_move(s, c, false);
    } else {
      return false;
    }
    return true;
  }

  // Tensors (LeducObserver::WriteTensor, leduc_poker.cc:92-192).  which = 0: observation {player(n), private_card(cards),
  // community_card(cards), pot_contribution(n)}; which = 1: information state {player(n), private_card, community_card,
  // betting(2 x max_bets x 2)} with call = 10, raise = 01, fold = 00.
  struct ObsPack { S s; int player; int which; };
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg&, int player, int which, ObsPack& p) {
    p.s = s; p.player = player; p.which = which;
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg& c, int e) {
    const S& s = p.s;
    if (e < c.n) return e == p.player ? 1.f : 0.f;
    e -= c.n;
    if (e < c.cards) return s.priv[p.player] == e ? 1.f : 0.f;
    e -= c.cards;
    if (e < c.cards) return s.pub == e ? 1.f : 0.f;
    e -= c.cards;
    if (p.which == 0) return (float)s.ante[e];
    const int round = e / (2 * c.max_bets), k = e - round * 2 * c.max_bets, i = k >> 1, bit = k & 1;
    const int len = round == 0 ? s.r1len : s.r2len;
    const u32 seq = round == 0 ? s.r1seq : s.r2seq;
    if (i >= len) return 0.f;
    const int mv = (int)((seq >> (2 * i)) & 3u);
    return (mv == 1 && bit == 0) || (mv == 2 && bit == 1) ? 1.f : 0.f;
  }
};

}  // namespace b2s
