// leduc_poker (2 players) rule core.  Semantics: reference open_spiel/games/leduc_poker/leduc_poker.cc
// (DoApplyAction :298-414, LegalActions :416-457, IsTerminal :498-500, Returns :502-514, NextPlayer :573-591,
// RankHand :593-626, ResolveWinner :628-678, ReadyForNextRound :680-683, observer tensors :92-192).
// Packed into one uint64, split so that no field straddles the two 32-bit halves (every read is one bit-field extract
// and the state is never unpacked: a step touches only the fields its branch needs).  field: bits
//   lo: priv0 0-2, priv1 3-5, pub 6-8 (7 = not dealt) | r1len 9-11, r1seq 12-19 | r2len 20-22, r2seq 23-30
//   hi: cur 0-1 (0,1 player; 2 chance) | round2 2 | calls 3-4 | raises 5-6 | stakes 7-10 | ante0 11-14, ante1 15-18 |
//       folded0 19, folded1 20 | dealt 21-22
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
  static constexpr int kIlp = 4;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = true;
  struct Cfg { int starting_player; };
  struct S { u32 lo, hi; };
  static constexpr int kNone = 7, kChance = 2;
  // bit positions in hi
  static constexpr int kCur = 0, kRound2 = 2, kCalls = 3, kRaises = 5, kStakes = 7, kAnte = 11, kFolded = 19, kDealt = 21;

  __device__ static __forceinline__ int bits(u32 w, int sh, int n) { return (int)((w >> sh) & ((1u << n) - 1u)); }
  __device__ static __forceinline__ void put(u32& w, int sh, int n, int v) { w = (w & ~(((1u << n) - 1u) << sh)) | ((u32)v << sh); }
  __device__ static __forceinline__ int priv_of(const S& s, int p) { return bits(s.lo, 3 * p, 3); }
  __device__ static __forceinline__ int pub(const S& s) { return bits(s.lo, 6, 3); }
  __device__ static __forceinline__ int seq_len(const S& s, int round) { return bits(s.lo, round ? 20 : 9, 3); }
  __device__ static __forceinline__ int seq(const S& s, int round) { return bits(s.lo, round ? 23 : 12, 8); }
  __device__ static __forceinline__ int cur(const S& s) { return bits(s.hi, kCur, 2); }
  __device__ static __forceinline__ int round2(const S& s) { return bits(s.hi, kRound2, 1); }
  __device__ static __forceinline__ int calls(const S& s) { return bits(s.hi, kCalls, 2); }
  __device__ static __forceinline__ int raises(const S& s) { return bits(s.hi, kRaises, 2); }
  __device__ static __forceinline__ int stakes(const S& s) { return bits(s.hi, kStakes, 4); }
  __device__ static __forceinline__ int ante_of(const S& s, int p) { return bits(s.hi, kAnte + 4 * p, 4); }
  __device__ static __forceinline__ int folded_of(const S& s, int p) { return bits(s.hi, kFolded + p, 1); }
  __device__ static __forceinline__ int dealt(const S& s) { return bits(s.hi, kDealt, 2); }

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
  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) {
    const uint2 v = reinterpret_cast<const uint2*>(ctx.planes)[i];
    s.lo = v.x; s.hi = v.y;
  }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) { reinterpret_cast<uint2*>(ctx.planes)[i] = make_uint2(s.lo, s.hi); }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) {
    s.lo = (u32)kNone | (u32)kNone << 3 | (u32)kNone << 6;                              // no cards, empty betting sequences
    s.hi = (u32)kChance << kCur | 1u << kStakes | 1u << kAnte | 1u << (kAnte + 4);      // chance to deal, stakes 1, antes 1 / 1
  }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ int remaining(const S& s) { return 2 - __popc(s.hi & (3u << kFolded)); }
  __device__ static __forceinline__ bool ready_next(const S& s) {                      // leduc_poker.cc:680-683
    return calls(s) + (raises(s) > 0 ? 1 : 0) == remaining(s);
  }
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) {
    return remaining(s) == 1 || (round2(s) && cur(s) != kChance && ready_next(s) && pub(s) != kNone);
  }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg& c) {
    if (terminal(s, c)) return kTerminalPlayerId;
    return cur(s) == kChance ? kChancePlayerId : cur(s);
  }
  __device__ static __forceinline__ int rank(const S& s, int p) {
    int lo = pub(s), hi = priv_of(s, p);
    if (lo > hi) { int t = lo; lo = hi; hi = t; }
    if ((lo & 1) == 0 && hi == lo + 1) return 36 + lo;
    return (hi >> 1) * 6 + (lo >> 1);
  }
  __device__ static __forceinline__ void returns(const S& s, const Cfg& c, float* r) {
    r[0] = 0.f; r[1] = 0.f;
    if (!terminal(s, c)) return;
    const int a0 = ante_of(s, 0), a1 = ante_of(s, 1), pot = a0 + a1;
    int w;                                  // winner, or -1 for a split pot
    if (remaining(s) == 1) w = folded_of(s, 0) ? 1 : 0;
    else { int r0 = rank(s, 0), r1 = rank(s, 1); w = r0 == r1 ? -1 : (r0 > r1 ? 0 : 1); }
    if (w < 0) {                            // split pot: money += pot / 2.0 (leduc_poker.cc:670-676)
      r[0] = (float)pot * 0.5f - (float)a0;
      r[1] = (float)pot * 0.5f - (float)a1;
    } else {                                // the winner takes the other player's ante
      r[0] = w == 0 ? (float)a1 : (float)-a0;
      r[1] = w == 1 ? (float)a0 : (float)-a1;
    }
  }
  // cards still in the deck: a card field holding kNone (7) shifts its bit out of the six-card mask
  __device__ static __forceinline__ u32 deck(const S& s) {
    return 63u & ~(1u << priv_of(s, 0) | 1u << priv_of(s, 1) | 1u << pub(s));
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg&, u32* m) {
    if (cur(s) == kChance) { m[0] = deck(s); return; }
    u32 v = 2u;                                                     // call always
    if (stakes(s) > ante_of(s, cur(s))) v |= 1u;                    // fold only under pressure
    if (raises(s) < 2) v |= 4u;
    m[0] = v;
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (terminal(s, c)) m[0] = 0; else legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg& c, const Ctx&, long long) {
    const int p = cur(s);
    if (p == kChance) {
      if ((unsigned)a > 5u || !((deck(s) >> a) & 1u)) return false;
      const int d = dealt(s);                                       // 0, 1: the private cards; 2: the public card
      put(s.lo, 3 * d, 3, a);
      if (d < 2) {
        put(s.hi, kDealt, 2, d + 1);
        if (d == 1) put(s.hi, kCur, 2, c.starting_player);
      } else {
        // NextPlayer() from the chance node (leduc_poker.cc:573-591) starts after starting_player + 1: with nobody
        // folded (a fold ends a two-player hand) that is starting_player
        put(s.hi, kCur, 2, c.starting_player);
      }
      return true;
    }
    if ((unsigned)a > 2u) return false;
    const int st = stakes(s), r2 = round2(s);
    if (a == 0 ? !(st > ante_of(s, p)) : (a == 2 && raises(s) >= 2)) return false;
    {                                                               // sequence_append_move
      const int lsh = r2 ? 20 : 9, ssh = r2 ? 23 : 12;
      s.lo |= (u32)a << (ssh + 2 * bits(s.lo, lsh, 3));
      s.lo += 1u << lsh;
    }
    if (a == 0) {                                                   // fold: the other player wins, cur_player_ stays
      s.hi |= 1u << (kFolded + p);
      return true;
    }
    if (a == 1) {
      put(s.hi, kAnte + 4 * p, 4, st);
      s.hi += 1u << kCalls;
      if (ready_next(s)) {
        if (r2) return true;                                        // betting over in round 2: terminal, cur_player_ stays
        // NewRound(): round 2, counters cleared, chance deals the public card (leduc_poker.cc:385-396)
        s.hi = (s.hi & ~(3u << kCalls | 3u << kRaises | 3u << kCur)) | 1u << kRound2 | (u32)kChance << kCur;
        return true;
      }
    } else {
      const int ns = st + (r2 ? 4 : 2);
      put(s.hi, kStakes, 4, ns);
      put(s.hi, kAnte + 4 * p, 4, ns);
      s.hi = (s.hi & ~(3u << kCalls)) + (1u << kRaises);
    }
    put(s.hi, kCur, 2, 1 - p);                                      // NextPlayer(): the other player (nobody has folded)
    return true;
  }
  // Tensors (LeducObserver::WriteTensor, leduc_poker.cc:92-192).  which = 0: observation {player(2),
  // private_card(6), community_card(6), pot_contribution(2)}; which = 1: information state {player(2),
  // private_card(6), community_card(6), betting(2x4x2)} with call = 10, raise = 01, fold = 00.
  struct ObsPack { S s; int player; int which; };
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg&, int player, int which, ObsPack& p) {
    p.s = s; p.player = player; p.which = which;
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) {
    const S& s = p.s;
    if (e < 2) return e == p.player ? 1.f : 0.f;
    if (e < 8) return priv_of(s, p.player) == e - 2 ? 1.f : 0.f;
    if (e < 14) return pub(s) == e - 8 ? 1.f : 0.f;
    if (p.which == 0) return (float)ante_of(s, e - 14);
    int k = e - 14, round = k >> 3, i = (k >> 1) & 3, bit = k & 1;
    if (i >= seq_len(s, round)) return 0.f;
    int mv = (seq(s, round) >> (2 * i)) & 3;
    return (mv == 1 && bit == 0) || (mv == 2 && bit == 1) ? 1.f : 0.f;
  }
};

}  // namespace b2s
