// kuhn_poker (2 players) rule core.  Semantics: reference open_spiel/games/kuhn_poker/kuhn_poker.cc
// (CurrentPlayer :178-185, DoApplyAction :190-229, LegalActions :231-242, Returns :272-283, DidBet :339-349,
// ChanceOutcomes :329-337, observer tensors :72-107).  The reference state is a function of the action
// history, so the packed state IS the history: len (3 bits) | card0 (2) | card1 (2) | bets (1 bit per
// betting action, up to 3), in one uint32.
#pragma once
#include "common.cuh"

namespace b2s {

struct KuhnRules {
  static constexpr int kGameId = B2S_KUHN_POKER;
  typedef u32 Chunk;
  static constexpr int kChunks = 1;
  static constexpr int kMaskWords = 1;
  static constexpr int kPlayers = 2;
  static constexpr int kMaxPath = 0;   // MCTS path stack (>= max_game_length + 2); 0 = no device MCTS
  static constexpr int kMaxLegal = 3;   // most legal actions any state can have (MCTS children block size)
  static constexpr int kFilterWords = 0;   // no per-lane history filter (see rules_go.cuh)
  static constexpr int kIlp = 4;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = true;
  struct Cfg { int dummy; };
  struct S { u32 h; };

  static __host__ const char* make_cfg(const b2s_params& p, Cfg& c, b2s_game_info& gi) {
    c.dummy = 0;
    int n = p.players >= 0 ? p.players : 2;
    if (n != 2) return "kuhn_poker: the device path supports players=2 only";
    gi.num_players = 2;
    gi.num_distinct_actions = 2;                 // kuhn_poker.h:103
    gi.max_chance_outcomes = 3;                  // kuhn_poker.h:105
    gi.max_game_length = 3;                      // kuhn_poker.h:112
    gi.information_state_tensor_size = 11;       // 6n-1, kuhn_poker.cc:395-401
    gi.observation_tensor_size = 7;              // 3n+1, kuhn_poker.cc:403-410
    gi.obs_shape[0] = 7;
    gi.min_utility = -2; gi.max_utility = 2;
    return nullptr;
  }
  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) { s.h = reinterpret_cast<const u32*>(ctx.planes)[i]; }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) { reinterpret_cast<u32*>(ctx.planes)[i] = s.h; }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) { s.h = 0; }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ int len(const S& s) { return s.h & 7; }
  __device__ static __forceinline__ int card(const S& s, int p) { return (s.h >> (3 + 2 * p)) & 3; }
  __device__ static __forceinline__ int bet(const S& s, int k) { return (s.h >> (7 + k)) & 1; }   // k-th betting action
  // Betting sequences that end the game: pp, bp, bb, pbp, pbb.
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg&) {
    int nb = len(s) - 2;
    if (nb < 2) return false;
    if (nb == 2) return !(bet(s, 0) == 0 && bet(s, 1) == 1);
    return true;
  }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg& c) {
    if (terminal(s, c)) return kTerminalPlayerId;
    return len(s) < 2 ? kChancePlayerId : (len(s) & 1);
  }
  // did player p put the extra chip in (kuhn_poker.cc:339-349)
  __device__ static __forceinline__ bool did_bet(const S& s, int p) {
    int nb = len(s) - 2;
    if (nb >= 1 && bet(s, 0)) return p == 0 ? true : (nb >= 2 && bet(s, 1));          // first bettor = player 0
    if (nb >= 2 && bet(s, 1)) return p == 1 ? true : (nb >= 3 && bet(s, 2));          // first bettor = player 1
    return false;
  }
  __device__ static __forceinline__ void returns(const S& s, const Cfg& c, float* r) {
    if (!terminal(s, c)) { r[0] = 0.f; r[1] = 0.f; return; }
    bool b0 = did_bet(s, 0), b1 = did_bet(s, 1);
    int winner;
    if (b0 == b1) winner = card(s, 0) > card(s, 1) ? 0 : 1;      // showdown among equals
    else winner = b0 ? 0 : 1;                                     // the only player who stayed in
    int pot = 2 + (b0 ? 1 : 0) + (b1 ? 1 : 0);
    for (int p = 0; p < 2; ++p) {
      int bt = (p == 0 ? b0 : b1) ? 2 : 1;
      r[p] = p == winner ? (float)(pot - bt) : (float)(-bt);
    }
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg&, u32* m) {
    if (len(s) == 0) m[0] = 7u;
    else if (len(s) == 1) m[0] = 7u & ~(1u << card(s, 0));
    else m[0] = 3u;
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (terminal(s, c)) m[0] = 0; else legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg&, const Ctx&, long long) {
    int l = len(s);
    if (l < 2) {
      if (a < 0 || a > 2 || (l == 1 && a == card(s, 0))) return false;
      s.h |= (u32)a << (3 + 2 * l);
    } else {
      if (a < 0 || a > 1) return false;
      s.h |= (u32)a << (7 + (l - 2));
    }
    s.h = (s.h & ~7u) | (u32)(l + 1);
    return true;
  }
  // Tensors (KuhnObserver::WriteTensor, kuhn_poker.cc:72-107).  which = 0: observation
  // {player(2), private_card(3), pot_contribution(2)}; which = 1: information state
  // {player(2), private_card(3), betting(3x2)}.
  struct ObsPack { u32 h; int player; int which; };
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg&, int player, int which, ObsPack& p) {
    p.h = s.h; p.player = player; p.which = which;
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg&, int e) {
    S s; s.h = p.h;
    if (e < 2) return e == p.player ? 1.f : 0.f;
    if (e < 5) return (len(s) > p.player && card(s, p.player) == e - 2) ? 1.f : 0.f;
    if (p.which == 0) return did_bet(s, e - 5) ? 2.f : 1.f;          // ante_[p]
    int k = (e - 5) >> 1, act = (e - 5) & 1;
    return (len(s) - 2 > k && bet(s, k) == act) ? 1.f : 0.f;
  }
};

}  // namespace b2s
