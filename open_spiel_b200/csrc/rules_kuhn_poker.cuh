// kuhn_poker (2..5 players) rule core.  Semantics: reference open_spiel/games/kuhn_poker/kuhn_poker.cc
// (CurrentPlayer :178-185, DoApplyAction :190-229, LegalActions :231-242, Returns :272-283, DidBet :339-349,
// ChanceOutcomes :329-337, observer tensors :72-107).  The reference state is a function of the action
// history, so the packed state IS the history, in one uint32:
//   len (bits 0-4) | card of player p (3 bits at 5 + 3p, p < 5) | betting action k (1 bit at 20 + k, k < 2n-1 <= 9)
// n players draw from n+1 cards; betting action k is made by player k mod n.
#pragma once
#include "common.cuh"

namespace b2s {

struct KuhnRules {
  static constexpr int kGameId = B2S_KUHN_POKER;
  typedef u32 Chunk;
  static constexpr int kChunks = 1;
  static constexpr int kMaskWords = 1;
  static constexpr int kPlayers = 5;   // most players the packed layout holds (returns arrays); the actual count is Cfg::n
  static constexpr int kMaxPath = 0;   // MCTS path stack (>= max_game_length + 2); 0 = no device MCTS
  static constexpr int kMaxLegal = 6;   // most legal actions any state can have (MCTS children block size)
  static constexpr int kFilterWords = 0;   // no per-lane history filter (see rules_go.cuh)
  static constexpr int kIlp = 4;
  static constexpr int kMinBlocks = 4;
  static constexpr bool kHasInfoState = true;

  struct Cfg { int n; };
  struct S { u32 h; };

  static __host__ const char* make_cfg(const b2s_params& p, Cfg& c, b2s_game_info& gi) {
    c.n = p.players >= 0 ? p.players : 2;
    if (c.n < 2 || c.n > kPlayers) return "kuhn_poker: the packed layout holds players = 2..5";
    gi.num_players = c.n;
    gi.num_distinct_actions = 2;                 // kuhn_poker.h:107
    gi.max_chance_outcomes = c.n + 1;            // kuhn_poker.h:111
    gi.max_game_length = 2 * c.n - 1;            // kuhn_poker.h:121
    gi.information_state_tensor_size = 6 * c.n - 1;   // kuhn_poker.cc:395-401
    gi.observation_tensor_size = 3 * c.n + 1;         // kuhn_poker.cc:403-410
    gi.obs_shape[0] = 3 * c.n + 1;
    gi.min_utility = -2; gi.max_utility = 2 * (c.n - 1);
    return nullptr;
  }
  __device__ static __forceinline__ int num_players(const Cfg& c) { return c.n; }

  __device__ static __forceinline__ void load(S& s, const Ctx& ctx, long long i) { s.h = reinterpret_cast<const u32*>(ctx.planes)[i]; }
  __device__ static __forceinline__ void store(const S& s, const Ctx& ctx, long long i) { reinterpret_cast<u32*>(ctx.planes)[i] = s.h; }
  __device__ static __forceinline__ void init(S& s, const Cfg&, const Ctx&, long long) { s.h = 0; }
  __device__ static __forceinline__ void copy_history(const Ctx&, long long, const Ctx&, long long, const S&, const Cfg&) {}

  __device__ static __forceinline__ int len(const S& s) { return s.h & 31; }
  __device__ static __forceinline__ int card(const S& s, int p) { return (s.h >> (5 + 3 * p)) & 7; }
  __device__ static __forceinline__ int bet(const S& s, int k) { return (s.h >> (20 + k)) & 1; }   // k-th betting action
  __device__ static __forceinline__ int num_bet_actions(const S& s, const Cfg& c) { int k = len(s) - c.n; return k > 0 ? k : 0; }
  // player of the first bet (first_bettor_), -1 if nobody has bet
  __device__ static __forceinline__ int first_bettor(const S& s, const Cfg& c) {
    const u32 bets = (s.h >> 20) & ((1u << num_bet_actions(s, c)) - 1u);
    return bets ? (__ffs((int)bets) - 1) % c.n : -1;        // a first bet can only come in the first round: action k, player k
  }
  // The game ends after n passes, or once everybody after the first bettor has answered (kuhn_poker.cc:216-228).
  __device__ static __forceinline__ bool terminal(const S& s, const Cfg& c) {
    const int k = num_bet_actions(s, c), fb = first_bettor(s, c);
    return fb < 0 ? k == c.n : k == c.n + fb;
  }
  __device__ static __forceinline__ int cur_player(const S& s, const Cfg& c) {
    if (terminal(s, c)) return kTerminalPlayerId;
    return len(s) < c.n ? kChancePlayerId : (len(s) % c.n);
  }
  // did player p put the extra chip in (kuhn_poker.cc:339-349)
  __device__ static __forceinline__ bool did_bet(const S& s, const Cfg& c, int p) {
    const int fb = first_bettor(s, c);
    if (fb < 0) return false;
    if (p == fb) return true;
    const int k = p > fb ? p : c.n + p;                     // the action in which p answered the bet
    return k < num_bet_actions(s, c) && bet(s, k);
  }
  // pot contribution of player p so far (ante_[p]): 1 + one chip per bet / call
  __device__ static __forceinline__ int ante(const S& s, const Cfg& c, int p) {
    const int k = num_bet_actions(s, c);
    int a = 1;
    if (p < k && bet(s, p)) ++a;
    if (c.n + p < k && bet(s, c.n + p)) ++a;
    return a;
  }
  __device__ static __forceinline__ void returns(const S& s, const Cfg& c, float* r) {
    for (int p = 0; p < c.n; ++p) r[p] = 0.f;
    if (!terminal(s, c)) return;
    // winner: the highest card among the players who bet, or among all players when nobody did (:216-228)
    const bool any_bet = first_bettor(s, c) >= 0;
    int winner = -1, best_card = -1, pot = c.n;
    for (int p = 0; p < c.n; ++p) {
      const bool in = !any_bet || did_bet(s, c, p);
      if (did_bet(s, c, p)) ++pot;
      if (in && card(s, p) > best_card) { best_card = card(s, p); winner = p; }
    }
    for (int p = 0; p < c.n; ++p) {
      const int bt = did_bet(s, c, p) ? 2 : 1;
      r[p] = p == winner ? (float)(pot - bt) : (float)(-bt);
    }
  }
  __device__ static __forceinline__ void legal_nonterminal(const S& s, const Cfg& c, u32* m) {
    const int l = len(s);
    if (l < c.n) {
      u32 deck = (1u << (c.n + 1)) - 1u;
      for (int p = 0; p < l; ++p) deck &= ~(1u << card(s, p));
      m[0] = deck;
    } else {
      m[0] = 3u;
    }
  }
  __device__ static __forceinline__ void legal(const S& s, const Cfg& c, u32* m) {
    if (terminal(s, c)) m[0] = 0; else legal_nonterminal(s, c, m);
  }
  __device__ static __forceinline__ bool apply(S& s, int a, const Cfg& c, const Ctx&, long long) {
    const int l = len(s);
    if (l < c.n) {
      if (a < 0 || a > c.n) return false;
      for (int p = 0; p < l; ++p) if (card(s, p) == a) return false;
      s.h |= (u32)a << (5 + 3 * l);
    } else {
      if (a < 0 || a > 1) return false;
      s.h |= (u32)a << (20 + (l - c.n));
    }
    s.h = (s.h & ~31u) | (u32)(l + 1);
    return true;
  }

  // Tensors (KuhnObserver::WriteTensor, kuhn_poker.cc:72-107).  which = 0: observation
  // {player(n), private_card(n+1), pot_contribution(n)}; which = 1: information state
  // {player(n), private_card(n+1), betting(2n-1 x 2)}.
  struct ObsPack { u32 h; int player; int which; };
  __device__ static __forceinline__ void obs_pack(const S& s, const Cfg&, int player, int which, ObsPack& p) {
    p.h = s.h; p.player = player; p.which = which;
  }
  __device__ static __forceinline__ float obs_elem(const ObsPack& p, const Cfg& c, int e) {
    S s; s.h = p.h;
    if (e < c.n) return e == p.player ? 1.f : 0.f;
    if (e < 2 * c.n + 1) return (len(s) > p.player && card(s, p.player) == e - c.n) ? 1.f : 0.f;
    const int j = e - (2 * c.n + 1);
    if (p.which == 0) return (float)ante(s, c, j);
    const int k = j >> 1, act = j & 1;
    return (num_bet_actions(s, c) > k && bet(s, k) == act) ? 1.f : 0.f;
  }
};

}  // namespace b2s
