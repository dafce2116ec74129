// Device-resident MCTS: one search tree per thread (root parallelism), nodes in a global-memory arena.
// Semantics: reference open_spiel/algorithms/mcts.cc — ApplyTreePolicy :273-351 (expand on the second visit,
// children in a random order, first-max UCT selection), RandomRolloutEvaluator::Evaluate :43-72, MCTSearch
// :353-467 (backup from the point of view of the player who chose each node, MCTS-Solver propagation, early
// exit when the root is proven or has a single child), BestChild / CompareFinal :114-143.
// A single tree is inherently sequential (every simulation sees the statistics of all earlier ones), so one
// thread runs one tree exactly in the reference's order and throughput comes from running thousands of
// independent roots per GPU.  Random decisions are an explicit function of (seed, tree, simulation, position)
// through Philox (common.cuh), the same function oracle/algorithms/mcts.cc uses, so trees match bit for bit:
//   expansion #e:  Fisher-Yates over the ascending legal list, j = rng(key, e, i, 1, i+1) for i = n-1..1
//   simulation #t, rollout #r, ply p:  k = rng(key, t, p + 4096 q, 2+r, C) over the C playout candidates
//     (the legal actions; for go: empty non-ko points + pass), q = 0,1,.. until the candidate is legal
// UCT arithmetic is done with explicitly rounded double operations (no FMA contraction) and log(N_parent)
// comes from a table the HOST fills with std::log, so values equal the CPU's to the last bit.
// Both child selection policies (UCT, PUCT with the rollout evaluator's uniform prior) are implemented.
// Not implemented: chance nodes in the tree, Dirichlet noise, the reference's node-budget garbage
// collection (mcts.cc:441-482) — a tree that exhausts the arena stops and is reported as an error.
#pragma once
#include "common.cuh"

namespace b2s {

struct __align__(16) MctsNode {     // SearchNode (mcts.h:114-146) without the heap vectors: 32 B
  double total_reward;
  u32 visits;                       // explore_count
  u32 first_child;                  // arena index of the first child; children are contiguous
  float out0, out1;                 // proven outcome (returns) when has_outcome
  short action;
  unsigned char nchild;
  signed char player;               // the player who chose `action`
  unsigned char has_outcome;
  unsigned char pad[3];
};

struct MctsArgs {
  int sims, n_rollouts, solve, num_actions, mask_words, max_plies, puct;
  double uct_c, max_utility;
  u64 seed;
  long long tree_offset;
  const double* log_table;          // log_table[k] = std::log((double)k), k <= sims (host-computed)
  MctsNode* pool;
  unsigned long long* pool_top;
  unsigned long long pool_cap;
  int* visits_out;                  // [n][A]
  double* reward_out;               // [n][A]
  float* outcome_out;               // [n][A] (NaN = unproven), nullable
  int* best_out;                    // [n]
  int* sims_out;                    // [n], nullable
  ErrBuf* err;
};

__device__ __forceinline__ u32 rng_uniform(u64 key, u32 a, u32 b, u32 c, u32 n) {
  return philox_uniform(key, (u64)a | ((u64)b << 32), c, n);
}

__device__ __forceinline__ double uct_value(const MctsNode& ch, u32 parent_visits, const MctsArgs& P) {
  if (ch.has_outcome) return (double)(ch.player == 0 ? ch.out0 : ch.out1);
  if (ch.visits == 0) return __longlong_as_double(0x7ff0000000000000LL);
  double n = (double)ch.visits;
  double q = __ddiv_rn(ch.total_reward, n);
  double u = __dsqrt_rn(__ddiv_rn(P.log_table[parent_visits], n));
  return __dadd_rn(q, __dmul_rn(P.uct_c, u));
}

// PUCTValue (mcts.cc:103-112) with RandomRolloutEvaluator's uniform prior 1/|children| (mcts.cc:74-87); `cp` is
// (uct_c * prior) * sqrt(N_parent), the part shared by all children of one parent, in the reference's
// left-to-right evaluation order.
__device__ __forceinline__ double puct_value(const MctsNode& ch, double cp) {
  if (ch.has_outcome) return (double)(ch.player == 0 ? ch.out0 : ch.out1);
  double q = ch.visits ? __ddiv_rn(ch.total_reward, (double)ch.visits) : 0.0;
  return __dadd_rn(q, __ddiv_rn(cp, (double)(ch.visits + 1u)));
}

template <class R, int MAXPATH, int MINBLOCKS>
__global__ void __launch_bounds__(128, MINBLOCKS) k_mcts(Ctx rootctx, Ctx workctx, typename R::Cfg cfg, MctsArgs P, long long n_trees) {
  long long tree = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tree >= n_trees) return;
  typename R::S root;
  R::load(root, rootctx, tree);
  MctsNode* pool = P.pool;
  const u64 key = P.seed + (u64)(tree + P.tree_offset) * 0x9E3779B97F4A7C15ull;
  const u32 root_idx = (u32)tree;                  // the first n_trees arena slots are the roots
  for (int a = 0; a < P.num_actions; ++a) {
    P.visits_out[tree * P.num_actions + a] = 0;
    P.reward_out[tree * P.num_actions + a] = 0.0;
    if (P.outcome_out) P.outcome_out[tree * P.num_actions + a] = __int_as_float(0x7fc00000);
  }
  P.best_out[tree] = -1;
  if (P.sims_out) P.sims_out[tree] = 0;
  if (R::terminal(root, cfg)) return;              // nothing to search (the reference would index returns[-4])
  {
    MctsNode r;
    r.total_reward = 0; r.visits = 0; r.first_child = 0; r.out0 = r.out1 = 0; r.action = -1; r.nchild = 0;
    r.player = (signed char)R::cur_player(root, cfg); r.has_outcome = 0; r.pad[0] = r.pad[1] = r.pad[2] = 0;
    pool[root_idx] = r;
  }
  u32 path[MAXPATH];
  u32 expansions = 0;
  int sim = 0;
  bool failed = false;
  for (; sim < P.sims && !failed; ++sim) {
    typename R::S s = root;
    int depth = 0;
    u32 cur = root_idx;
    path[depth++] = cur;
    bool term = false;
    // ---- tree policy (mcts.cc:273-351) ----
    while (!term && pool[cur].visits > 0) {
      if (pool[cur].nchild == 0) {
        u32 m[R::kMaskWords];
        R::legal_nonterminal(s, cfg, m);
        int n = 0;
        for (int w = 0; w < P.mask_words; ++w) n += __popc(m[w]);
        unsigned long long base = atomicAdd(P.pool_top, (unsigned long long)n);
        if (base + n > P.pool_cap || depth >= MAXPATH - 1) { failed = true; break; }
        signed char player = (signed char)R::cur_player(s, cfg);
        int k = 0;
        for (int w = 0; w < P.mask_words; ++w) {
          u32 bits = m[w];
          while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            MctsNode c;
            c.total_reward = 0; c.visits = 0; c.first_child = 0; c.out0 = c.out1 = 0;
            c.action = (short)(w * 32 + b); c.nchild = 0; c.player = player; c.has_outcome = 0;
            c.pad[0] = c.pad[1] = c.pad[2] = 0;
            pool[base + k++] = c;
          }
        }
        u32 e = expansions++;
        for (int i = n - 1; i >= 1; --i) {          // random child order (std::shuffle's role, mcts.cc:294)
          u32 j = rng_uniform(key, e, (u32)i, 1u, (u32)(i + 1));
          short t = pool[base + i].action; pool[base + i].action = pool[base + j].action; pool[base + j].action = t;
        }
        pool[cur].first_child = (u32)base;
        pool[cur].nchild = (unsigned char)n;
      }
      u32 first = pool[cur].first_child, pv = pool[cur].visits;
      int nch = pool[cur].nchild;
      double best = __longlong_as_double(0xfff0000000000000LL);
      u32 chosen = first;
      if (P.puct) {
        double cp = __dmul_rn(__dmul_rn(P.uct_c, __ddiv_rn(1.0, (double)nch)), __dsqrt_rn((double)pv));
        for (int i = 0; i < nch; ++i) {
          double v = puct_value(pool[first + i], cp);
          if (v > best) { best = v; chosen = first + i; }
        }
      } else {
        for (int i = 0; i < nch; ++i) {
          double v = uct_value(pool[first + i], pv, P);
          if (v > best) { best = v; chosen = first + i; }
        }
      }
      cur = chosen;
      apply_known_legal<R>(s, (int)pool[cur].action, cfg, workctx, tree);
      path[depth++] = cur;
      term = R::terminal(s, cfg);
    }
    if (failed) break;
    // ---- evaluate (mcts.cc:372-381) ----
    double ret[2];
    bool solved;
    if (term) {
      float r[2];
      R::returns(s, cfg, r);
      ret[0] = r[0]; ret[1] = r[1];
      pool[cur].out0 = r[0]; pool[cur].out1 = r[1]; pool[cur].has_outcome = 1;
      solved = P.solve != 0;
    } else {
      ret[0] = 0; ret[1] = 0;
      for (int ro = 0; ro < P.n_rollouts; ++ro) {
        typename R::S w = s;
        u32 ply = 0;
        while (!R::terminal(w, cfg) && (int)ply < P.max_plies) {
          auto draw = [&](u32 b, u32 n) { return rng_uniform(key, (u32)sim, b, 2u + (u32)ro, n); };
          playout_step<R>(w, cfg, workctx, tree, P.mask_words, draw, ply);
          ++ply;
        }
        float r[2];
        R::returns(w, cfg, r);
        ret[0] = __dadd_rn(ret[0], (double)r[0]);
        ret[1] = __dadd_rn(ret[1], (double)r[1]);
      }
      ret[0] = __ddiv_rn(ret[0], (double)P.n_rollouts);
      ret[1] = __ddiv_rn(ret[1], (double)P.n_rollouts);
      solved = false;
    }
    // ---- backup + solver (mcts.cc:384-434) ----
    while (depth > 0) {
      u32 ni = path[--depth];
      MctsNode nd = pool[ni];
      nd.total_reward = __dadd_rn(nd.total_reward, ret[nd.player]);
      nd.visits += 1;
      if (solved && nd.nchild > 0) {
        int player = pool[nd.first_child].player;
        int best = -1;
        float best_v = 0;
        bool all_solved = true;
        for (int i = 0; i < nd.nchild; ++i) {
          const MctsNode& ch = pool[nd.first_child + i];
          if (!ch.has_outcome) all_solved = false;
          else {
            float v = player == 0 ? ch.out0 : ch.out1;
            if (best < 0 || v > best_v) { best = i; best_v = v; }
          }
        }
        if (best >= 0 && (all_solved || (double)best_v == P.max_utility)) {
          nd.out0 = pool[nd.first_child + best].out0; nd.out1 = pool[nd.first_child + best].out1; nd.has_outcome = 1;
        } else {
          solved = false;
        }
      }
      pool[ni] = nd;
    }
    if (pool[root_idx].has_outcome || pool[root_idx].nchild == 1) { ++sim; break; }
  }
  if (failed) { flag_error(P.err, tree); }
  // ---- report the root's children + BestChild (mcts.cc:127-143) ----
  MctsNode r = pool[root_idx];
  int best = -1;
  for (int i = 0; i < r.nchild; ++i) {
    const MctsNode& ch = pool[r.first_child + i];
    long long o = tree * P.num_actions + ch.action;
    P.visits_out[o] = (int)ch.visits;
    P.reward_out[o] = ch.total_reward;
    if (P.outcome_out && ch.has_outcome) P.outcome_out[o] = ch.out0;
    if (best < 0) { best = i; continue; }
    const MctsNode& b = pool[r.first_child + best];       // CompareFinal(b, ch): is b "less than" ch?
    double ob = b.has_outcome ? (double)(b.player == 0 ? b.out0 : b.out1) : 0.0;
    double oc = ch.has_outcome ? (double)(ch.player == 0 ? ch.out0 : ch.out1) : 0.0;
    bool less = ob != oc ? ob < oc : (b.visits != ch.visits ? b.visits < ch.visits : b.total_reward < ch.total_reward);
    if (less) best = i;
  }
  if (best >= 0) P.best_out[tree] = pool[r.first_child + best].action;
  if (P.sims_out) P.sims_out[tree] = sim;
}

}  // namespace b2s
