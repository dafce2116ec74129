// GPU half of the adapter tests (the CPU half — the reference's RandomSimTest harness and lock-step against the stock
// games — is adapter_host_test.cc): the device MCTS / CFR behind the reference's Bot / solver interfaces, driven by the
// reference's own EvaluateBots / Exploitability, and the packed-lane bridge between scalar states and device batches.
#include <cstring>
#include <iostream>
#include <random>

#include "b200_algorithms.h"
#include "b200_games.h"
#include "open_spiel/algorithms/cfr.h"
#include "open_spiel/algorithms/evaluate_bots.h"
#include "open_spiel/algorithms/mcts.h"
#include "open_spiel/algorithms/tabular_exploitability.h"
#include "open_spiel/games/connect_four/connect_four.h"
#include "open_spiel/games/tic_tac_toe/tic_tac_toe.h"
#include "open_spiel/spiel.h"
#include "open_spiel/tests/basic_tests.h"

using namespace open_spiel;

// The reference's own self-play driver (algorithms/evaluate_bots.cc:28-66) with the device MCTS plugged in as a Bot.
static void BotTests(const Game& stock_c4, const Game& stock_ttt) {
  // connect_four: 400 simulations per move against a uniform random bot, both seats
  int wins = 0, games = 0;
  for (int seat = 0; seat < 2; ++seat) {
    for (int g = 0; g < 8; ++g) {
      b200::B200MCTSBot mcts(stock_c4, /*n_rollouts=*/1, /*uct_c=*/2.0, /*max_simulations=*/400, /*max_memory_mb=*/100,
                             /*solve=*/true, /*seed=*/100 + g, /*verbose=*/false);
      auto rnd = MakeUniformRandomBot(1 - seat, 7 + g);
      std::vector<Bot*> bots(2);
      bots[seat] = &mcts;
      bots[1 - seat] = rnd.get();
      std::vector<double> r = EvaluateBots(stock_c4, bots, 11 + g);
      wins += r[seat] > 0;
      ++games;
    }
  }
  std::cout << "B200MCTSBot vs uniform random on connect_four: " << wins << "/" << games << " wins" << std::endl;
  SPIEL_CHECK_GE(wins, games - 1);
  // tic_tac_toe: against the reference's own MCTSBot with the same budget, perfect play from both sides is a draw
  for (int seat = 0; seat < 2; ++seat) {
    b200::B200MCTSBot ours(stock_ttt, 20, 2.0, 2000, 100, true, 5, false);
    auto evaluator = std::make_shared<algorithms::RandomRolloutEvaluator>(20, 42);
    algorithms::MCTSBot theirs(stock_ttt, evaluator, 2.0, 2000, 100, true, 42, false);
    std::vector<Bot*> bots(2);
    bots[seat] = &ours;
    bots[1 - seat] = &theirs;
    std::vector<double> r = EvaluateBots(stock_ttt, bots, 3);
    SPIEL_CHECK_EQ(r[0], 0.0);
    SPIEL_CHECK_EQ(r[1], 0.0);
  }
  // PUCT selection through the same interface
  b200::B200MCTSBot puct(stock_c4, 1, 2.0, 300, 100, true, 9, false, algorithms::ChildSelectionPolicy::PUCT);
  auto st = stock_c4.NewInitialState();
  Action a = puct.Step(*st);
  SPIEL_CHECK_TRUE(a >= 0 && a < 7);
  int total = 0;
  for (int v : puct.LastVisitCounts()) total += v;
  SPIEL_CHECK_EQ(total, 299);                    // every simulation after the first descends into one root child
}

// B200CFRSolver hands the reference a TabularPolicy keyed by information-state strings; the reference's own
// Exploitability (tabular_exploitability.cc) must give exactly what it gives for its own CFRSolver's average policy.
static void CfrTests() {
  for (const char* name : {"kuhn_poker", "leduc_poker"}) {
    std::shared_ptr<const Game> game = LoadGame(name);
    const int iters = std::string(name) == "kuhn_poker" ? 200 : 20;
    b200::B200CFRSolver ours(*game);
    algorithms::CFRSolver theirs(*game);
    ours.EvaluateAndUpdatePolicy(iters);
    for (int i = 0; i < iters; ++i) theirs.EvaluateAndUpdatePolicy();
    TabularPolicy avg = ours.AveragePolicy();
    double e_ours = algorithms::Exploitability(*game, avg);
    double e_theirs = algorithms::Exploitability(*game, *theirs.AveragePolicy());
    std::cout << name << ": exploitability after " << iters << " iterations: device tables " << e_ours << ", reference "
              << e_theirs << ", device NashConv/2 " << ours.NashConv() / 2 << std::endl;
    SPIEL_CHECK_EQ(e_ours, e_theirs);
    SPIEL_CHECK_TRUE(std::abs(ours.NashConv() / 2 - e_theirs) < 1e-9);
  }
}

// A scalar B200State and a device lane are the same packed bytes: host state -> lane -> device kernels must agree with
// the host rule core on every observable, and one device ApplyAction must produce the blob the host produces.
static void LaneBridgeTests(std::mt19937* rng) {
  for (const char* name : {"tic_tac_toe", "connect_four", "breakthrough", "hex(board_size=7,swap=true)", "go(board_size=9)",
                           "go(board_size=5)", "kuhn_poker", "kuhn_poker(players=4)", "leduc_poker", "leduc_poker(players=3)", "mnk", "mnk(m=6,n=5,k=4)", "othello", "y(board_size=9)", "havannah(board_size=4,swap=true)", "havannah"}) {
    std::shared_ptr<const Game> game = LoadGame(name);
    const auto* bg = dynamic_cast<const b200::B200Game*>(game.get());
    SPIEL_CHECK_TRUE(bg != nullptr);
    const b2s_game_info& gi = bg->info();
    const int kLanes = 64;
    void* batch = bg->NewBatch(kLanes);
    std::vector<std::unique_ptr<State>> states;
    for (int i = 0; i < kLanes; ++i) {               // random positions of random depth
      auto s = game->NewInitialState();
      int plies = (int)((*rng)() % (unsigned)(gi.max_game_length + 1));
      for (int k = 0; k < plies && !s->IsTerminal(); ++k) {
        auto la = s->LegalActions();
        s->ApplyAction(la[(*rng)() % la.size()]);
      }
      static_cast<const b200::B200State&>(*s).ToBatchLane(batch, i);
      states.push_back(std::move(s));
    }
    void* dev = nullptr;
    const size_t W = gi.mask_words, P = gi.num_players, F = gi.observation_tensor_size;
    size_t bytes = kLanes * (4 * W + 1 + 1 + 4 * P + 4 * F + 4) + 256;
    SPIEL_CHECK_EQ(b2s_device_alloc(0, &dev, bytes), 0);
    char* d = (char*)dev;
    uint32_t* mask_d = (uint32_t*)d; d += kLanes * 4 * W;
    float* rets_d = (float*)d; d += kLanes * 4 * P;
    float* obs_d = (float*)d; d += kLanes * 4 * F;
    int32_t* act_d = (int32_t*)d; d += kLanes * 4;
    int8_t* cur_d = (int8_t*)d; d += kLanes;
    uint8_t* term_d = (uint8_t*)d;
    SPIEL_CHECK_EQ(b2s_legal_mask(batch, mask_d, kLanes, nullptr), 0);
    SPIEL_CHECK_EQ(b2s_status(batch, cur_d, term_d, rets_d, kLanes, nullptr), 0);
    SPIEL_CHECK_EQ(b2s_observation(batch, 0, obs_d, kLanes, nullptr), 0);
    std::vector<uint32_t> mask(kLanes * W);
    std::vector<float> rets(kLanes * P), obs(kLanes * F);
    std::vector<int8_t> cur(kLanes);
    std::vector<uint8_t> term(kLanes);
    b2s_memcpy_d2h(0, mask.data(), mask_d, mask.size() * 4, nullptr);
    b2s_memcpy_d2h(0, rets.data(), rets_d, rets.size() * 4, nullptr);
    b2s_memcpy_d2h(0, obs.data(), obs_d, obs.size() * 4, nullptr);
    b2s_memcpy_d2h(0, cur.data(), cur_d, kLanes, nullptr);
    b2s_memcpy_d2h(0, term.data(), term_d, kLanes, nullptr);
    SPIEL_CHECK_EQ(b2s_stream_synchronize(0, nullptr), 0);
    std::vector<int32_t> acts(kLanes, -1);
    for (int i = 0; i < kLanes; ++i) {
      const State& s = *states[i];
      SPIEL_CHECK_EQ((int)cur[i], (int)s.CurrentPlayer());
      SPIEL_CHECK_EQ((bool)term[i], s.IsTerminal());
      std::vector<Action> legal;
      for (size_t w = 0; w < W; ++w)
        for (int b = 0; b < 32; ++b) if ((mask[i * W + w] >> b) & 1u) legal.push_back(w * 32 + b);
      SPIEL_CHECK_TRUE(legal == s.LegalActions());
      std::vector<double> r = s.Returns();
      for (size_t p = 0; p < P; ++p) SPIEL_CHECK_EQ((double)rets[i * P + p], r[p]);
      std::vector<float> o = s.ObservationTensor(0);
      for (size_t f = 0; f < F; ++f) SPIEL_CHECK_EQ(obs[i * F + f], o[f]);
      if (!legal.empty()) acts[i] = (int32_t)legal[(*rng)() % legal.size()];
    }
    b2s_memcpy_h2d(0, act_d, acts.data(), kLanes * 4, nullptr);
    SPIEL_CHECK_EQ(b2s_apply_actions(batch, act_d, kLanes, nullptr), 0);
    int64_t bad = -1;
    SPIEL_CHECK_EQ(b2s_error_count(batch, &bad, nullptr, nullptr), 0);
    SPIEL_CHECK_EQ(bad, 0);
    for (int i = 0; i < kLanes; ++i) {
      if (acts[i] < 0) continue;
      states[i]->ApplyAction(acts[i]);                           // host rule core
      auto probe = game->NewInitialState();
      auto& lane = static_cast<b200::B200State&>(*probe);
      lane.FromBatchLane(batch, i);                              // device kernel
      const auto& host = static_cast<const b200::B200State&>(*states[i]);
      SPIEL_CHECK_EQ(memcmp(lane.blob(), host.blob(), bg->rules().state_bytes()), 0);
      SPIEL_CHECK_EQ(lane.ToString().substr(lane.ToString().find('\n') + 1),
                     host.ToString().substr(host.ToString().find('\n') + 1));   // (go prints history_.size() on line 1)
    }
    b2s_device_free(0, dev);
    b2s_batch_destroy(batch);
    std::cout << "lane bridge ok " << name << std::endl;
  }
}

int main() {
  // stock game objects, built by the stock factories before the names are taken over
  std::shared_ptr<const Game> stock_c4 = LoadGame("connect_four");
  std::shared_ptr<const Game> stock_ttt = LoadGame("tic_tac_toe");
  BotTests(*stock_c4, *stock_ttt);
  CfrTests();
  b200::RegisterB200Games();
  std::shared_ptr<const Game> c4 = LoadGame("connect_four");
  SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(c4.get()) != nullptr);     // LoadGame now returns the adapter
  std::mt19937 rng(7);
  LaneBridgeTests(&rng);
  // the algorithm adapters on the drop-in games as well (roots copied to the device as packed lanes)
  BotTests(*c4, *LoadGame("tic_tac_toe"));
  CfrTests();
  {
    // MCTSearch returns the reference's SearchNode: children = legal actions, visits sum to simulations - 1
    std::shared_ptr<const Game> go = LoadGame("go(board_size=9)");
    b200::B200MCTSBot bot(*go, 1, 2.0, 500, 100, true, 3, false);
    auto st = go->NewInitialState();
    st->ApplyAction(40);
    auto root = bot.MCTSearch(*st);
    SPIEL_CHECK_EQ(root->children.size(), st->LegalActions().size());
    int total = 0;
    for (const auto& ch : root->children) total += ch.explore_count;
    SPIEL_CHECK_EQ(total, 499);
    SPIEL_CHECK_EQ(root->BestChild().action, bot.Step(*st) >= 0 ? root->BestChild().action : -1);
  }
  std::cout << "adapter_test ok" << std::endl;
  return 0;
}
