// Runs the REFERENCE'S OWN test harness (open_spiel/tests/basic_tests.cc: RandomSimTest — legal actions sorted,
// clone equality, serialization round trip, tensor sizes, returns bounds, game length, ...) on the B200 adapter
// games obtained through open_spiel::LoadGame, then plays the stock C++ game and the adapter in lock-step.
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

static void LockStep(const Game& ours, const Game& stock, int games, std::mt19937* rng) {
  for (int g = 0; g < games; ++g) {
    auto a = ours.NewInitialState();
    auto b = stock.NewInitialState();
    while (true) {
      SPIEL_CHECK_EQ(a->IsTerminal(), b->IsTerminal());
      SPIEL_CHECK_EQ(a->CurrentPlayer(), b->CurrentPlayer());
      SPIEL_CHECK_TRUE(a->LegalActions() == b->LegalActions());
      SPIEL_CHECK_TRUE(a->Returns() == b->Returns());
      SPIEL_CHECK_EQ(a->ToString(), b->ToString());
      for (Player p = 0; p < 2; ++p) SPIEL_CHECK_TRUE(a->ObservationTensor(p) == b->ObservationTensor(p));
      if (a->IsTerminal()) break;
      auto la = b->LegalActions();
      Action act = la[(*rng)() % la.size()];
      SPIEL_CHECK_EQ(a->ActionToString(act), b->ActionToString(act));
      a->ApplyAction(act);
      b->ApplyAction(act);
    }
    SPIEL_CHECK_TRUE(a->History() == b->History());
  }
}

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

int main() {
  // stock game objects, built directly from their classes before the names are taken over
  std::shared_ptr<const Game> stock_c4 = LoadGame("connect_four");
  std::shared_ptr<const Game> stock_ttt = LoadGame("tic_tac_toe");
  std::shared_ptr<const Game> stock_c4_small = LoadGame("connect_four(rows=4,columns=5,x_in_row=3)");
  BotTests(*stock_c4, *stock_ttt);
  CfrTests();
  b200::RegisterB200Games();
  std::shared_ptr<const Game> c4 = LoadGame("connect_four");
  std::shared_ptr<const Game> ttt = LoadGame("tic_tac_toe");
  std::shared_ptr<const Game> c4_small = LoadGame("connect_four(rows=4,columns=5,x_in_row=3)");
  SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(c4.get()) != nullptr);     // LoadGame now returns the adapter
  SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(ttt.get()) != nullptr);
  SPIEL_CHECK_TRUE(dynamic_cast<const b200::B200Game*>(stock_c4.get()) == nullptr);
  SPIEL_CHECK_EQ(c4->NumDistinctActions(), 7);
  SPIEL_CHECK_EQ(c4->MaxGameLength(), 42);
  std::mt19937 rng(7);
  LockStep(*c4, *stock_c4, 40, &rng);
  LockStep(*ttt, *stock_ttt, 40, &rng);
  LockStep(*c4_small, *stock_c4_small, 20, &rng);
  testing::RandomSimTest(*c4, 15);        // the reference's own harness on the drop-in
  testing::RandomSimTest(*ttt, 15);
  testing::RandomSimTest(*c4_small, 10);
  std::cout << "adapter_test ok" << std::endl;
  return 0;
}
