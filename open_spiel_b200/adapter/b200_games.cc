#include "b200_games.h"

#include <cstring>

namespace open_spiel {
namespace b200 {
namespace {

void Check(int rc) {
  if (rc != 0) SpielFatalError(std::string("b2s: ") + b2s_last_error());   // the reference's own error path
}

b2s_params ToCParams(const std::string& name, const GameParameters& params) {
  b2s_params p;
  b2s_params_default(&p);
  auto geti = [&](const char* k, int32_t* out) {
    auto it = params.find(k);
    if (it == params.end()) return;
    if (it->second.has_int_value()) *out = it->second.int_value();
    else if (it->second.has_bool_value()) *out = it->second.bool_value() ? 1 : 0;
  };
  if (name == "connect_four") {
    geti("rows", &p.rows); geti("columns", &p.columns); geti("x_in_row", &p.x_in_row);
    geti("egocentric_obs_tensor", &p.egocentric_obs_tensor);
  }
  return p;
}

}  // namespace

B200Game::B200Game(const GameType& type, const GameParameters& params) : Game(type, params) {
  gid_ = b2s_game_id(type.short_name.c_str());
  if (gid_ < 0) SpielFatalError("b200: unsupported game " + type.short_name);
  cparams_ = ToCParams(type.short_name, params);
  Check(b2s_game_info_get(gid_, &cparams_, &info_));
}

std::unique_ptr<State> B200Game::NewInitialState() const {
  return std::unique_ptr<State>(new B200State(shared_from_this()));
}

std::vector<int> B200Game::ObservationTensorShape() const {
  std::vector<int> s;
  for (int d : info_.obs_shape) if (d > 0) s.push_back(d);
  return s;
}

void* B200Game::NewBatch(int64_t n, int device) const {
  void* b = nullptr;
  Check(b2s_batch_create(gid_, &cparams_, n, device, &b));
  return b;
}

static size_t ScratchBytes(const b2s_game_info& gi) {
  size_t obs = sizeof(float) * (size_t)gi.observation_tensor_size;
  return obs > 64 ? obs : 64;
}

B200State::B200State(std::shared_ptr<const Game> game) : State(game) {
  batch_ = bgame().NewBatch(1);
  Check(b2s_device_alloc(0, &scratch_d_, ScratchBytes(bgame().info())));
}

B200State::B200State(const B200State& other) : State(other) {
  batch_ = bgame().NewBatch(1);
  Check(b2s_device_alloc(0, &scratch_d_, ScratchBytes(bgame().info())));
  Check(b2s_copy_states(batch_, 0, other.batch_, 0, 1, nullptr));
}

B200State::~B200State() {
  if (batch_) b2s_batch_destroy(batch_);
  if (scratch_d_) b2s_device_free(0, scratch_d_);
}

Player B200State::CurrentPlayer() const {
  int8_t cur;
  Check(b2s_status(batch_, (int8_t*)scratch_d_, nullptr, nullptr, 1, nullptr));
  Check(b2s_memcpy_d2h(0, &cur, scratch_d_, 1, nullptr));
  Check(b2s_stream_synchronize(0, nullptr));
  return cur;
}

bool B200State::IsTerminal() const {
  uint8_t t;
  Check(b2s_status(batch_, nullptr, (uint8_t*)scratch_d_, nullptr, 1, nullptr));
  Check(b2s_memcpy_d2h(0, &t, scratch_d_, 1, nullptr));
  Check(b2s_stream_synchronize(0, nullptr));
  return t != 0;
}

std::vector<double> B200State::Returns() const {
  float r[2];
  Check(b2s_status(batch_, nullptr, nullptr, (float*)scratch_d_, 1, nullptr));
  Check(b2s_memcpy_d2h(0, r, scratch_d_, sizeof r, nullptr));
  Check(b2s_stream_synchronize(0, nullptr));
  return {(double)r[0], (double)r[1]};
}

std::vector<Action> B200State::LegalActions() const {
  const int words = bgame().info().mask_words;
  std::vector<uint32_t> m(words);
  Check(b2s_legal_mask(batch_, (uint32_t*)scratch_d_, 1, nullptr));
  Check(b2s_memcpy_d2h(0, m.data(), scratch_d_, sizeof(uint32_t) * words, nullptr));
  Check(b2s_stream_synchronize(0, nullptr));
  std::vector<Action> out;                         // ascending ids, empty at terminal states (spiel.h:374-388)
  for (int w = 0; w < words; ++w)
    for (int b = 0; b < 32; ++b) if ((m[w] >> b) & 1u) out.push_back(w * 32 + b);
  return out;
}

void B200State::DoApplyAction(Action action_id) {
  int32_t a = (int32_t)action_id;
  Check(b2s_memcpy_h2d(0, scratch_d_, &a, sizeof a, nullptr));
  Check(b2s_apply_actions(batch_, (const int32_t*)scratch_d_, 1, nullptr));
  int64_t bad = 0;
  Check(b2s_error_count(batch_, &bad, nullptr, nullptr));
  if (bad) SpielFatalError("b200: illegal action " + std::to_string(action_id));   // connect_four.cc:131-133's CHECK
}

void B200State::ObservationTensor(Player player, absl::Span<float> values) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, num_players_);
  SPIEL_CHECK_EQ((int)values.size(), bgame().info().observation_tensor_size);
  Check(b2s_observation(batch_, player, (float*)scratch_d_, 1, nullptr));
  Check(b2s_memcpy_d2h(0, values.data(), scratch_d_, sizeof(float) * values.size(), nullptr));
  Check(b2s_stream_synchronize(0, nullptr));
}

std::unique_ptr<State> B200State::Clone() const { return std::unique_ptr<State>(new B200State(*this)); }

// Strings are host-side decoding of the packed lane (b2s_state_get); formats follow connect_four.cc:158-161, 212-222
// and tic_tac_toe.cc:150-176 / Game::ActionToString.
std::string B200State::ToString() const {
  const auto& gi = bgame().info();
  std::string s;
  if (gi.game_id == B2S_CONNECT_FOUR) {
    uint64_t w[2];
    Check(b2s_state_get(batch_, 0, w, sizeof w));
    const int rows = gi.obs_shape[1], cols = gi.obs_shape[2], h1 = rows + 1;
    for (int r = rows - 1; r >= 0; --r) {
      for (int c = 0; c < cols; ++c) {
        int bit = c * h1 + r;
        s += ((w[0] >> bit) & 1) ? "x" : ((w[1] >> bit) & 1) ? "o" : ".";
      }
      s += "\n";
    }
  } else if (gi.game_id == B2S_TIC_TAC_TOE) {
    uint32_t b;
    Check(b2s_state_get(batch_, 0, &b, sizeof b));
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        int cell = r * 3 + c;
        s += ((b >> cell) & 1) ? "x" : ((b >> (9 + cell)) & 1) ? "o" : ".";
      }
      if (r < 2) s += "\n";
    }
  } else {
    s = "<b200 state: " + HistoryString() + ">";
  }
  return s;
}

std::string B200State::ActionToString(Player player, Action action_id) const {
  const auto& gi = bgame().info();
  const char* mark = player == 0 ? "x" : "o";
  if (gi.game_id == B2S_CONNECT_FOUR) return std::string(mark) + std::to_string(action_id);
  if (gi.game_id == B2S_TIC_TAC_TOE)
    return std::string(mark) + "(" + std::to_string(action_id / 3) + "," + std::to_string(action_id % 3) + ")";
  return std::to_string(action_id);
}

void RegisterB200Games() {
  for (const char* name : {"tic_tac_toe", "connect_four"}) {
    GameType type = LoadGame(name)->GetType();          // the stock registration's GameType, unchanged
    type.provides_observation_string = true;
    GameRegisterer::RegisterGame(type, [type](const GameParameters& params) {
      return std::shared_ptr<const Game>(new B200Game(type, params));
    });
  }
}

}  // namespace b200
}  // namespace open_spiel
