// C++ tests of the map-server host mirror (ingvio_b200/host/ingvio_map_server.hpp) on top of the C-ABI track table.
// Restates the reference's gtest TEST_F(TestMapServer, collectFeatureAndMarg)
// (ingvio_estimator/test/TestMapServer.cpp:184-308) line by line against the mirrored classes, then follows a sliding
// window for a few frames (observation timestamps and anchors must survive marginalised clones).
// Built and run by tests/test_cpp_host_mirror.py: against libingvio_b200.so on the GPU, and -- to check the mirror's own
// logic without a GPU -- against tests/emul/igv_shim.cpp (the same C symbols backed by the kernel source run on the CPU).
#include <cmath>
#include <cstdio>
#include <functional>

#include "../../ingvio_b200/host/ingvio_map_server.hpp"

using namespace ingvio;

static unsigned long long g_seed = 2463534242ull;
static double urand() {
  g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17;
  return 2.0 * ((g_seed >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}
static int g_fail = 0;
#define CHECK(cond, ...)                                                   \
  do {                                                                     \
    if (!(cond)) { ++g_fail; std::printf("  FAILED %s:%d: %s | ", __FILE__, __LINE__, #cond); std::printf(__VA_ARGS__); std::printf("\n"); } \
  } while (0)
static void run(const char* name, const std::function<void()>& f) {
  const int before = g_fail;
  f();
  std::printf("[%s] %s\n", g_fail == before ? "  OK  " : "FAILED", name);
}

// generateRandomMonoFrame / generateRandomStereoFrame (TestMapServer.cpp:79-99, :132-154)
static feature_tracker::MonoFrame generateRandomMonoFrame(const std::vector<int>& ids, double timestamp) {
  feature_tracker::MonoFrame f;
  f.header.stamp = timestamp;
  for (int id : ids) { feature_tracker::MonoMeas m; m.id = id; m.u0 = urand(); m.v0 = urand(); f.mono_features.push_back(m); }
  return f;
}
static feature_tracker::StereoFrame generateRandomStereoFrame(const std::vector<int>& ids, double timestamp) {
  feature_tracker::StereoFrame f;
  f.header.stamp = timestamp;
  for (int id : ids) {
    feature_tracker::StereoMeas m; m.id = id; m.u0 = urand(); m.v0 = urand(); m.u1 = urand(); m.v1 = urand();
    f.stereo_features.push_back(m);
  }
  return f;
}
// imu_propa->propagateToExpectedPoseAndAugment(state, t, T) as far as the map server sees it: the state reaches time t
// with some pose and a clone is appended (ImuPropagator.cpp:316-341 -> StateManager::augmentSlidingWindowPose)
static void propagateToExpectedPoseAndAugment(std::shared_ptr<State> state, double t) {
  Vec3d p; for (int i = 0; i < 3; ++i) p[i] = urand();
  state->_extended_pose->setValueTrans1(p);
  state->_timestamp = t;
  StateManager::augmentSlidingWindowPose(state);
}
static std::shared_ptr<State> make_state(int cam_nums, int max_sw) {
  StateParams p;
  p._cam_nums = cam_nums;
  p._max_sw_poses = max_sw;
  auto state = std::make_shared<State>(p, /*max_feats*/ 32, /*max_sats*/ 4);
  Mat3 I; for (int i = 0; i < 9; ++i) I.m[i] = (i % 4 == 0);
  state->initStateAndCov(0.0, I, Vec3d(), Vec3d(), Vec3d(), Vec3d());
  return state;
}

static void collectFeatureAndMarg(bool stereo) {
  std::vector<int> id1(4), id2(4);
  for (int i = 0; i < 4; ++i) { id1[i] = i + 1; id2[i] = i + 2; }                        // :186-192
  auto state = make_state(stereo ? 2 : 1, 20);
  auto map_server = std::make_shared<MapServer>(64);
  auto collect = [&](const std::vector<int>& ids, double t) {
    if (stereo) MapServerManager::collectStereoMeas(map_server, state, generateRandomStereoFrame(ids, t));
    else MapServerManager::collectMonoMeas(map_server, state, generateRandomMonoFrame(ids, t));
  };
  auto frames = [&](int i) { return stereo ? map_server->at(i)->numOfStereoFrames() : map_server->at(i)->numOfMonoFrames(); };
  auto has = [&](int i, double t) { return stereo ? map_server->at(i)->hasStereoObsAt(t) : map_server->at(i)->hasMonoObsAt(t); };

  propagateToExpectedPoseAndAugment(state, 2.0);                                         // :201
  collect(id1, 2.0);                                                                     // :203
  CHECK(map_server->size() == 4, "size %zu", map_server->size());                        // :205
  for (int i = 1; i < 5; ++i) {                                                          // :207-213
    CHECK(map_server->at(i)->getId() == i, "id");
    CHECK(map_server->at(i)->getFeatureType() == FeatureInfo::MSCKF, "type");
    CHECK(frames(i) == 1, "frames %d", frames(i));
    CHECK(has(i, 2.0), "obs at 2.0");
  }
  propagateToExpectedPoseAndAugment(state, 4.0);                                         // :215
  collect(id2, 4.0);                                                                     // :217
  CHECK(map_server->size() == 5, "size %zu", map_server->size());                        // :219
  for (int i = 1; i < 6; ++i) {                                                          // :221-234
    CHECK(map_server->at(i)->getId() == i, "id");
    CHECK(map_server->at(i)->getFeatureType() == FeatureInfo::MSCKF, "type");
    CHECK(frames(i) == ((i == 1 || i == 5) ? 1 : 2), "frames of %d: %d", i, frames(i));
    if (i > 1) CHECK(has(i, 4.0), "obs at 4.0 of %d", i);
    CHECK(!map_server->at(i)->isToMarg(), "toMarg of %d", i);
  }
  if (stereo) MapServerManager::markMargStereoFeatures(map_server, state);              // :236 / :286
  else MapServerManager::markMargMonoFeatures(map_server, state);
  CHECK(map_server->size() == 5, "size %zu", map_server->size());                        // :238
  for (int i = 1; i < 6; ++i) CHECK(map_server->at(i)->isToMarg() == (i == 1), "toMarg of %d", i);   // :240-246
  for (int i = 1; i < 6; ++i)                                                            // :300-306
    CHECK(map_server->at(i)->anchor() == state->_sw_camleft_poses.at(i < 5 ? 2.0 : 4.0), "anchor of %d", i);
}

int main() {
  run("testMapServer.collectFeatureAndMarg, mono (TestMapServer.cpp:184-246, :300-306)", [] { collectFeatureAndMarg(false); });
  run("testMapServer.collectFeatureAndMarg, stereo (TestMapServer.cpp:248-306)", [] { collectFeatureAndMarg(true); });

  run("observation timestamps and anchors follow a sliding window (margSlidingWindowPose)", [] {
    auto state = make_state(1, 3);
    auto map_server = std::make_shared<MapServer>(64);
    std::map<int, std::map<double, std::array<double, 2>>> truth;    // id -> t -> uv
    for (int k = 1; k <= 7; ++k) {
      const double t = 0.5 * k;
      propagateToExpectedPoseAndAugment(state, t);
      std::vector<int> ids;
      for (int id = k; id < k + 4; ++id) ids.push_back(id);           // each id lives for four frames
      auto fr = generateRandomMonoFrame(ids, t);
      fr.mono_features.push_back(fr.mono_features[1]);                // a repeated id: the first measurement wins
      fr.mono_features.back().u0 = 9.0;
      MapServerManager::collectMonoMeas(map_server, state, fr);
      for (int i = 0; i < 4; ++i) truth[ids[i]][t] = {fr.mono_features[i].u0, fr.mono_features[i].v0};
      if (state->nextMargTime() != std::numeric_limits<double>::infinity()) {
        const double mt = state->nextMargTime();
        StateManager::margSlidingWindowPose(state);                   // the handle drops the observations at that clone
        for (auto& kv : truth) kv.second.erase(mt);
      }
      for (auto& kv : truth) {
        if (kv.second.empty()) continue;
        CHECK(map_server->count(kv.first), "id %d missing at frame %d", kv.first, k);
        if (!map_server->count(kv.first)) continue;
        auto f = map_server->at(kv.first);
        CHECK(f->numOfMonoFrames() == (int)kv.second.size(), "id %d: %d frames, expected %zu", kv.first, f->numOfMonoFrames(), kv.second.size());
        for (auto& ob : kv.second) {
          CHECK(f->hasMonoObsAt(ob.first), "id %d lacks t=%.1f", kv.first, ob.first);
          CHECK(f->monoMeasAt(ob.first) == ob.second, "id %d uv at t=%.1f", kv.first, ob.first);
        }
      }
      CHECK((int)state->_sw_camleft_poses.size() <= 3, "window");
    }
    MapServerManager::markMargMonoFeatures(map_server, state);
    for (auto it = map_server->begin(); it != map_server->end(); ++it)
      CHECK(it->second->isToMarg() == !it->second->hasMonoObsAt(state->_timestamp), "toMarg of %d", it->first);
  });

  if (g_fail == 0) std::printf("ALL TESTS PASSED\n");
  else std::printf("%d CHECK(S) FAILED\n", g_fail);
  return g_fail == 0 ? 0 : 1;
}
