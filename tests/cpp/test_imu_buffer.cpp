// ImuPropagator mirror (ingvio_b200/host/ingvio_filter.hpp) against the oracle's restatement of
// ImuPropagator::propagateUntil (ImuPropagator.cpp:232-292): replays a list of storeImu / propagateUntil events and writes,
// after every propagateUntil, the state time, the samples left in the buffer and the (gyro, accel, dt) steps that were
// handed to igv_propagate_imu. Runs against tests/emul/igv_shim.cpp only (it records those arguments); the comparison with
// the oracle is in tests/test_cpp_updaters.py::test_imu_buffer_matches_oracle.
//   usage: test_imu_buffer <events.bin> <out.bin>     events: [kind, ...]: 0 stamp g3 a3 = storeImu, 1 t_end = propagateUntil
#include <cstdio>

#include "../../ingvio_b200/host/ingvio_filter.hpp"

extern "C" int igv_shim_last_propagate(igv_batch* h, double* gyro, double* accel, double* dt, int cap);

using namespace ingvio;

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<double> ev;
  double v;
  while (std::fread(&v, sizeof(double), 1, f) == 1) ev.push_back(v);
  std::fclose(f);
  StateParams sp;
  sp._max_sw_poses = 3;
  auto state = std::make_shared<State>(sp, 8, 1);
  Mat3 I;
  state->initStateAndCov(ev[0], I, Vec3d(), Vec3d(), Vec3d(), Vec3d());     // first value: initial state time
  ImuPropagator propa;
  FILE* out = std::fopen(argv[2], "wb");
  const int cap = 4096;
  std::vector<double> g(3 * cap), a(3 * cap), d(cap);
  for (std::size_t i = 1; i < ev.size();) {
    if (ev[i] == 0.0) {
      Vec3d w, acc;
      for (int k = 0; k < 3; ++k) { w[k] = ev[i + 2 + k]; acc[k] = ev[i + 5 + k]; }
      propa.storeImu(ImuCtrl(ev[i + 1], w, acc));
      i += 8;
    } else {
      propa.propagateUntil(state, ev[i + 1]);
      const int n = igv_shim_last_propagate(StateManager::handle(state), g.data(), a.data(), d.data(), cap);
      const double hdr[3] = {state->_timestamp, (double)propa.bufferSize(), (double)n};
      std::fwrite(hdr, sizeof(double), 3, out);
      for (int s = 0; s < n; ++s) {
        std::fwrite(&g[3 * s], sizeof(double), 3, out);
        std::fwrite(&a[3 * s], sizeof(double), 3, out);
        std::fwrite(&d[s], sizeof(double), 1, out);
      }
      i += 2;
    }
  }
  std::fclose(out);
  std::printf("EVENTS DONE\n");
  return 0;
}
