// DEVELOPMENT PROBE (not product, never loaded by the package): host build of csrc/minsnap_core.cuh so
// the K1 arithmetic can be compared with tests/golden/planning.npz in the GPU-less build container.
//   g++ -O2 -shared -fPIC -I uav-autonomous-control_b200/csrc tests/devtools/host_probe_minsnap.cpp -o /tmp/libprobe_minsnap.so
#include "minsnap_core.cuh"
using namespace uavb;
template <int MAXS> static int run(const double* w, double vel, int S, double factor, double* c, double* t) {
  return minsnap_solve_one<MAXS>(S, vel, factor, [w](int i, int ax) { return w[3 * i + ax]; },
                                 [c](int seg, int j, int ax, double v) { c[seg * 24 + j * 3 + ax] = v; },
                                 [t](int seg, double T) { t[seg] = T; });
}
extern "C" int probe_solve(const double* w, double vel, int S, double factor, double* c, double* t, int big) {
  if (big) return run<64>(w, vel, S, factor, c, t);
  if (S == 1) return run<1>(w, vel, S, factor, c, t);
  if (S == 2) return run<2>(w, vel, S, factor, c, t);
  if (S <= 4) return run<4>(w, vel, S, factor, c, t);
  return run<8>(w, vel, S, factor, c, t);
}
