// tsan_two_contexts_driver.cpp -- test infrastructure (profiles/tools/host_sanitize.sh, TSan leg; tests/test_abi.py):
// the C ABI's threading contract -- any number of contexts, each used by one host thread at a time -- on the HOST
// code of the library, over tests/stub_cudart.cpp (no GPU, kernels are no-ops).  Two threads, each with its own
// context and its own SLAM-shaped graph, run graph_load + graph_prepare (structure pass, symbolic phase on the
// shared host pool, upload packing) + graph_optimize + map / frame staging at the same time, under ThreadSanitizer;
// the bytes each thread uploads must equal what it uploads alone.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../include/slam_b200.h"

struct G {
  std::vector<int32_t> pid, lid, of, ot, ep, el, fixed;
  std::vector<double> pe, le, oz, oi, ez, ei;
};
static G make(int P, int L, int laps, unsigned seed) {
  G g;
  std::mt19937 rng(seed);
  for (int l = 0; l < L; l++) { g.lid.push_back(l); g.le.push_back(l); g.le.push_back(-l); }
  const int base = 1000 > L ? 1000 : L;
  for (int p = 0; p < P; p++) { g.pid.push_back(base + p); g.pe.push_back(0.3 * p); g.pe.push_back(0.1 * p); g.pe.push_back(0.01 * p); }
  const double I3[9] = {5, 0, 0, 0, 5, 0, 0, 0, 5}, I2[4] = {0.01, 0, 0, 0.01};
  for (int p = 0; p + 1 < P; p++) {
    g.of.push_back(base + p); g.ot.push_back(base + p + 1);
    g.oz.push_back(0.3); g.oz.push_back(0.0); g.oz.push_back(0.01);
    g.oi.insert(g.oi.end(), I3, I3 + 9);
  }
  for (int p = 0; p < P; p++) {
    const int c = (int)((long)(p % (P / laps)) * L / (P / laps)), k = 4 + (int)(rng() % 6);
    for (int q = 0; q < k; q++) {
      g.ep.push_back(base + p); g.el.push_back((c + q) % L);
      g.ez.push_back(1.0 + q); g.ez.push_back(0.5 - q);
      g.ei.insert(g.ei.end(), I2, I2 + 4);
    }
  }
  g.fixed = {base, base + 1, 0, 1};
  return g;
}
extern "C" unsigned long stub_h2d_hash();   // process-wide in the stub: only meaningful single-threaded
extern "C" void stub_reset();

static long run(const G& g, int rounds, double* chi_out) {
  slam_b200_ctx* c = nullptr;
  if (slam_b200_create(0, nullptr, &c)) return -1;
  long n = 0;
  for (int r = 0; r < rounds; r++) {
    if (slam_b200_graph_load(c, (int)g.pid.size(), g.pid.data(), g.pe.data(), (int)g.lid.size(), g.lid.data(), g.le.data(),
                             (int)g.of.size(), g.of.data(), g.ot.data(), g.oz.data(), g.oi.data(), (int)g.ep.size(), g.ep.data(),
                             g.el.data(), g.ez.data(), g.ei.data(), (int)g.fixed.size(), g.fixed.data())) return -2;
    n = slam_b200_graph_prepare(c);
    if (n < 0) return -3;
    double chi2[4];
    slam_b200_graph_optimize(c, 2, chi2);   // kernels are no-ops: only the host side of an iteration runs
    double st[16];
    slam_b200_graph_stats(c, st);
    *chi_out = st[5];                        // nnz(L): a fingerprint of the analysis
    std::vector<double> x(100, 1.0), y(100, 2.0);
    std::vector<int32_t> t(100, 1);
    slam_b200_map_clear(c);
    slam_b200_map_append(c, x.data(), y.data(), t.data(), 100);
  }
  slam_b200_destroy(c);
  return n;
}

int main() {
  const G a = make(6000, 300, 6, 1), b = make(900, 250, 1, 2);
  double fa = 0, fb = 0, ca = 0, cb = 0;
  const long na = run(a, 1, &fa), nb = run(b, 1, &fb);
  long ma = 0, mb = 0;
  std::thread t1([&] { ma = run(a, 4, &ca); });
  std::thread t2([&] { mb = run(b, 30, &cb); });
  t1.join();
  t2.join();
  if (na <= 0 || nb <= 0 || ma != na || mb != nb || ca != fa || cb != fb) {
    printf("mismatch: alone (%ld, %ld, %.0f, %.0f) together (%ld, %ld, %.0f, %.0f)\n", na, nb, fa, fb, ma, mb, ca, cb);
    return 1;
  }
  printf("ok %ld %ld %.0f %.0f\n", na, nb, fa, fb);
  return 0;
}
