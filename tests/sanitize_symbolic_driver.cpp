// sanitize_symbolic_driver.cpp -- test infrastructure (tests/test_symbolic.py::test_symbolic_phase_under_sanitizers):
// runs the host symbolic phase (csrc/symbolic.cpp, incl. its thread pool) over SLAM-shaped and adversarial block
// patterns -- 10-lap and 1-lap pose/landmark graphs, random sparse graphs, a near-clique, a single vertex, the empty
// graph, no edges, thousands of components -- built with -fsanitize=address,undefined or -fsanitize=thread.
// Prints "ok <hash>" (xor of the per-repetition hashes of the results: 0 for an even number of repetitions when
// the analysis is deterministic).
#include "symbolic.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <thread>
#include <utility>
// slam-like pattern: L landmarks (dim 2) then P poses (dim 3) in g2o order; chain + each pose sees k landmarks
static void slam_pattern(int P, int L, int laps, std::mt19937& rng, std::vector<int>& dim, std::vector<int>& a, std::vector<int>& b) {
  dim.assign(L, 2); dim.insert(dim.end(), P, 3);
  std::set<std::pair<int,int>> S;
  for (int p = 0; p + 1 < P; p++) S.insert({L + p, L + p + 1});
  for (int p = 0; p < P; p++) {
    int c = (int)((long)(p % (P / laps)) * L / (P / laps));
    int k = 4 + rng() % 6;
    for (int q = 0; q < k; q++) { int l = (c + q) % L; S.insert({l, L + p}); }
  }
  for (auto& pr : S) { a.push_back(pr.first); b.push_back(pr.second); }
}
static void random_pattern(int nb, int deg, std::mt19937& rng, std::vector<int>& dim, std::vector<int>& a, std::vector<int>& b) {
  dim.resize(nb); for (auto& d : dim) d = 2 + rng() % 2;
  std::set<std::pair<int,int>> S;
  for (long k = 0; k < (long)nb * deg / 2; k++) { int u = rng() % nb, v = rng() % nb; if (u == v) continue; S.insert({std::min(u, v), std::max(u, v)}); }
  for (auto& pr : S) { a.push_back(pr.first); b.push_back(pr.second); }
}
static unsigned long run(const std::vector<int>& dim, const std::vector<int>& a, const std::vector<int>& b, int leaf) {
  int nb = (int)dim.size();
  std::vector<int> hd(nb), ho(a.size()); int cur = 0;
  for (int k = 0; k < nb; k++) { hd[k] = cur; cur += dim[k] * dim[k]; }
  for (size_t k = 0; k < a.size(); k++) { ho[k] = cur; cur += dim[a[k]] * dim[b[k]]; }
  Symbolic S;
  symbolic_analyze(nb, dim.data(), (int)a.size(), a.data(), b.data(), hd.data(), ho.data(), leaf, S);
  unsigned long h = (unsigned long)S.nnzL * 31 + S.nf;
  for (int v : S.pos) h = h * 1000003 + v;
  for (int v : S.rel) h = h * 1000003 + v;
  return h;
}
int main(int argc, char** argv) {
  int reps = argc > 1 ? atoi(argv[1]) : 3;
  std::mt19937 rng(5);
  unsigned long H = 0;
  for (int r = 0; r < reps; r++) {
    { std::vector<int> d, a, b; std::mt19937 g(1); slam_pattern(10000, 300, 10, g, d, a, b); H ^= run(d, a, b, 1024); }
    { std::vector<int> d, a, b; std::mt19937 g(2); slam_pattern(1000, 300, 1, g, d, a, b); H ^= run(d, a, b, 130) * 3; }
    { std::vector<int> d, a, b; std::mt19937 g(3); random_pattern(6000, 4, g, d, a, b); H ^= run(d, a, b, 256) * 5; }
    { std::vector<int> d, a, b; std::mt19937 g(4); random_pattern(50, 49, g, d, a, b); H ^= run(d, a, b, 8) * 7; }      // near-clique
    { std::vector<int> d(1, 3), a, b; H ^= run(d, a, b, 8) * 11; }                                                       // single vertex
    { std::vector<int> d, a, b; H ^= run(d, a, b, 8) * 13; }                                                             // empty
    { std::vector<int> d(500, 2), a, b; H ^= run(d, a, b, 8) * 17; }                                                     // no edges
    { std::vector<int> d, a, b; std::mt19937 g(6); random_pattern(5000, 1, g, d, a, b); H ^= run(d, a, b, 64) * 19; }    // many components
  }
  // two host threads analysing different graphs at the same time (two contexts share the process-wide pool):
  // each must get what it gets alone
  {
    std::vector<int> d1, a1, b1, d2, a2, b2;
    { std::mt19937 g(1); slam_pattern(6000, 300, 6, g, d1, a1, b1); }
    { std::mt19937 g(3); random_pattern(6000, 4, g, d2, a2, b2); }
    const unsigned long s1 = run(d1, a1, b1, 600), s2 = run(d2, a2, b2, 256);
    unsigned long c1 = 0, c2 = 0;
    std::thread t1([&] { for (int k = 0; k < 3; k++) c1 = run(d1, a1, b1, 600); });
    std::thread t2([&] { for (int k = 0; k < 3; k++) c2 = run(d2, a2, b2, 256); });
    t1.join(); t2.join();
    if (c1 != s1 || c2 != s2) { printf("concurrent analysis differs from serial\n"); return 1; }
  }
  printf("ok %lx\n", H);
}
