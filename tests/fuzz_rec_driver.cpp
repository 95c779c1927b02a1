// fuzz_rec_driver.cpp -- test infrastructure (tests/test_rec_reader.py::test_mutated_streams_under_sanitizers):
// mutates the golden recording (byte flips, random streams, insertions, hostile 24-bit lengths / varints) and runs
// reader + replay over an exact-size heap copy of each mutant, built with -fsanitize=address,undefined.
//   fuzz_rec_driver <golden.rec> <seed> <mutants>
#include "rec_reader.hpp"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <cstring>
int main(int argc, char** argv) {
  slamrec::Reader src; if (!src.open(argv[1])) return 2;
  FILE* f = fopen(argv[1], "rb"); std::vector<uint8_t> base(70000); size_t n = fread(base.data(), 1, base.size(), f); base.resize(n); fclose(f);
  std::mt19937 rng(atoi(argv[2])); long frames = 0, env = 0;
  for (int it = 0; it < atoi(argv[3]); it++) {
    size_t len = 200 + rng() % (base.size() - 200);
    std::vector<uint8_t> d(base.begin(), base.begin() + len);
    int mode = rng() % 4;
    if (mode == 0) { int k = 1 + rng() % 40; while (k--) d[rng() % d.size()] = (uint8_t)rng(); }
    else if (mode == 1) { d.resize(rng() % 4000); for (auto& b : d) b = (uint8_t)rng(); }
    else if (mode == 2) { size_t p = rng() % d.size(); std::vector<uint8_t> ins(1 + rng() % 50); for (auto& b : ins) b = (uint8_t)rng(); d.insert(d.begin() + p, ins.begin(), ins.end()); }
    else { size_t p = rng() % (d.size() - 16); uint8_t h[12] = {0x0d, 0xa4, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff}; for (int k = 0; k < 12; k++) d[p + k] = h[k]; }
    // exact-size heap copy so that any over-read trips the sanitizer
    uint8_t* heap = (uint8_t*)malloc(d.size() ? d.size() : 1); memcpy(heap, d.data(), d.size());
    slamrec::Reader r; r.attach(heap, d.size());
    slamrec::ReplayConfig cfg; cfg.gatheringTimeMs = 10;
    slamrec::ReplayStats s = slamrec::replay(r, cfg, [&](const slamrec::ReplayFrame& fr) { frames += fr.cones.cols(); });
    env += s.envelopes;
    free(heap);
  }
  printf("ok %ld envelopes %ld cols\n", env, frames);
}
