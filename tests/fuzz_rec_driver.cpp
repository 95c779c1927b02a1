// fuzz_rec_driver.cpp -- test infrastructure (tests/test_rec_reader.py::test_mutated_streams_under_sanitizers):
// mutates the golden recording (byte flips, random streams, insertions, hostile 24-bit lengths / varints, VALID 10-byte
// varints carrying lengths near 2^64 in front of length-delimited fields -- the wrap-around of `off + len`) and runs
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
  auto run = [&](const std::vector<uint8_t>& d) {
    // exact-size heap copy so that any over-read trips the sanitizer
    uint8_t* heap = (uint8_t*)malloc(d.size() ? d.size() : 1); memcpy(heap, d.data(), d.size());
    slamrec::Reader r; r.attach(heap, d.size());
    slamrec::ReplayConfig cfg; cfg.gatheringTimeMs = 10;
    slamrec::ReplayStats s = slamrec::replay(r, cfg, [&](const slamrec::ReplayFrame& fr) { frames += fr.cones.cols(); });
    env += s.envelopes;
    free(heap);
  };
  // the advisor's 16-byte vector: envelope of 11 bytes whose field 3 carries the length 2^64 - 11 as a valid varint
  run({0x0D, 0xA4, 0x0B, 0x00, 0x00, 0x1A, 0xF5, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0x01});
  // a valid 10-byte varint for a 64-bit value: nine continuation bytes + a last byte of 0 or 1
  auto huge_varint = [&](uint8_t* o) {
    uint64_t v = ~(uint64_t)0 - (rng() % 70000);
    if (rng() % 4 == 0) v = ((uint64_t)rng() << 32 | rng()) | ((uint64_t)1 << 63);
    for (int k = 0; k < 9; k++) { o[k] = (uint8_t)(v & 0x7f) | 0x80; v >>= 7; }
    o[9] = (uint8_t)(v & 1);
  };
  for (int it = 0; it < atoi(argv[3]); it++) {
    size_t len = 200 + rng() % (base.size() - 200);
    std::vector<uint8_t> d(base.begin(), base.begin() + len);
    int mode = rng() % 6;
    if (mode == 0) { int k = 1 + rng() % 40; while (k--) d[rng() % d.size()] = (uint8_t)rng(); }
    else if (mode == 1) { d.resize(rng() % 4000); for (auto& b : d) b = (uint8_t)rng(); }
    else if (mode == 2) { size_t p = rng() % d.size(); std::vector<uint8_t> ins(1 + rng() % 50); for (auto& b : ins) b = (uint8_t)rng(); d.insert(d.begin() + p, ins.begin(), ins.end()); }
    else if (mode == 4) {  // huge but VALID length in front of a length-delimited field, anywhere in the stream
      int k = 1 + rng() % 6;
      while (k--) { size_t p = rng() % (d.size() - 16); d[p] = (uint8_t)(((1 + rng() % 6) << 3) | 2); huge_varint(&d[p + 1]); }
    } else if (mode == 5) {  // the same right behind an envelope header, where field walks start
      for (size_t p = 0; p + 20 < d.size(); p++)
        if (d[p] == 0x0D && d[p + 1] == 0xA4 && rng() % 8 == 0) { d[p + 5] = (uint8_t)(((1 + rng() % 6) << 3) | 2); huge_varint(&d[p + 6]); }
    }
    else { size_t p = rng() % (d.size() - 16); uint8_t h[12] = {0x0d, 0xa4, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff}; for (int k = 0; k < 12; k++) d[p + k] = h[k]; }
    run(d);
  }
  printf("ok %ld envelopes %ld cols\n", env, frames);
}
