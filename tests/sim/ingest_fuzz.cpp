// Host-only fuzz harness for the text readers of breakmer_b200/csrc/ingest.cuh, built with
// g++ -fsanitize=address,undefined by tests/test_ingest.py.  Random texts over alphabets that
// exercise every branch of the parsers (line structure, header fields, white space, empty
// records) go through the single-text parsers and the batch layout; the harness checks the
// layout invariants and relies on the sanitizers for memory errors.  Test infrastructure only.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <string>
#include <vector>

#define BK_SIM 1
#include "../../include/breakmer_b200.h"
#include "../../breakmer_b200/csrc/ingest.cuh"

using namespace bk;

static std::string random_text(std::mt19937& rng, int mode) {
  static const char* alphabets[] = {
      "ACGTNacgt\n\n\n@>+:/_#01 \t\r",                 // everything
      "ACGT\n",                                          // sequence lines only
      "@a:1:2:3:4/1_0\nACGT\n+\nIIII\n",                 // characters of a valid record, shuffled
      "\n\r\t @>:",                                      // structure characters only
  };
  const char* al = alphabets[mode & 3];
  const size_t na = strlen(al);
  const int n = (int)(rng() % 400);
  std::string s;
  if (mode & 4) {                                        // start from well-formed records and corrupt a few bytes
    const int recs = (int)(rng() % 8);
    for (int i = 0; i < recs; ++i) {
      std::string seq;
      for (int j = (int)(rng() % 50); j > 0; --j) seq.push_back("ACGT"[rng() & 3]);
      s += "@M:" + std::to_string(rng() % 9) + ":" + std::to_string(rng() % 99) + ":" + std::to_string(rng() % 999) + ":" +
           std::to_string(rng() % 9999) + "/" + std::to_string(1 + (rng() & 1)) + "_" + std::to_string(rng() & 1) + "\n" + seq +
           "\n+\n" + std::string(seq.size(), 'I') + "\n";
    }
    for (int j = (int)(rng() % 4); j > 0 && !s.empty(); --j) s[rng() % s.size()] = al[rng() % na];
    return s;
  }
  for (int i = 0; i < n; ++i) s.push_back(al[rng() % na]);
  return s;
}

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "ingest_fuzz: check failed at line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  std::mt19937 rng(12345);
  long n_ok = 0, n_err = 0;
  for (int it = 0; it < iters; ++it) {
    const int R = 1 + (int)(rng() % 6);
    std::vector<std::string> txt((size_t)R * 4);
    std::vector<TextView> v[4];
    for (int s = 0; s < 4; ++s) v[s].resize(R);
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < 4; ++s) {
        std::string& t = txt[(size_t)r * 4 + s];
        t = random_text(rng, (int)(rng() % 8));
        const bool absent = (rng() % 7) == 0;
        v[s][r] = absent ? TextView{nullptr, 0} : TextView{t.data(), t.size()};
      }
    Ingest g;
    g.n_threads = 1 + (int)(rng() % 4);
    g.pinned = false;
    bk_batch_input in;
    IngestText text{};
    try {
      ingest_texts(g, R, v[0].data(), v[1].data(), v[2].data(), (rng() & 1) ? v[3].data() : nullptr, &in, &text);
    } catch (const ApiError& e) {
      CHECK(e.code == BK_ERR_FORMAT);
      ++n_err;
      continue;
    }
    ++n_ok;
    CHECK(in.n_regions == R);
    CHECK(in.read_reg_off[0] == 0 && in.read_reg_off[R] == text.n_reads);
    CHECK(in.ref_off[0] == 0 && in.read_off[0] == 0 && in.sc_off[0] == 0);
    for (int r = 0; r < R; ++r) {
      CHECK(in.ref_off[r] <= in.ref_off[r + 1]);
      CHECK(in.read_reg_off[r] <= in.read_reg_off[r + 1] && in.sc_reg_off[r] <= in.sc_reg_off[r + 1]);
      int32_t mx = 0;
      for (int64_t i = in.read_reg_off[r]; i < in.read_reg_off[r + 1]; ++i) {
        const int64_t len = in.read_off[i + 1] - in.read_off[i];
        CHECK(len >= 0 && text.id_off[i] <= text.id_off[i + 1] && text.qual_off[i] <= text.qual_off[i + 1]);
        CHECK(text.id_off[i + 1] - text.id_off[i] >= 10);          // "a:1:2:3:4/" at the very least
        CHECK(in.read_flags[i] <= 1);
        if (len > mx) mx = (int32_t)len;
        for (int64_t b = in.read_off[i]; b < in.read_off[i + 1]; ++b) CHECK(in.read_bases[b] != '\n');
      }
      CHECK(in.read_len[r] == mx);
    }
    for (int64_t i = 0; i < in.sc_reg_off[R]; ++i) CHECK(in.sc_off[i] <= in.sc_off[i + 1]);
    if (in.normal_bases)
      for (int64_t i = 0; i < in.normal_reg_off[R]; ++i) CHECK(in.normal_off[i] <= in.normal_off[i + 1]);
  }
  printf("ingest_fuzz: %d batches, %ld parsed, %ld rejected (BK_ERR_FORMAT)\n", iters, n_ok, n_err);
  return (n_ok > 0 && n_err > 0) ? 0 : 2;
}
