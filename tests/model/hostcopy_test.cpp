// CPU harness for nka_b200/csrc/nka_hostcopy.h (the host threads that move a pageable caller's bytes
// into / out of the pinned staging slots).  Compiled and run by tests/test_hostcopy.py; plain C++.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <functional>
#include <thread>
#include <vector>

#include "../../nka_b200/csrc/nka_hostcopy.h"

static uint64_t mix(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv)
{
  const int nthreads = argc > 1 ? atoi(argv[1]) : 3;
  NkaHostCopier hc(nthreads);
  if (hc.threads() != nthreads + 1) { printf("FAIL threads\n"); return 1; }
  // sizes around the 4 MiB threshold below which a plain memcpy is used, odd byte counts, unaligned bases
  const size_t sizes[] = {0, 1, 4095, (4u << 20) - 1, 4u << 20, (4u << 20) + 1, (12u << 20) + 12345, (33u << 20) + 7, 64u << 20};
  std::vector<unsigned char> src((64u << 20) + 64), dst((64u << 20) + 64);
  int rounds = 0;
  for (int rep = 0; rep < 3; ++rep)
    for (size_t bytes : sizes)
      for (int sa = 0; sa < 2; ++sa) {
        const size_t so = sa ? 3 : 0, dof = sa ? 5 : 0;
        for (size_t i = 0; i < bytes; i += 8) {
          const uint64_t v = mix(i + bytes + rep);
          for (size_t b = 0; b < 8 && i + b < bytes; ++b) src[so + i + b] = (unsigned char)(v >> (8 * b));
        }
        for (size_t i = 0; i < bytes + 16; ++i) dst[dof + i] = 0xEE;
        if (dof) for (size_t i = 0; i < dof; ++i) dst[i] = 0xEE;
        hc.copy(dst.data() + dof, src.data() + so, bytes);
        for (size_t i = 0; i < bytes; ++i)
          if (dst[dof + i] != src[so + i]) { printf("FAIL bytes=%zu at %zu\n", bytes, i); return 1; }
        for (size_t i = 0; i < 16; ++i)
          if (dst[dof + bytes + i] != 0xEE) { printf("FAIL overrun bytes=%zu\n", bytes); return 1; }
        for (size_t i = 0; i < dof; ++i)
          if (dst[i] != 0xEE) { printf("FAIL underrun bytes=%zu\n", bytes); return 1; }
        ++rounds;
      }
  // two application threads sharing the pool (two handles updated from two host threads): copies take turns
  {
    const size_t bytes = (24u << 20) + 8;
    std::vector<unsigned char> s2(bytes), d2(bytes), s3(bytes), d3(bytes);
    for (size_t i = 0; i < bytes; ++i) { s2[i] = (unsigned char)mix(i); s3[i] = (unsigned char)mix(i + 77); }
    int bad = 0;
    auto job = [&](const std::vector<unsigned char>& s, std::vector<unsigned char>& d) {
      for (int r = 0; r < 6; ++r) {
        memset(d.data(), 0, bytes);
        hc.copy(d.data(), s.data(), bytes);
        if (memcmp(d.data(), s.data(), bytes) != 0) ++bad;
      }
    };
    std::thread a(job, std::cref(s2), std::ref(d2)), b(job, std::cref(s3), std::ref(d3));
    a.join(); b.join();
    if (bad) { printf("FAIL concurrent callers: %d bad copies\n", bad); return 1; }
  }
  printf("hostcopy ok: %d copies with %d threads\n", rounds, hc.threads());
  return 0;
}
