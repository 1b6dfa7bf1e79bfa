// nka_res_items.h -- work items of the residual strip kernel (host side, plain C++; also compiled by the CPU
// harness tests/model/res_items_test.cpp).
//
// An item is (band of `band` GRID diagonals, strip of `cols` grid columns).  Strip s has cells on diagonals
// [cols*s, cols*s + cols - 1 + ny); band b is [b*band, (b+1)*band), the last one cut at nx + ny - 1.  Only the
// pairs that hold cells are numbered, band by band and strip by strip inside a band, so that neighbouring strips of
// one band -- which write the two halves of the 32-byte sectors straddling their common edge -- are handed out one
// after the other.  first[b] = number of the first item of band b (first[nbands] = count),
// s_lo[b] = first strip with cells in band b; the kernel finds an item's band by binary search in `first`.

#pragma once

#include <stddef.h>

#include <vector>

struct NkaResItems {
  int nbands = 0;
  size_t count = 0;
  std::vector<unsigned> first;     // [nbands + 1]
  std::vector<int> s_lo;           // [nbands]
};

inline NkaResItems nka_res_items(int nx, int ny, int cols, int band)
{
  NkaResItems it;
  const int nstrips = (nx + cols - 1) / cols;
  it.nbands = (nx + ny - 1 + band - 1) / band;
  it.first.resize((size_t)it.nbands + 1);
  it.s_lo.resize((size_t)it.nbands);
  size_t count = 0;
  for (int b = 0; b < it.nbands; ++b) {
    const long long lo_t = (long long)b * band, hi_t = lo_t + band;   // [lo_t, hi_t)
    long long s0 = lo_t - (cols - 1) - ny;                             // cols*s > s0  <=>  the strip has not ended before the band
    s0 = s0 < 0 ? 0 : s0 / cols + 1;
    long long s1 = (hi_t - 1) / cols;                                  // cols*s <= hi_t - 1  <=>  the strip has started by the band's end
    if (s1 > nstrips - 1) s1 = nstrips - 1;
    it.first[b] = (unsigned)count;
    it.s_lo[b] = (int)s0;
    if (s1 >= s0) count += (size_t)(s1 - s0 + 1);
  }
  it.first[it.nbands] = (unsigned)count;
  it.count = count;
  return it;
}
