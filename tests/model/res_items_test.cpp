// CPU harness for nka_b200/csrc/nka_res_items.h: the numbering of the residual kernel's (band, strip) items against a
// brute-force enumeration that mirrors the kernel's own range clipping (ex_residual_strip_kernel: t0, t1).
#include <stdio.h>
#include <stdlib.h>

#include <set>
#include <utility>

#include "../../nka_b200/csrc/nka_res_items.h"

static int check(int nx, int ny, int cols, int band)
{
  const NkaResItems it = nka_res_items(nx, ny, cols, band);
  const int nstrips = (nx + cols - 1) / cols, tmax = nx + ny - 2;
  std::set<std::pair<int, int>> want;
  long long cells = 0;
  for (int b = 0; b < it.nbands; ++b)
    for (int s = 0; s < nstrips; ++s) {
      int t0 = b * band, t1 = t0 + band;
      const int tend = s * cols + cols - 1 + ny;
      if (t0 < s * cols) t0 = s * cols;
      if (t1 > tend) t1 = tend;
      if (t1 > tmax + 1) t1 = tmax + 1;
      if (t0 < t1) {
        want.insert({b, s});
        for (int t = t0; t < t1; ++t)                       // cells of strip s on diagonal t
          for (int j = s * cols; j < s * cols + cols && j < nx; ++j) cells += (t - j >= 0 && t - j < ny);
      }
    }
  if (cells != (long long)nx * ny) { printf("FAIL cover %d x %d band %d: %lld cells\n", nx, ny, band, cells); return 1; }
  std::set<std::pair<int, int>> got;
  for (size_t item = 0; item < it.count; ++item) {
    int lo = 0, hi = it.nbands - 1;                         // the kernel's binary search
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (it.first[mid] <= item) lo = mid; else hi = mid - 1;
    }
    const int b = lo, s = it.s_lo[b] + (int)(item - it.first[b]);
    if (s < 0 || s >= nstrips || !got.insert({b, s}).second) { printf("FAIL item %zu -> (%d, %d)\n", item, b, s); return 1; }
  }
  // every pair with cells is numbered; numbered pairs without cells (possible at the clipped ends) are allowed but rare
  for (const auto& p : want)
    if (!got.count(p)) { printf("FAIL missing (%d, %d) for %d x %d band %d\n", p.first, p.second, nx, ny, band); return 1; }
  if (got.size() > want.size() + (size_t)it.nbands) { printf("FAIL too many empty items\n"); return 1; }
  return 0;
}

int main()
{
  const int shapes[][2] = {{3, 3}, {5, 4}, {4, 9}, {31, 17}, {32, 32}, {33, 70}, {50, 50}, {96, 40}, {257, 129}, {300, 300},
                           {700, 64}, {9700, 7}, {1000, 7}, {7, 1000}, {4096, 4096}, {2048, 129}, {61, 61}, {30, 30}, {29, 31}};
  const int bands[] = {1, 2, 24, 32, 33, 64, 128};
  int n = 0;
  for (const auto& sh : shapes)
    for (int band : bands) {
      if ((long long)sh[0] * sh[1] > 4000000 && band < 24) continue;
      if (check(sh[0], sh[1], 30, band)) return 1;
      ++n;
    }
  printf("res items ok: %d geometries\n", n);
  return 0;
}
