// nka_pass_b.cu -- instantiations of the pass B kernel (one per number of streamed Z columns).
#include "nka_dispatch.h"
#include "nka_kernels.cuh"

static PassBFn g_pass_b[NKA_MAXSLOT + 1][3];

template <int N> struct FillB {
  static void run() {
    g_pass_b[N - 1][1] = nka_pass_b<N - 1, 1>;
    g_pass_b[N - 1][2] = nka_pass_b<N - 1, 2>;
    FillB<N - 1>::run();
  }
};
template <> struct FillB<0> { static void run() {} };

PassBFn nka_get_pass_b(int nz, int v)
{
  static bool ready = false;
  if (!ready) { FillB<NKA_INSTANTIATE_MAX>::run(); ready = true; }
  if (nz < 0 || nz >= NKA_MAXSLOT || v < 1 || v > 2) return nullptr;
  return g_pass_b[nz][v];
}
