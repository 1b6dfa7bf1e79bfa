// nka_pass_a.cu -- instantiations of the pass A kernel (one per number of streamed columns).
#include "nka_dispatch.h"
#include "nka_kernels.cuh"

static PassAFn g_pass_a[NKA_MAXSLOT + 1][3];

template <int N> struct FillA {
  static void run() {
    g_pass_a[N][1] = nka_pass_a<N, 1>;
    g_pass_a[N][2] = nka_pass_a<N, 2>;
    FillA<N - 1>::run();
  }
};
template <> struct FillA<0> { static void run() {} };

PassAFn nka_get_pass_a(int nc, int v)
{
  static bool ready = false;
  if (!ready) { FillA<NKA_INSTANTIATE_MAX>::run(); ready = true; }
  if (nc < 1 || nc > NKA_MAXSLOT || v < 1 || v > 2) return nullptr;
  return g_pass_a[nc][v];
}

#ifdef NKA_TRACE
// tools/pass_a_trace.py (tuning builds only): read and re-arm the phase stamps
extern "C" void nka_debug_trace(unsigned long long out[8])
{
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_nka_trace, sizeof(unsigned long long) * 8);
  unsigned long long init[8] = {~0ull, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_nka_trace, init, sizeof init);
}
#endif
