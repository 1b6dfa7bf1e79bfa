// nka_dispatch.h -- kernel tables shared by the translation units of libnka_b200.so.
#pragma once
#include <stddef.h>
#include "nka_state.h"

typedef void (*PassAFn)(const double*, const double*, size_t, size_t, NkaDevState*, double*, unsigned*, double*, int, NkaPeerCtx*,
                        const double*, unsigned);
typedef void (*PassBFn)(double*, double*, double*, size_t, size_t, const NkaDevState*);

#ifndef NKA_INSTANTIATE_MAX
#define NKA_INSTANTIATE_MAX NKA_MAXSLOT     // tuning builds instantiate fewer sizes to compile faster
#endif

// nc in 1..NKA_MAXSLOT, nz in 0..NKA_MAXSLOT-1, v in {1,2}; nullptr if not instantiated in this build
PassAFn nka_get_pass_a(int nc, int v);
PassBFn nka_get_pass_b(int nz, int v);
// experimental (NKA_PASS_B_TMA=1): operand tiles staged by cp.async.bulk; nullptr if nz is not instantiated
PassBFn nka_get_pass_b_tma(int nz, int* threads, int* smem_bytes);
