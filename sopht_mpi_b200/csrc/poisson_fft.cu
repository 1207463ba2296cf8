// placeholder for the pruned in-kernel FFT pipeline (filled in below)
#include "poisson.h"
int sb_poisson_fft_create(sb200_poisson*, void*) { sb_set_error("fft backend not built"); return -1; }
int sb_poisson_fft_destroy(sb200_poisson*) { return 0; }
int sb_poisson_fft_solve(sb200_poisson*, void*, const void*, int, void*) { return -1; }
int64_t sb_poisson_fft_bytes(const sb200_poisson*) { return 0; }
