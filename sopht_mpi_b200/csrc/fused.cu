// Fused streaming vorticity update (filled in later in the round).
#include "sb200_common.h"
extern "C" int sb200_vorticity_rhs_fused_3d(const sb200_grid_t*, void*, const void*, const void*,
                                            const void*, double, double, void*) {
  sb_set_error("fused kernel not built yet");
  return -1;
}
