// tcgen05 (UMMA) grid kernel -- placeholder until the 3xTF32 kernel lands; the dispatcher falls back to SIMT.
#include "bogp_internal.h"

namespace bogp {
bool umma_supported(const bogp_modeset_s*) { return false; }
int launch_grid_umma(bogp_modeset_s*, const double*, int, const double*, int, int, int, float*, cudaStream_t) {
  set_error("UMMA engine not built");
  return BOGP_ERR_UNSUPPORTED;
}
}  // namespace bogp
