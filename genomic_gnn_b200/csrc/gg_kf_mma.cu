// K4 (tensor path): TMA-fed tcgen05 Gram tiles with the fused exact-threshold epilogue.
#include "gg_kf_emit.cuh"

namespace ggkf {

int emit_tcgen05(const KfArgs& a, cudaStream_t s) {
  (void)a;
  (void)s;
  gg::set_error("tcgen05 KF kernel not built yet");
  return GG_ERR_ARG;
}

}  // namespace ggkf
