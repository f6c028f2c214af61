// Tensor-core (tcgen05 / TMEM, 3xTF32) flavour of the reconstruction forward pass.  Placeholder until the
// kernel lands: the dispatcher never selects it.
#ifndef FR_RECON_TC_CUH_
#define FR_RECON_TC_CUH_
#include "fr_common.cuh"
namespace fr {
inline size_t recon_tc_workspace_bytes(int, const BasisGeom&) { return 0; }
inline bool recon_tc_applicable(int, const BasisGeom&, unsigned) { return false; }
inline int launch_recon_fwd_tc(const float*, const float*, const float*, void*, float*, int, int, const BasisGeom&, float,
                               unsigned, int, cudaStream_t) {
  return fail(FR_ERR_UNSUPPORTED, "tensor-core reconstruction path not built");
}
}  // namespace fr
#endif
