// Deformable-conv tcgen05 path (gather producer -> swizzled smem K-slabs -> tcgen05.mma).  Not built yet in this
// revision: op_deform() falls through to the SIMT gather kernel (kernels_simt.cu) in both precisions.
#include "brn_common.h"

namespace brn {
bool tc_deform_supported(const DeformArgs&) { return false; }
void tc_deform(const LaunchCtx&, const DeformArgs&) { throw Error(7, "tc_deform: not implemented"); }
}  // namespace brn
