// stencil_kernels.cu -- per-block stencil sweeps (placeholder until the
// kernels land; every entry fails loudly rather than falling back).
#include "cup_internal.h"
namespace cup {
int advdiff(CupCtx *) { set_error("advdiff: not built yet"); return CUP_ERR_UNSUPPORTED; }
int projection(CupCtx *, CupSolveInfo *) { set_error("projection: not built yet"); return CUP_ERR_UNSUPPORTED; }
int stencil_run(CupCtx *, CupStencilId, const long long *, long long) { set_error("stencil_run: not built yet"); return CUP_ERR_UNSUPPORTED; }
}
