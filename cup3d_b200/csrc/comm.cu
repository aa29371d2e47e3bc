// comm.cu -- one rank per GPU over NCCL (placeholder until the halo exchange lands).
#include "cup_internal.h"
namespace cup {
int comm_init(CupCtx *, int, int nranks, const void *, size_t) {
  if (nranks == 1) return CUP_OK;
  set_error("comm_init: multi-rank not built yet"); return CUP_ERR_UNSUPPORTED;
}
int comm_unique_id(void *, size_t) { set_error("comm_unique_id: not built yet"); return CUP_ERR_UNSUPPORTED; }
}
