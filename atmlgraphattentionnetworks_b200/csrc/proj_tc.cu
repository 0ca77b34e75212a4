// tcgen05 projection — placeholder until the TMA/TMEM kernel lands: reports every shape as unsupported so that
// proj.cu routes to the fp32 CUDA-core GEMM.
#include "proj_tc.cuh"

namespace b200gat {

bool proj_tc_fwd_supported(const b200gat_layer&, int64_t) { return false; }
size_t proj_tc_fwd_workspace_bytes(const b200gat_layer&, int64_t) { return 0; }
int proj_tc_fwd(const b200gat_proj_fwd_args&, cudaStream_t) {
  return fail(B200GAT_E_UNSUPPORTED, "proj_tc_fwd: not built");
}
bool proj_tc_bwd_supported(const b200gat_layer&, int64_t) { return false; }
size_t proj_tc_bwd_workspace_bytes(const b200gat_layer&, int64_t) { return 0; }
int proj_tc_bwd(const b200gat_proj_bwd_args&, cudaStream_t) {
  return fail(B200GAT_E_UNSUPPORTED, "proj_tc_bwd: not built");
}

}  // namespace b200gat
