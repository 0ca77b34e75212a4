// Interface of the tcgen05 (5th-gen tensor core, TMEM accumulator, TMA-fed) projection kernels, proj_tc.cu.
#pragma once
#include "common.cuh"

namespace b200gat {

bool proj_tc_fwd_supported(const b200gat_layer& L, int64_t N);
size_t proj_tc_fwd_workspace_bytes(const b200gat_layer& L, int64_t N);
size_t proj_tc_split_bytes(const b200gat_layer& L, int64_t N);   // the x_split blob kept from forward to backward
int proj_tc_fwd(const b200gat_proj_fwd_args& a, cudaStream_t stream);

// the stand-alone attention-logit pass (proj.cu), used when heads straddle the GEMM's output tiles
int launch_logits(const b200gat_proj_fwd_args& a, cudaStream_t stream);

bool proj_tc_bwd_supported(const b200gat_layer& L, int64_t N);
bool proj_tc_can_fuse_prep(const b200gat_layer& consumer, int64_t N, const b200gat_layer& producer);
size_t proj_tc_bwd_workspace_bytes(const b200gat_layer& L, int64_t N);
int proj_tc_bwd(const b200gat_proj_bwd_args& a, cudaStream_t stream);

}  // namespace b200gat
