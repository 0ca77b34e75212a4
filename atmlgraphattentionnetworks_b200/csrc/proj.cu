// K1 / K4 — the dense per-head projection (GAT.py:42-52) and its backward.
//   forward : Wh[N,Dp] = X[N,F] · Wp[Dp,F]^T + bw ;  s_src = <Wh_h, a1_h> + b1_h ;  s_dst = <Wh_h, a2_h> + b2_h
//   backward: gX[N,F] = gT[N,Dp] · Wp[Dp,F] ;  gW[Dp,F] = gT^T · X
// Dispatch: the tcgen05 (3xFP16 split) kernel in proj_tc.cu takes the shapes it supports; everything else runs on
// the fp32 CUDA-core GEMM below.  Both produce fp32 results within the 1e-5 parity bar.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "proj_tc.cuh"

namespace b200gat {

// one lane GROUP of G = pow2ceil(min(c_pad / 4, 32)) lanes per (node, head): s_src / s_dst from the freshly written Wh row
// (L2-resident at this point).  (One warp per node looping over the heads left 2 of 32 lanes busy on the 8 x 8 layers of
// GATNet: 21.8 us for 15 k nodes; the group form keeps every lane loading.)
template <int G>
__global__ void __launch_bounds__(256)
logits_kernel(const float* __restrict__ wh, const float* __restrict__ a1, const float* __restrict__ a2,
              const float* __restrict__ b1, const float* __restrict__ b2, float* __restrict__ s_src,
              float* __restrict__ s_dst, int64_t items, int H, int Cp) {
  constexpr int GPW = 32 / G;
  constexpr int U = 4;                                     // items per lane group and trip: their loads are all in flight
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;     // together (one item per trip left the kernel
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;   // latency-bound: 82 us for 342 k items)
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int Q = Cp >> 2;
  for (int64_t base = warp * (GPW * U); base < items; base += nwarps * (GPW * U)) {
    float d1[U], d2[U];
    int hh[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t item = base + u * GPW + gi;
      valid[u] = item < items;
      hh[u] = valid[u] ? static_cast<int>(item % H) : 0;
      d1[u] = d2[u] = 0.f;
    }
    for (int q = gl; q < Q; q += G) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        v[u] = valid[u] ? ldg4(wh + (base + u * GPW + gi) * Cp + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);   // rows are [N, H, Cp]
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 p = ldg4(a1 + hh[u] * Cp + 4 * q);
        const float4 r = ldg4(a2 + hh[u] * Cp + 4 * q);
        d1[u] += v[u].x * p.x + v[u].y * p.y + v[u].z * p.z + v[u].w * p.w;
        d2[u] += v[u].x * r.x + v[u].y * r.y + v[u].z * r.z + v[u].w * r.w;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float s1 = group_sum<G>(d1[u]), s2 = group_sum<G>(d2[u]);
      if (valid[u] && gl == 0) {
        const int64_t item = base + u * GPW + gi;
        s_src[item] = s1 + b1[hh[u]];
        s_dst[item] = s2 + b2[hh[u]];
      }
    }
  }
}

// ---- heads of >= 9 float4 slots (c_pad > 32): a lane group per HEAD, walking the nodes
template <int G>
__global__ void __launch_bounds__(256, 3)
logits_head_kernel(const float* __restrict__ wh, const float* __restrict__ a1, const float* __restrict__ a2,
              const float* __restrict__ b1, const float* __restrict__ b2, float* __restrict__ s_src,
              float* __restrict__ s_dst, int64_t N, int H, int Cp) {
  constexpr int GPW = 32 / G;
  constexpr int U = 8;                                     // nodes per lane group and trip: their loads are all in flight
  const int lane = threadIdx.x & 31, gl = lane & (G - 1);  // together (one item per trip left the kernel latency-bound)
  // A lane group keeps ONE head for the whole kernel (its slices of a1 / a2 stay in registers: re-loading them per item
  // was two L1 loads per row load, 2.4 TB/s on the 6 x 121 PPI layer) and walks the nodes; adjacent groups take adjacent
  // heads of the same node, so the grid still streams one window of consecutive rows.
  const int64_t group = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = (int64_t(gridDim.x) * blockDim.x) / G;
  const int64_t per_head = ngroups / H;                    // groups per head (the host launches >= H groups)
  if (group >= per_head * H) return;                       // < H left-over groups idle (whole groups: shuffles stay group-local)
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
  const int h = static_cast<int>(group % H);
  const int Q = Cp >> 2;
  const float bb1 = __ldg(b1 + h), bb2 = __ldg(b2 + h);
  if (Q <= G) {                                            // the usual case: one 128-bit slot per lane
    const bool on = gl < Q;
    const int c = 4 * (on ? gl : 0);
    const float4 p = ldg4(a1 + h * Cp + c), r = ldg4(a2 + h * Cp + c);
    for (int64_t n0 = (group / H) * U; n0 < N; n0 += per_head * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t n = n0 + u < N ? n0 + u : N - 1;       // tail: re-read the last node, never stored
        v[u] = ldg4(wh + (n * H + h) * Cp + c);              // rows are [N, H, Cp]
      }
      float d1[U], d2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        d1[u] = on ? v[u].x * p.x + v[u].y * p.y + v[u].z * p.z + v[u].w * p.w : 0.f;
        d2[u] = on ? v[u].x * r.x + v[u].y * r.y + v[u].z * r.z + v[u].w * r.w : 0.f;
      }
#pragma unroll
      for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          d1[u] += __shfl_xor_sync(gmask, d1[u], o, G);
          d2[u] += __shfl_xor_sync(gmask, d2[u], o, G);
        }
      }
      if (gl < U && n0 + gl < N) {                            // lane u of the group writes node n0 + u (G >= U: see the host)
        float s1 = d1[0], s2 = d2[0];
#pragma unroll
        for (int u = 1; u < U; ++u)
          if (gl == u) { s1 = d1[u]; s2 = d2[u]; }
        s_src[(n0 + gl) * H + h] = s1 + bb1;
        s_dst[(n0 + gl) * H + h] = s2 + bb2;
      }
    }
    return;
  }
  // wide heads (Cp > 128, G = 32): several slots per lane, U / 2 nodes in flight
  constexpr int UW = U / 2;
  for (int64_t n0 = (group / H) * UW; n0 < N; n0 += per_head * UW) {
    float d1[UW], d2[UW];
#pragma unroll
    for (int u = 0; u < UW; ++u) d1[u] = d2[u] = 0.f;
    for (int q = gl; q < Q; q += G) {
      const float4 p = ldg4(a1 + h * Cp + 4 * q), r = ldg4(a2 + h * Cp + 4 * q);
      float4 v[UW];
#pragma unroll
      for (int u = 0; u < UW; ++u) {
        const int64_t n = n0 + u < N ? n0 + u : N - 1;
        v[u] = ldg4(wh + (n * H + h) * Cp + 4 * q);
      }
#pragma unroll
      for (int u = 0; u < UW; ++u) {
        d1[u] += v[u].x * p.x + v[u].y * p.y + v[u].z * p.z + v[u].w * p.w;
        d2[u] += v[u].x * r.x + v[u].y * r.y + v[u].z * r.z + v[u].w * r.w;
      }
    }
#pragma unroll
    for (int u = 0; u < UW; ++u) {
      const float s1 = group_sum<G>(d1[u]), s2 = group_sum<G>(d2[u]);
      if (gl == 0 && n0 + u < N) {
        s_src[(n0 + u) * H + h] = s1 + bb1;
        s_dst[(n0 + u) * H + h] = s2 + bb2;
      }
    }
  }
}

// fp32 rows -> bf16 copy (the CUDA-core projection path of the bf16 gather mode; the tensor-core path writes it in its epilogue)
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n4) {
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n4; t += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = ldg4(src + 4 * t);
    *reinterpret_cast<uint2*>(dst + 4 * t) = pack_bf16x4(v.x, v.y, v.z, v.w);
  }
}

int launch_logits(const b200gat_proj_fwd_args& a, cudaStream_t stream) {
  const int64_t N = a.num_nodes;
  const int H = static_cast<int>(a.layer.heads), Cp = static_cast<int>(a.layer.c_pad), Q = Cp / 4;
  const int64_t items = N * H;
  const int threads = 256;
  const int64_t cap = int64_t(sm_count()) * 8;
  auto grid = [&](int g) {
    const int64_t want = ceil_div(ceil_div(items, (32 / g) * 4), threads / 32);
    return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  };
  auto grid_head = [&](int g) {
    // one lane group per (head, 8-node strip): at least H groups, at most `cap` CTAs
    const int64_t groups = ceil_div(N, 8) * H, gpc = threads / g;
    const int64_t want = ceil_div(groups, gpc), least = ceil_div(int64_t(H), gpc);
    const int64_t n = want < cap ? want : cap;
    return static_cast<int>(n > least ? n : least);
  };
#define B200GAT_LOGITS(G) logits_kernel<G><<<grid(G), threads, 0, stream>>>(a.wh, a.a1, a.a2, a.b1, a.b2, a.s_src, a.s_dst, items, H, Cp)
#define B200GAT_LOGITS_HEAD(G) logits_head_kernel<G><<<grid_head(G), threads, 0, stream>>>(a.wh, a.a1, a.a2, a.b1, a.b2, a.s_src, a.s_dst, N, H, Cp)
  if (Q <= 1) B200GAT_LOGITS(1);
  else if (Q <= 2) B200GAT_LOGITS(2);
  else if (Q <= 4) B200GAT_LOGITS(4);
  else if (Q <= 8) B200GAT_LOGITS(8);
  else if (Q <= 16) B200GAT_LOGITS_HEAD(16);
  else B200GAT_LOGITS_HEAD(32);
#undef B200GAT_LOGITS_HEAD
#undef B200GAT_LOGITS
  return check_launch("logits_kernel");
}

}  // namespace b200gat

using namespace b200gat;

extern "C" size_t b200gat_proj_fwd_workspace_bytes(const b200gat_layer* L, int64_t N) {
  if (!L || N < 0) return 0;
  return proj_tc_fwd_workspace_bytes(*L, N);
}

extern "C" size_t b200gat_proj_split_bytes(const b200gat_layer* L, int64_t N) {
  if (!L || N < 0) return 0;
  return proj_tc_split_bytes(*L, N);
}

extern "C" int b200gat_proj_fwd(const b200gat_proj_fwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "proj_fwd: NULL args");
  int rc = validate_layer(a->layer);
  if (rc) return rc;
  const b200gat_layer& L = a->layer;
  const int64_t N = a->num_nodes, F = L.in_channels, H = L.heads, Cp = L.c_pad, Dp = H * Cp;
  B200GAT_REQUIRE(N >= 0, B200GAT_E_SHAPE, "proj_fwd: negative num_nodes");
  if (N == 0) return 0;
  B200GAT_REQUIRE(a->x && a->w && a->bw && a->a1 && a->a2 && a->b1 && a->b2 && a->wh && a->s_src && a->s_dst,
                  B200GAT_E_NULL, "proj_fwd: NULL pointer");
  B200GAT_REQUIRE(a->ldx >= F, B200GAT_E_SHAPE, "proj_fwd: ldx < in_channels");
  B200GAT_REQUIRE(aligned16(a->wh) && aligned16(a->a1) && aligned16(a->a2), B200GAT_E_ALIGN,
                  "proj_fwd: wh/a1/a2 must be 16-byte aligned");
  B200GAT_REQUIRE(a->x_activation == ACT_NONE || a->x_activation == ACT_ELU, B200GAT_E_UNSUPPORTED,
                  "proj_fwd: unknown x_activation %d", a->x_activation);
  if (proj_tc_fwd_supported(L, N)) return proj_tc_fwd(*a, stream);
  B200GAT_REQUIRE(a->num_peers == 0, B200GAT_E_UNSUPPORTED, "proj_fwd: wh_peers needs the tensor-core path (shape too small)");
  rc = gemm_simt<true, true>(a->x, a->ldx, a->w, F, a->wh, Dp, a->bw, N, Dp, F, 1, stream, a->x_activation);
  if (rc) return rc;
  if (a->wh_bf16) {
    const int64_t n4 = N * Dp / 4, want = ceil_div(n4, 256), cap = int64_t(sm_count()) * 8;
    to_bf16_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, stream>>>(a->wh, static_cast<uint16_t*>(a->wh_bf16), n4);
    if ((rc = check_launch("to_bf16_kernel"))) return rc;
  }
  return launch_logits(*a, stream);
}

extern "C" size_t b200gat_proj_bwd_workspace_bytes(const b200gat_layer* L, int64_t N) {
  if (!L || N < 0) return 0;
  return proj_tc_bwd_workspace_bytes(*L, N);
}

extern "C" int b200gat_proj_bwd_can_fuse_prep(const b200gat_layer* consumer, int64_t N, const b200gat_layer* producer) {
  if (!consumer || !producer || N <= 0) return 0;
  return proj_tc_can_fuse_prep(*consumer, N, *producer) ? 1 : 0;
}

extern "C" int b200gat_proj_bwd(const b200gat_proj_bwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "proj_bwd: NULL args");
  int rc = validate_layer(a->layer);
  if (rc) return rc;
  const b200gat_layer& L = a->layer;
  const int64_t N = a->num_nodes, F = L.in_channels, Dp = L.heads * L.c_pad;
  B200GAT_REQUIRE(N >= 0, B200GAT_E_SHAPE, "proj_bwd: negative num_nodes");
  B200GAT_REQUIRE(a->parts >= 0 && a->parts <= B200GAT_PROJ_BWD_GW, B200GAT_E_SHAPE, "proj_bwd: unknown parts %d", a->parts);
  B200GAT_REQUIRE((a->g_w || a->parts == B200GAT_PROJ_BWD_GX) && a->w, B200GAT_E_NULL, "proj_bwd: NULL pointer");
  if (N == 0) {
    if (a->parts == B200GAT_PROJ_BWD_GX) return 0;
    cudaError_t e = cudaMemsetAsync(a->g_w, 0, size_t(Dp) * F * sizeof(float), stream);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "proj_bwd: memset: %s", cudaGetErrorString(e));
    return 0;
  }
  B200GAT_REQUIRE(a->x_activation == ACT_NONE || a->x_activation == ACT_ELU, B200GAT_E_UNSUPPORTED,
                  "proj_bwd: unknown x_activation %d", a->x_activation);
  B200GAT_REQUIRE((a->g_t || a->g_t_split) && (a->x || a->x_split), B200GAT_E_NULL, "proj_bwd: NULL pointer");
  B200GAT_REQUIRE(a->ldx >= F && (!a->g_x || a->ldgx >= F), B200GAT_E_SHAPE, "proj_bwd: leading dimension < in_channels");
  if (proj_tc_bwd_supported(L, N)) return proj_tc_bwd(*a, stream);
  B200GAT_REQUIRE(a->g_t && a->x, B200GAT_E_NULL, "proj_bwd: the CUDA-core path needs fp32 g_t and x");
  B200GAT_REQUIRE(!a->fuse_prep, B200GAT_E_UNSUPPORTED, "proj_bwd: fuse_prep needs the tensor-core path");
  if (a->g_x && a->parts != B200GAT_PROJ_BWD_GW) {
    rc = gemm_simt<true, false>(a->g_t, Dp, a->w, F, a->g_x, a->ldgx, nullptr, N, F, Dp, 1, stream);
    if (rc) return rc;
  }
  if (a->parts == B200GAT_PROJ_BWD_GX) return 0;
  // gW[Dp,F] = sum_n gT[n,:]^T X[n,:] — reduction over nodes, split so that the grid covers the machine
  const int64_t tiles = ceil_div(Dp, GM) * ceil_div(F, GN);
  int64_t splits = ceil_div(int64_t(sm_count()) * 4, tiles);
  const int64_t max_splits = ceil_div(N, 64);       // short k-loops: the tiny shapes that land here are latency-bound
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return gemm_simt<false, false>(a->g_t, Dp, a->x, a->ldx, a->g_w, F, nullptr, Dp, F, N, static_cast<int>(splits), stream,
                                 ACT_NONE, a->x_activation);
}
