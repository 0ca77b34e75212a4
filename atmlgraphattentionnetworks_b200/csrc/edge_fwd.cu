// K2 — fused edge forward (GAT.py:53-67): for every destination row i and head h
//     z_k = s_dst[i,h] + s_src[j_k,h]          e_k = LeakyReLU(z_k)                        (GAT.py:57-58)
//     alpha_k = exp(e_k - max e) / (sum exp(e - max) + 1e-16)                             (GAT.py:60 [PyG] softmax)
//     O[i,h,:] = sum_k alpha_k * mask_k * Wh[j_k,h,:]                                       (GAT.py:61-62, aggr='add')
//     out[i]   = concat ? O[i].reshape(H*C) + bias : mean_h O[i,h,:] + bias                 (GAT.py:63-66,54)
// over the destination-sorted CSR.  No per-edge tensor is written: only out [N,D_out] and the softmax
// statistics rowmax/rowsum [N,H] that let the backward recompute alpha.
//
// Scheduling: one lane GROUP of G = pow2ceil(min(c_pad/4, 32)) lanes per (row, head) work item, 32/G items per
// warp, items = row-major (i, h) so groups of one warp share the row's col[] lines.  Inside an item the G lanes
// first run EDGE-parallel (coalesced col[] load, 4-byte s_src gather, one exp per (edge, head)), then
// FEATURE-parallel: the (j, p) pair of each edge is broadcast by width-G shuffles and every lane gathers its
// 128-bit slices of Wh[j,h,:] (ld.global.nc.v4.f32) into NV float4 accumulators.  The softmax is online per chunk of
// G edges (running max / sum; only rows longer than G ever rescale), so every row is walked exactly once.
//
// Kernels in this file, by degree class and regime:
//   edge_fwd_kernel / edge_fwd_stream_kernel ... lane group per (row, head): L2-resident / HBM-streaming gathers
//   edge_fwd_row_stream_kernel ................. lane group per ROW (all heads), narrow heads on streaming graphs
//   edge_fwd_hub_kernel ........................ one CTA per (row, head) for 512 < in-degree <= 4096
//   edge_fwd_giant_{max,acc,finish}_kernel ..... in-degree > 4096: one CTA per 4096-edge segment, atomics to combine
//   edge_fwd_act_kernel ........................ LogSigmoid / Tanh logits (run_act_func_experiment.py)
//   head_mean_kernel ........................... concat == False, H > 1 (GAT.py:65-66)
#include "common.cuh"
#include "split_blob.cuh"
#include <math.h>

namespace b200gat {

struct EdgeFwdParams {
  int64_t N, items;
  int H, C, Cp, Dp;
  float slope; int act;
  const int32_t* rowptr; const int32_t* col; const int32_t* eid;
  const float* wh; const float* s_src; const float* s_dst; const float* bias;
  const void* wh16;   // optional bf16 copy of wh [N, Dp]: the gathers read it instead (ROW16 instantiations)
  DropoutSpec drop;   // attention dropout (GAT.py:61): mask tensor or in-kernel Philox (common.cuh)
  float* out; int64_t ldo;
  float* rowmax; float* rowsum; float* o_heads;
  int heads_mode;   // 1: write per-head aggregate to o_heads (mean over heads done by head_mean_kernel)
  int vec_out;      // 1: out rows / head offsets are 16-byte aligned -> float4 stores
  uint32_t* out_amax;   // optional: bit pattern of max|out| (atomicMax; zeroed by the host)
  // scheduling by degree (b200gat_graph.hub_rows): rowend[i] = rowptr[i + 1] except for hub rows, which look EMPTY to
  // the row-per-group kernels (rowend[i] = rowptr[i]; they write a placeholder) and are then walked — and their outputs
  // overwritten — by edge_fwd_hub_kernel, one CTA per (row, head).  No compare, no extra register in the hot kernels.
  const int32_t* rowend; const int32_t* hub; int64_t nhub;
  int max_deg;          // largest in-degree (>= any hub row's): rows above B200GAT_GIANT_DEGREE are split into segments
};

// STREAM: the schedule for graphs whose gathered rows come from HBM (b200gat_graph.span): __launch_bounds__(256, 3)
// makes ptxas issue ALL U*NV gathers of a batch back to back (76 registers) — 3x the bytes in flight per warp at 3/4
// of the occupancy.  Measured (tools/microbench/gather_bench.cu, power-law graph): 5.3 vs 3.7 TB/s gathered.  On the
// L2-resident PPI-shaped batch the same build is 13 % SLOWER than the occupancy-first one, hence two instantiations.
template <int G, int NV, bool HAS_MASK, bool STREAM, bool GENERIC = false, bool ROW16 = false>
__device__ __forceinline__ void edge_fwd_body(const EdgeFwdParams& p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int H = p.H, Cp = p.Cp, Q = p.Cp >> 2;
  const int64_t Dp = p.Dp;
  const uint32_t row_bytes = static_cast<uint32_t>(p.Dp) * (ROW16 ? 2u : 4u);
  const float slope = p.slope;
  const int act = GENERIC ? p.act : 0;
  // lanes beyond the head width gather a clamped (valid) column and are never stored
  int off[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) off[v] = 4 * ((gl + v * G < Q) ? gl + v * G : Q - 1);
  const DropoutKey dkey = HAS_MASK ? dropout_key(p.drop) : DropoutKey{0u, 0u, 0u, 0u};

  float amax = 0.f;
  for (int64_t base = warp * GPW; base < p.items; base += nwarps * GPW) {
    const int64_t item = base + gi;
    const bool valid = item < p.items;
    const int64_t i = valid ? item / H : 0;
    const int h = valid ? static_cast<int>(item - i * H) : 0;
    const int beg = valid ? __ldg(p.rowptr + i) : 0;
    const int end = valid ? __ldg(p.rowend + i) : 0;            // hub rows are empty here (edge_fwd_hub_kernel)
    const int deg = end - beg;
    const int maxdeg = GPW == 1 ? deg : __reduce_max_sync(FULL, deg);
    const float sd = valid ? __ldg(p.s_dst + i * H + h) : 0.f;
    const float* ssrc_h = p.s_src + h;

    // ---- one walk over the row, G edges per chunk: chunk max -> online rescale (only rows longer than G ever
    //      rescale) -> exp, row sum, weighted aggregation.  A separate max pass would add two dependent memory
    //      round trips (col -> s_src) per item before the first gather can issue. ----
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float m = -INFINITY, l = 0.f;
    const char* wb[NV];                                   // this lane's column slices of row 0 of head h
#pragma unroll
    for (int v = 0; v < NV; ++v) wb[v] = row_base<ROW16>(p.wh, p.wh16, h * Cp + off[v]);
    // For 256-wide heads (NV = 2) the FIRST gather batch of a chunk is issued as soon as col[] is known — before the
    // dependent s_src gather, the exp and the softmax bookkeeping, none of which the Wh gathers need.  Tuned with
    // tools/microbench/gather_bench.cu (PPI-shaped batch, same loop): NV=2: weights-first U=4 0.279 ms, early U=4 0.317,
    // early U=2 0.233 (a bare gather loop: 0.204); NV=1: weights-first U=8 is the best on the streaming graph.
    // Double-buffering the batches in registers was measured slower (80 registers: 3 instead of 4 CTAs per SM).
    if constexpr (!HAS_MASK && (NV == 2 || (NV == 1 && G == 32)) && !STREAM) {
      constexpr int U = NV == 2 ? 2 : 4;
      for (int k0 = 0; k0 < maxdeg; k0 += G) {
        const int k = beg + k0 + gl;
        const bool ok = k < end;
        int j = static_cast<int>(i);
        if (ok) j = __ldg(p.col + k);
        const int cnt = (maxdeg - k0) < G ? (maxdeg - k0) : G;
        float4 w[U][NV];
        auto load_batch = [&](int t) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            // slots past the chunk's edge count read a valid row (lane t+u's own j, or the destination itself), weight 0
            const int jt = __shfl_sync(FULL, j, t + u, G);
#pragma unroll
            for (int v = 0; v < NV; ++v) w[u][v] = gather4<ROW16>(wb[v], jt, row_bytes);
          }
        };
        load_batch(0);
        float e = -INFINITY;
        if (ok) e = logit_act<GENERIC>(sd + __ldg(ssrc_h + int64_t(j) * H), slope, act);   // (!HAS_MASK path: never GENERIC)
        const float m_new = fmaxf(m, group_max<G>(e));
        if (k0 > 0 && m_new != m) {           // group-uniform; exp(-inf) = 0 covers the "nothing accumulated yet" case
          const float scale = expf(m - m_new);
          l *= scale;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x *= scale; acc[v].y *= scale; acc[v].z *= scale; acc[v].w *= scale;
          }
        }
        m = m_new;
        const float pp = ok ? expf(e - m) : 0.f;
        l += pp;
        for (int t = 0; t < cnt; t += U) {
          float pt[U];
#pragma unroll
          for (int u = 0; u < U; ++u) pt[u] = __shfl_sync(FULL, pp, t + u, G);   // slots past the edges carry pp = 0
          if (t > 0) load_batch(t);
#pragma unroll
          for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              acc[v].x = fmaf(pt[u], w[u][v].x, acc[v].x);
              acc[v].y = fmaf(pt[u], w[u][v].y, acc[v].y);
              acc[v].z = fmaf(pt[u], w[u][v].z, acc[v].z);
              acc[v].w = fmaf(pt[u], w[u][v].w, acc[v].w);
            }
          }
        }
      }
    } else {
      for (int k0 = 0; k0 < maxdeg; k0 += G) {
        const int k = beg + k0 + gl;
        const bool ok = k < end;
        int j = static_cast<int>(i);
        float e = -INFINITY;
        if (ok) {
          j = __ldg(p.col + k);
          if constexpr (GENERIC) e = edge_logit<true>(p.s_dst + i * H, p.s_src + int64_t(j) * H, sd, h, H, slope, act);
          else e = leaky(sd + __ldg(ssrc_h + int64_t(j) * H), slope);
        }
        const float m_new = fmaxf(m, group_max<G>(e));
        if (k0 > 0 && m_new != m) {           // group-uniform; exp(-inf) = 0 covers the "nothing accumulated yet" case
          const float scale = expf(m - m_new);
          l *= scale;
  #pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x *= scale; acc[v].y *= scale; acc[v].z *= scale; acc[v].w *= scale;
          }
        }
        m = m_new;
        float pp = 0.f, pm = 0.f;
        if (ok) {
          pp = expf(e - m);
          pm = pp;
          if (HAS_MASK && (!GENERIC || p.drop.active())) pm *= dropout_mult(p.drop, dkey, __ldg(p.eid + k), h, H);
        }
        l += pp;
        const int cnt = (maxdeg - k0) < G ? (maxdeg - k0) : G;
        // U edges per step: all U*NV 128-bit gathers are issued before the first FMA consumes one (memory-level
        // parallelism; a per-edge branch here serialises load -> FMA -> next load and leaves the kernel latency-bound)
        constexpr int U = NV >= 4 ? 2 : (NV == 1 ? 8 : 4);
        for (int t = 0; t < cnt; t += U) {
          int jt[U];
          float pt[U];
  #pragma unroll
          for (int u = 0; u < U; ++u) {
            jt[u] = __shfl_sync(FULL, j, t + u, G);
            pt[u] = __shfl_sync(FULL, pm, t + u, G);       // slots past the chunk's edges carry pm = 0 ...
            if (U > G && t + u >= cnt) pt[u] = 0.f;        // ... unless the batch is wider than the group (shuffle wraps)
          }
          float4 w[U][NV];
  #pragma unroll
          for (int u = 0; u < U; ++u) {
  #pragma unroll
            for (int v = 0; v < NV; ++v) {
              if (HAS_MASK) {   // dropped edges (60 % under the reference's p = 0.6) contribute nothing: skip the gather
                w[u][v] = pt[u] != 0.f ? gather4<ROW16>(wb[v], jt[u], row_bytes) : make_float4(0.f, 0.f, 0.f, 0.f);
              } else {          // no predicate, no branch: padded slots gather the row's own (valid) Wh and weigh it by 0
                w[u][v] = gather4<ROW16>(wb[v], jt[u], row_bytes);
              }
            }
          }
  #pragma unroll
          for (int u = 0; u < U; ++u) {
  #pragma unroll
            for (int v = 0; v < NV; ++v) {
              acc[v].x = fmaf(pt[u], w[u][v].x, acc[v].x);
              acc[v].y = fmaf(pt[u], w[u][v].y, acc[v].y);
              acc[v].z = fmaf(pt[u], w[u][v].z, acc[v].z);
              acc[v].w = fmaf(pt[u], w[u][v].w, acc[v].w);
            }
          }
        }
      }
    }
    l = group_sum<G>(l);
    if (!valid) continue;
    const float inv = 1.f / (l + 1e-16f);
    if (gl == 0) {
      p.rowmax[i * H + h] = m;
      p.rowsum[i * H + h] = l;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int q = gl + v * G;
      if (q >= Q) continue;
      float4 o = make_float4(acc[v].x * inv, acc[v].y * inv, acc[v].z * inv, acc[v].w * inv);
      if (p.heads_mode) {
        *reinterpret_cast<float4*>(p.o_heads + i * Dp + h * Cp + 4 * q) = o;
      } else {
        const int c = 4 * q;
        const float* b = p.bias + h * p.C + c;
        float* dst = p.out + i * p.ldo + h * p.C + c;
        if (p.vec_out) {
          const float4 bb = ldg4(b);
          const float4 r = make_float4(o.x + bb.x, o.y + bb.y, o.z + bb.z, o.w + bb.w);
          *reinterpret_cast<float4*>(dst) = r;
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))));
        } else {
          const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (c + u < p.C) {
              const float r = ov[u] + __ldg(b + u);
              dst[u] = r;
              amax = fmaxf(amax, fabsf(r));
            }
        }
      }
    }
  }
  if (p.out_amax && !p.heads_mode) warp_atomic_amax(p.out_amax, amax);
}

// two entry points: __launch_bounds__(256) (no minimum: ptxas's own occupancy-first choice, 60-64 registers) and
// __launch_bounds__(256, 3) for the streaming schedule (a minimum of 1 is NOT the same as none: it raises the register
// budget and changed the code of the occupancy-first variant for the worse)
template <int G, int NV, bool HAS_MASK>
__global__ void __launch_bounds__(256) edge_fwd_kernel(const EdgeFwdParams p) { edge_fwd_body<G, NV, HAS_MASK, false>(p); }
template <int G, int NV>
__global__ void __launch_bounds__(256, 3) edge_fwd_stream_kernel(const EdgeFwdParams p) { edge_fwd_body<G, NV, false, true>(p); }
// bf16-stored gathered rows (no mask, LeakyReLU): the same two schedules
template <int G, int NV>
__global__ void __launch_bounds__(256) edge_fwd16_kernel(const EdgeFwdParams p) { edge_fwd_body<G, NV, false, false, false, true>(p); }
template <int G, int NV>
__global__ void __launch_bounds__(256, 3) edge_fwd16_stream_kernel(const EdgeFwdParams p) { edge_fwd_body<G, NV, false, true, false, true>(p); }
// other logit activations (run_act_func_experiment.py): one mask-capable instantiation per geometry (mask may be NULL)
template <int G, int NV>
__global__ void __launch_bounds__(256) edge_fwd_act_kernel(const EdgeFwdParams p) { edge_fwd_body<G, NV, true, false, true>(p); }

// concat == False with H > 1 (GAT.py:65-66): out[i,c] = mean_h O[i,h,c] + bias[c].  One thread per FOUR channels of a node
// (o_heads rows are 16-byte aligned, pad channels zero): the H 128-bit loads of a thread are independent and in flight
// together.
__global__ void __launch_bounds__(256)
head_mean_kernel(const float* __restrict__ o_heads, const float* __restrict__ bias, float* __restrict__ out,
                 int64_t ldo, int64_t N, int H, int C, int Cp, uint32_t* __restrict__ out_amax) {
  const int Q = Cp >> 2;
  const int64_t total = N * Q;
  float amax = 0.f;
  // (node, slot) of item t are stepped incrementally (no 64-bit division per item); up to 8 heads' loads in flight together
  const int64_t nth = int64_t(gridDim.x) * blockDim.x, t0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t di = nth / Q, dq = nth - di * Q;
  int64_t i = t0 / Q, q = t0 - i * Q;
  for (int64_t t = t0; t < total; t += nth, i += di, q += dq) {
    if (q >= Q) { q -= Q; ++i; }
    const int c = 4 * static_cast<int>(q);
    const float* src = o_heads + i * int64_t(H) * Cp + c;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int h0 = 0; h0 < H; h0 += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = h0 + u < H ? ldg4(src + (h0 + u) * Cp) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c + u < C) {
        const float r = sv[u] / static_cast<float>(H) + __ldg(bias + c + u);
        out[i * ldo + c + u] = r;
        amax = fmaxf(amax, fabsf(r));
      }
  }
  if (out_amax) warp_atomic_amax(out_amax, amax);
}

template <int G, int NV>
static int launch_edge_fwd(const EdgeFwdParams& p, bool streaming, cudaStream_t stream) {
  constexpr int GPW = 32 / G;
  const int threads = 256;
  const int64_t want = ceil_div(ceil_div(p.items, GPW), threads / 32);
  const int64_t cap = int64_t(sm_count()) * 8;
  const int blocks = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  if (p.wh16) {                                            // (the entry point only sets wh16 without mask / other activations)
    if (streaming && G >= 16) edge_fwd16_stream_kernel<(G >= 16 ? G : 32), (G >= 16 ? NV : 1)><<<blocks, threads, 0, stream>>>(p);
    else edge_fwd16_kernel<G, NV><<<blocks, threads, 0, stream>>>(p);
  }
  else if (p.act != B200GAT_LOGIT_LEAKY_RELU) edge_fwd_act_kernel<G, NV><<<blocks, threads, 0, stream>>>(p);
  else if (p.drop.active()) edge_fwd_kernel<G, NV, true><<<blocks, threads, 0, stream>>>(p);
  else if (streaming && G >= 16) edge_fwd_stream_kernel<(G >= 16 ? G : 32), (G >= 16 ? NV : 1)><<<blocks, threads, 0, stream>>>(p);
  else edge_fwd_kernel<G, NV, false><<<blocks, threads, 0, stream>>>(p);
  return check_launch("edge_fwd_kernel");
}

// ---- hub rows (in-degree > B200GAT_HUB_DEGREE): one CTA per (row, head).  Its 8 warps walk interleaved 32-edge
// chunks of the row with the same online softmax as above and merge their (max, sum, aggregate) partials through
// shared memory.  A 40 k-edge row of the power-law graph costs 1257 dependent chunk walks in the row-per-group
// kernels (~20 ms on one warp, longer than the rest of the grid needs for the other 2.4 M rows); here it is 157 per warp.
template <int NV, bool ROW16 = false>
__global__ void __launch_bounds__(256) edge_fwd_hub_kernel(const EdgeFwdParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int U = NV >= 4 ? 2 : (NV == 1 ? 8 : 4);
  __shared__ float sm_m[8], sm_l[8];
  __shared__ float4 sm_acc[8][32 * NV];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int H = p.H, Cp = p.Cp, Q = p.Cp >> 2;
  const int64_t Dp = p.Dp;
  const uint32_t row_bytes = static_cast<uint32_t>(p.Dp) * (ROW16 ? 2u : 4u);
  const float slope = p.slope;
  const int act = p.act;
  int off[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) off[v] = 4 * ((lane + v * 32 < Q) ? lane + v * 32 : Q - 1);
  const DropoutKey dkey = dropout_key(p.drop);
  float amax = 0.f;
  for (int64_t item = blockIdx.x; item < p.nhub * H; item += gridDim.x) {
    const int64_t i = __ldg(p.hub + item / H);
    const int h = static_cast<int>(item % H);
    const int beg = __ldg(p.rowptr + i), end = __ldg(p.rowptr + i + 1);
    if (end - beg > B200GAT_GIANT_DEGREE) continue;       // CTA-uniform: split into segments by the giant kernels below
    const float sd = __ldg(p.s_dst + i * H + h);
    const float* ssrc_h = p.s_src + h;
    const char* wb[NV];
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      wb[v] = row_base<ROW16>(p.wh, p.wh16, h * Cp + off[v]);
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = beg + w * 32; k0 < end; k0 += 256) {
      const int k = k0 + lane;
      const bool ok = k < end;
      int j = static_cast<int>(i);
      float e = -INFINITY;
      if (ok) {
        j = __ldg(p.col + k);
        e = edge_logit<true>(p.s_dst + i * H, p.s_src + int64_t(j) * H, sd, h, H, slope, act);
      }
      const float m_new = fmaxf(m, group_max<32>(e));
      if (m_new != m) {                                   // exp(-inf) = 0 covers the first chunk
        const float scale = expf(m - m_new);
        l *= scale;
#pragma unroll
        for (int v = 0; v < NV; ++v) { acc[v].x *= scale; acc[v].y *= scale; acc[v].z *= scale; acc[v].w *= scale; }
      }
      m = m_new;
      float pp = 0.f, pm = 0.f;
      if (ok) {
        pp = expf(e - m);
        pm = p.drop.active() ? pp * dropout_mult(p.drop, dkey, __ldg(p.eid + k), h, H) : pp;
      }
      l += pp;
      const int cnt = (end - k0) < 32 ? (end - k0) : 32;
      for (int t = 0; t < cnt; t += U) {                  // lanes past the chunk's edges carry weight 0 and a valid row id
        int jt[U];
        float pt[U];
        float4 wv[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          jt[u] = __shfl_sync(FULL, j, t + u);
          pt[u] = __shfl_sync(FULL, pm, t + u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int v = 0; v < NV; ++v) wv[u][v] = gather4<ROW16>(wb[v], jt[u], row_bytes);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(pt[u], wv[u][v].x, acc[v].x);
            acc[v].y = fmaf(pt[u], wv[u][v].y, acc[v].y);
            acc[v].z = fmaf(pt[u], wv[u][v].z, acc[v].z);
            acc[v].w = fmaf(pt[u], wv[u][v].w, acc[v].w);
          }
        }
      }
    }
    l = group_sum<32>(l);
    if (lane == 0) { sm_m[w] = m; sm_l[w] = l; }
#pragma unroll
    for (int v = 0; v < NV; ++v) sm_acc[w][lane + 32 * v] = acc[v];
    __syncthreads();
    float M = sm_m[0];
#pragma unroll
    for (int w2 = 1; w2 < 8; ++w2) M = fmaxf(M, sm_m[w2]);
    float L = 0.f, sc[8];
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) {
      sc[w2] = expf(sm_m[w2] - M);                        // warps that saw no chunk: exp(-inf) = 0
      L = fmaf(sm_l[w2], sc[w2], L);
    }
    const float inv = 1.f / (L + 1e-16f);
    const int q = threadIdx.x;
    if (q < Q) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) {
        const float4 a = sm_acc[w2][q];
        o.x = fmaf(a.x, sc[w2], o.x); o.y = fmaf(a.y, sc[w2], o.y); o.z = fmaf(a.z, sc[w2], o.z); o.w = fmaf(a.w, sc[w2], o.w);
      }
      o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
      if (p.heads_mode) {
        *reinterpret_cast<float4*>(p.o_heads + i * Dp + h * Cp + 4 * q) = o;
      } else {
        const int c = 4 * q;
        const float* b = p.bias + h * p.C + c;
        float* dst = p.out + i * p.ldo + h * p.C + c;
        if (p.vec_out) {
          const float4 bb = ldg4(b);
          const float4 r = make_float4(o.x + bb.x, o.y + bb.y, o.z + bb.z, o.w + bb.w);
          *reinterpret_cast<float4*>(dst) = r;
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))));
        } else {
          const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (c + u < p.C) {
              const float r = ov[u] + __ldg(b + u);
              dst[u] = r;
              amax = fmaxf(amax, fabsf(r));
            }
        }
      }
    }
    if (threadIdx.x == 0) {
      p.rowmax[i * H + h] = M;
      p.rowsum[i * H + h] = L;
    }
    __syncthreads();
  }
  if (p.out_amax && !p.heads_mode) warp_atomic_amax(p.out_amax, amax);
}

static int launch_edge_fwd_giant(const EdgeFwdParams& p, cudaStream_t stream);
static int launch_edge_fwd_hub(const EdgeFwdParams& p, cudaStream_t stream) {
  if (!p.hub || p.nhub <= 0) return 0;
  const int64_t want = p.nhub * p.H;
  const int64_t cap = int64_t(sm_count()) * 8;
  const int blocks = static_cast<int>(want < cap ? want : cap);
  const int Q = p.Cp / 4;
  if (p.wh16) {
    if (Q <= 32) edge_fwd_hub_kernel<1, true><<<blocks, 256, 0, stream>>>(p);
    else if (Q <= 64) edge_fwd_hub_kernel<2, true><<<blocks, 256, 0, stream>>>(p);
    else edge_fwd_hub_kernel<4, true><<<blocks, 256, 0, stream>>>(p);
  }
  else if (Q <= 32) edge_fwd_hub_kernel<1><<<blocks, 256, 0, stream>>>(p);
  else if (Q <= 64) edge_fwd_hub_kernel<2><<<blocks, 256, 0, stream>>>(p);
  else edge_fwd_hub_kernel<4><<<blocks, 256, 0, stream>>>(p);
  const int rc = check_launch("edge_fwd_hub_kernel");
  return rc ? rc : launch_edge_fwd_giant(p, stream);
}

// ---- giant rows (in-degree > B200GAT_GIANT_DEGREE): even one CTA per (row, head) would leave a 40 k-edge row as the
// tail of the launch, so these rows are cut into segments of B200GAT_GIANT_DEGREE edges, one CTA per (row, head, segment)
// (grid.y = segment; CTAs without work exit after three loads), in three steps: (1) row max of the logits by atomicMax,
// (2) exp / row sum / weighted aggregate per segment, added atomically into rowsum and the output row, (3) normalise
// + bias.  The only rows whose summation order is not fixed.
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// -> true when this CTA has a segment [kbeg, kend) of giant row i, head h
__device__ __forceinline__ bool giant_unit(const int32_t* hub, const int32_t* ptr, int H, int64_t& i, int& h, int& kbeg, int& kend) {
  i = __ldg(hub + blockIdx.x / H);
  h = static_cast<int>(blockIdx.x % H);
  const int beg = __ldg(ptr + i), end = __ldg(ptr + i + 1);
  kbeg = beg + static_cast<int>(blockIdx.y) * B200GAT_GIANT_DEGREE;
  kend = kbeg + B200GAT_GIANT_DEGREE < end ? kbeg + B200GAT_GIANT_DEGREE : end;
  return end - beg > B200GAT_GIANT_DEGREE && kbeg < end;
}

__global__ void __launch_bounds__(256) edge_fwd_giant_max_kernel(const EdgeFwdParams p) {
  __shared__ float sm_m[8];
  int64_t i; int h, kbeg, kend;
  if (!giant_unit(p.hub, p.rowptr, p.H, i, h, kbeg, kend)) return;
  const int H = p.H, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float sd = __ldg(p.s_dst + i * H + h);
  float m = -INFINITY;
  for (int k = kbeg + threadIdx.x; k < kend; k += 256)
    m = fmaxf(m, edge_logit<true>(p.s_dst + i * H, p.s_src + int64_t(__ldg(p.col + k)) * H, sd, h, H, p.slope, p.act));
  m = group_max<32>(m);
  if (lane == 0) sm_m[w] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w2 = 1; w2 < 8; ++w2) m = fmaxf(m, sm_m[w2]);
    atomic_max_float(p.rowmax + i * H + h, m);            // placeholder written by the row-per-group kernel: -inf
  }
  if (blockIdx.y == 0) {                                  // segment 0 clears the accumulation targets of step (2)
    if (threadIdx.x == 0) p.rowsum[i * H + h] = 0.f;
    if (p.heads_mode) {
      for (int c = threadIdx.x; c < p.Cp; c += 256) p.o_heads[i * p.Dp + h * p.Cp + c] = 0.f;
    } else {
      for (int c = threadIdx.x; c < p.C; c += 256) p.out[i * p.ldo + h * p.C + c] = 0.f;
    }
  }
}

template <int NV, bool ROW16 = false>
__global__ void __launch_bounds__(256) edge_fwd_giant_acc_kernel(const EdgeFwdParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int U = NV >= 4 ? 2 : (NV == 1 ? 8 : 4);
  __shared__ float sm_l[8];
  __shared__ float4 sm_acc[8][32 * NV];
  int64_t i; int h, kbeg, kend;
  if (!giant_unit(p.hub, p.rowptr, p.H, i, h, kbeg, kend)) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int H = p.H, Cp = p.Cp, Q = p.Cp >> 2;
  const uint32_t row_bytes = static_cast<uint32_t>(p.Dp) * (ROW16 ? 2u : 4u);
  const float sd = __ldg(p.s_dst + i * H + h);
  const float M = p.rowmax[i * H + h];                    // final: written by edge_fwd_giant_max_kernel
  const DropoutKey dkey = dropout_key(p.drop);
  const char* wb[NV];
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    wb[v] = row_base<ROW16>(p.wh, p.wh16, h * Cp + 4 * ((lane + v * 32 < Q) ? lane + v * 32 : Q - 1));
  }
  float l = 0.f;
  for (int k0 = kbeg + w * 32; k0 < kend; k0 += 256) {
    const int k = k0 + lane;
    const bool ok = k < kend;
    int j = static_cast<int>(i);
    float pp = 0.f, pm = 0.f;
    if (ok) {
      j = __ldg(p.col + k);
      pp = expf(edge_logit<true>(p.s_dst + i * H, p.s_src + int64_t(j) * H, sd, h, H, p.slope, p.act) - M);
      pm = p.drop.active() ? pp * dropout_mult(p.drop, dkey, __ldg(p.eid + k), h, H) : pp;
    }
    l += pp;
    const int cnt = (kend - k0) < 32 ? (kend - k0) : 32;
    for (int t = 0; t < cnt; t += U) {
      int jt[U];
      float pt[U];
      float4 wv[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        jt[u] = __shfl_sync(FULL, j, t + u);
        pt[u] = __shfl_sync(FULL, pm, t + u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < NV; ++v) wv[u][v] = gather4<ROW16>(wb[v], jt[u], row_bytes);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          acc[v].x = fmaf(pt[u], wv[u][v].x, acc[v].x);
          acc[v].y = fmaf(pt[u], wv[u][v].y, acc[v].y);
          acc[v].z = fmaf(pt[u], wv[u][v].z, acc[v].z);
          acc[v].w = fmaf(pt[u], wv[u][v].w, acc[v].w);
        }
      }
    }
  }
  l = group_sum<32>(l);
  if (lane == 0) sm_l[w] = l;
#pragma unroll
  for (int v = 0; v < NV; ++v) sm_acc[w][lane + 32 * v] = acc[v];
  __syncthreads();
  const int q = threadIdx.x;
  if (q < Q) {
    float4 o = sm_acc[0][q];
#pragma unroll
    for (int w2 = 1; w2 < 8; ++w2) {
      const float4 a = sm_acc[w2][q];
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    const float ov[4] = {o.x, o.y, o.z, o.w};
    float* dst = p.heads_mode ? p.o_heads + i * p.Dp + h * Cp + 4 * q : p.out + i * p.ldo + h * p.C + 4 * q;
    const int lim = p.heads_mode ? Cp : p.C;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (4 * q + u < lim) atomicAdd(dst + u, ov[u]);
  }
  if (threadIdx.x == 0) {
    float L = sm_l[0];
#pragma unroll
    for (int w2 = 1; w2 < 8; ++w2) L += sm_l[w2];
    atomicAdd(p.rowsum + i * H + h, L);
  }
}

__global__ void __launch_bounds__(128) edge_fwd_giant_finish_kernel(const EdgeFwdParams p) {
  const int H = p.H;
  const int64_t i = __ldg(p.hub + blockIdx.x / H);
  const int h = static_cast<int>(blockIdx.x % H);
  float amax = 0.f;
  if (__ldg(p.rowptr + i + 1) - __ldg(p.rowptr + i) > B200GAT_GIANT_DEGREE) {
    const float inv = 1.f / (p.rowsum[i * H + h] + 1e-16f);
    if (p.heads_mode) {
      for (int c = threadIdx.x; c < p.Cp; c += 128) p.o_heads[i * p.Dp + h * p.Cp + c] *= inv;
    } else {
      for (int c = threadIdx.x; c < p.C; c += 128) {
        float* dst = p.out + i * p.ldo + h * p.C + c;
        const float r = *dst * inv + __ldg(p.bias + h * p.C + c);
        *dst = r;
        amax = fmaxf(amax, fabsf(r));
      }
    }
  }
  if (p.out_amax && !p.heads_mode) warp_atomic_amax(p.out_amax, amax);
}

static int launch_edge_fwd_giant(const EdgeFwdParams& p, cudaStream_t stream) {
  if (!p.hub || p.nhub <= 0 || p.max_deg <= B200GAT_GIANT_DEGREE) return 0;
  const int64_t nseg = ceil_div(p.max_deg, B200GAT_GIANT_DEGREE);
  B200GAT_REQUIRE(nseg <= 65535, B200GAT_E_UNSUPPORTED, "edge_fwd: a row of %d edges is not supported", p.max_deg);
  const dim3 grid(static_cast<unsigned>(p.nhub * p.H), static_cast<unsigned>(nseg));
  edge_fwd_giant_max_kernel<<<grid, 256, 0, stream>>>(p);
  int rc = check_launch("edge_fwd_giant_max_kernel");
  if (rc) return rc;
  const int Q = p.Cp / 4;
  if (p.wh16) {
    if (Q <= 32) edge_fwd_giant_acc_kernel<1, true><<<grid, 256, 0, stream>>>(p);
    else if (Q <= 64) edge_fwd_giant_acc_kernel<2, true><<<grid, 256, 0, stream>>>(p);
    else edge_fwd_giant_acc_kernel<4, true><<<grid, 256, 0, stream>>>(p);
  }
  else if (Q <= 32) edge_fwd_giant_acc_kernel<1><<<grid, 256, 0, stream>>>(p);
  else if (Q <= 64) edge_fwd_giant_acc_kernel<2><<<grid, 256, 0, stream>>>(p);
  else edge_fwd_giant_acc_kernel<4><<<grid, 256, 0, stream>>>(p);
  if ((rc = check_launch("edge_fwd_giant_acc_kernel"))) return rc;
  edge_fwd_giant_finish_kernel<<<static_cast<unsigned>(p.nhub * p.H), 128, 0, stream>>>(p);
  return check_launch("edge_fwd_giant_finish_kernel");
}

// ---- row-wide schedule for NARROW heads (head width < 128 channels, 2..8 heads, at most 512 padded channels per row):
// one lane group per destination ROW covers all heads — lane slots run over the whole padded Wh row [H, Cp], so a
// gathered row is one contiguous H*Cp*4-byte read and the col[] -> s_src -> exp -> gather dependent chain is paid once
// per chunk of G edges instead of once per (chunk, head) with chunks of only Cp/4 edges.  The edge-parallel phase
// computes the H softmax weights of each edge and parks them (and the edge's source id) in shared memory; in the
// feature-parallel phase every lane reads the weight of ITS slot's head.  Used for streaming graphs (rows gathered
// from HBM), where the longer contiguous reads and the 4x larger chunks pay.
template <int G, int NV, int HH, bool HAS_MASK, bool ROW16 = false>
__device__ __forceinline__ void edge_fwd_row_body(const EdgeFwdParams& p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int GPW = 32 / G;
  constexpr int U = NV >= 4 ? 2 : (NV == 1 ? 8 : 4);
  __shared__ float ps_all[8][32 * HH];
  __shared__ int js_all[8][32];
  __shared__ float sc_all[8][GPW * HH];
  const int wid = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
  float* ps = ps_all[wid] + gi * G * HH;
  int* js = js_all[wid] + gi * G;
  float* sc = sc_all[wid] + gi * HH;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int Q = p.Cp >> 2, S = HH * Q;
  const int64_t Dp = p.Dp;
  const uint32_t row_bytes = static_cast<uint32_t>(p.Dp) * (ROW16 ? 2u : 4u);
  const float slope = p.slope;
  int off[NV], hv[NV];
  bool live[NV];
  const char* wb[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int s = gl + v * G;
    live[v] = s < S;
    const int sl = live[v] ? s : S - 1;                   // dead slots gather a clamped (valid) column and are never stored
    off[v] = 4 * sl;
    hv[v] = sl / Q;
    wb[v] = row_base<ROW16>(p.wh, p.wh16, off[v]);
  }

  const DropoutKey dkey = HAS_MASK ? dropout_key(p.drop) : DropoutKey{0u, 0u, 0u, 0u};
  float amax = 0.f;
  for (int64_t base = warp * GPW; base < p.N; base += nwarps * GPW) {
    const bool valid = base + gi < p.N;
    const int64_t i = valid ? base + gi : p.N - 1;
    const int beg = valid ? __ldg(p.rowptr + i) : 0;
    const int end = valid ? __ldg(p.rowend + i) : 0;            // hub rows are empty here (edge_fwd_hub_kernel)
    const int deg = end - beg;
    const int maxdeg = GPW == 1 ? deg : __reduce_max_sync(FULL, deg);
    float sd[HH], m[HH], l[HH];
#pragma unroll
    for (int h = 0; h < HH; ++h) {
      sd[h] = __ldg(p.s_dst + i * HH + h);
      m[h] = -INFINITY;
      l[h] = 0.f;
    }
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int k0 = 0; k0 < maxdeg; k0 += G) {
      const int k = beg + k0 + gl;
      const bool ok = k < end;
      int j = static_cast<int>(i);
      float e[HH];
#pragma unroll
      for (int h = 0; h < HH; ++h) e[h] = -INFINITY;
      int eid_k = 0;
      if (ok) {
        j = __ldg(p.col + k);
        const float* sj = p.s_src + int64_t(j) * HH;
#pragma unroll
        for (int h = 0; h < HH; ++h) e[h] = leaky(sd[h] + __ldg(sj + h), slope);
        if (HAS_MASK) eid_k = __ldg(p.eid + k);
      }
      js[gl] = j;
#pragma unroll
      for (int h = 0; h < HH; ++h) {
        const float m_new = fmaxf(m[h], group_max<G>(e[h]));
        if (k0 > 0) {                                     // group-uniform; exp(-inf - x) = 0 covers "nothing accumulated yet"
          const float scale = m_new != m[h] ? expf(m[h] - m_new) : 1.f;
          l[h] *= scale;
          if (gl == 0) sc[h] = scale;
        }
        m[h] = m_new;
        float pp = 0.f, pm = 0.f;
        if (ok) {
          pp = expf(e[h] - m_new);
          pm = HAS_MASK ? pp * dropout_mult(p.drop, dkey, eid_k, h, HH) : pp;
        }
        l[h] += pp;
        ps[gl * HH + h] = pm;
      }
      __syncwarp();
      if (k0 > 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float scale = sc[hv[v]];
          acc[v].x *= scale; acc[v].y *= scale; acc[v].z *= scale; acc[v].w *= scale;
        }
      }
      const int cnt = (maxdeg - k0) < G ? (maxdeg - k0) : G;
      // slots t + u past the chunk's edge count (always < G) hold weight 0 and the row's own (valid) id
      for (int t = 0; t < cnt; t += U) {
        float pt[U][NV];
        float4 w[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jt = js[t + u];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            pt[u][v] = ps[(t + u) * HH + hv[v]];
            if (HAS_MASK) w[u][v] = pt[u][v] != 0.f ? gather4<ROW16>(wb[v], jt, row_bytes) : make_float4(0.f, 0.f, 0.f, 0.f);
            else w[u][v] = gather4<ROW16>(wb[v], jt, row_bytes);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(pt[u][v], w[u][v].x, acc[v].x);
            acc[v].y = fmaf(pt[u][v], w[u][v].y, acc[v].y);
            acc[v].z = fmaf(pt[u][v], w[u][v].z, acc[v].z);
            acc[v].w = fmaf(pt[u][v], w[u][v].w, acc[v].w);
          }
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int h = 0; h < HH; ++h) {
      l[h] = group_sum<G>(l[h]);
      if (gl == 0) {
        sc[h] = 1.f / (l[h] + 1e-16f);
        if (valid) {
          p.rowmax[i * HH + h] = m[h];
          p.rowsum[i * HH + h] = l[h];
        }
      }
    }
    __syncwarp();
    if (valid) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!live[v]) continue;
        const float inv = sc[hv[v]];
        const float4 o = make_float4(acc[v].x * inv, acc[v].y * inv, acc[v].z * inv, acc[v].w * inv);
        if (p.heads_mode) {
          *reinterpret_cast<float4*>(p.o_heads + i * Dp + off[v]) = o;
        } else {
          const int c = off[v] - hv[v] * p.Cp;
          const float* b = p.bias + hv[v] * p.C + c;
          float* dst = p.out + i * p.ldo + hv[v] * p.C + c;
          if (p.vec_out) {
            const float4 bb = ldg4(b);
            const float4 r = make_float4(o.x + bb.x, o.y + bb.y, o.z + bb.z, o.w + bb.w);
            *reinterpret_cast<float4*>(dst) = r;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))));
          } else {
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (c + u < p.C) {
                const float r = ov[u] + __ldg(b + u);
                dst[u] = r;
                amax = fmaxf(amax, fabsf(r));
              }
          }
        }
      }
    }
    __syncwarp();
  }
  if (p.out_amax && !p.heads_mode) warp_atomic_amax(p.out_amax, amax);
}

template <int G, int NV, int HH, bool HAS_MASK>
__global__ void __launch_bounds__(256) edge_fwd_row_kernel(const EdgeFwdParams p) { edge_fwd_row_body<G, NV, HH, HAS_MASK>(p); }
template <int G, int NV, int HH>
__global__ void __launch_bounds__(256, 3) edge_fwd_row_stream_kernel(const EdgeFwdParams p) { edge_fwd_row_body<G, NV, HH, false>(p); }
template <int G, int NV, int HH>
__global__ void __launch_bounds__(256, 3) edge_fwd16_row_stream_kernel(const EdgeFwdParams p) { edge_fwd_row_body<G, NV, HH, false, true>(p); }

template <int G, int NV, int HH>
static int launch_edge_fwd_row(const EdgeFwdParams& p, bool streaming, cudaStream_t stream) {
  constexpr int GPW = 32 / G;
  const int threads = 256;
  const int64_t want = ceil_div(ceil_div(p.N, GPW), threads / 32);
  const int64_t cap = int64_t(sm_count()) * 8;
  const int blocks = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  (void)streaming;                                        // only dispatched for streaming graphs (see b200gat_edge_fwd)
  if (p.wh16) edge_fwd16_row_stream_kernel<G, NV, HH><<<blocks, threads, 0, stream>>>(p);
  else if (p.drop.active()) edge_fwd_row_kernel<G, NV, HH, true><<<blocks, threads, 0, stream>>>(p);
  else edge_fwd_row_stream_kernel<G, NV, HH><<<blocks, threads, 0, stream>>>(p);
  return check_launch("edge_fwd_row_kernel");
}

template <int HH>
static int dispatch_edge_fwd_row(const EdgeFwdParams& p, int S, bool streaming, cudaStream_t stream) {
  if (S <= 16) return launch_edge_fwd_row<16, 1, HH>(p, streaming, stream);
  if (S <= 32) return launch_edge_fwd_row<32, 1, HH>(p, streaming, stream);
  if (S <= 64) return launch_edge_fwd_row<32, 2, HH>(p, streaming, stream);
  return launch_edge_fwd_row<32, 4, HH>(p, streaming, stream);
}

// narrow heads: 2..8 heads (compile-time instantiations), head width < 128 channels, row <= 512 padded channels
static bool edge_fwd_row_supported(int H, int Q, int act) {
  return act == B200GAT_LOGIT_LEAKY_RELU && Q < 32 && H * Q <= 128 && (H == 2 || H == 3 || H == 4 || H == 6 || H == 8);
}

}  // namespace b200gat

using namespace b200gat;

extern "C" int b200gat_edge_fwd(const b200gat_edge_fwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "edge_fwd: NULL args");
  int rc = validate_layer(a->layer);
  if (rc) return rc;
  if ((rc = validate_graph(a->graph))) return rc;
  const b200gat_layer& L = a->layer;
  const int64_t N = a->graph.num_nodes;
  if (N == 0) return 0;
  const int H = static_cast<int>(L.heads), C = static_cast<int>(L.out_channels), Cp = static_cast<int>(L.c_pad);
  const bool heads_mode = !L.concat && H > 1;
  const int64_t d_out = L.concat ? int64_t(H) * C : C;
  B200GAT_REQUIRE(a->wh && a->s_src && a->s_dst && a->bias && a->out && a->rowmax && a->rowsum, B200GAT_E_NULL,
                  "edge_fwd: NULL pointer");
  B200GAT_REQUIRE(!heads_mode || a->o_heads, B200GAT_E_NULL, "edge_fwd: o_heads is required when !concat && heads > 1");
  DropoutSpec drop;
  if ((rc = make_dropout(a->mask, a->dropout, &drop))) return rc;
  B200GAT_REQUIRE(!drop.active() || a->graph.eid, B200GAT_E_NULL, "edge_fwd: dropout needs graph.eid");
  B200GAT_REQUIRE(a->ldo >= d_out, B200GAT_E_SHAPE, "edge_fwd: ldo < D_out");
  B200GAT_REQUIRE(aligned16(a->wh) && (!heads_mode || aligned16(a->o_heads)), B200GAT_E_ALIGN,
                  "edge_fwd: wh / o_heads must be 16-byte aligned");
  B200GAT_REQUIRE(N * int64_t(H) * Cp < (int64_t(1) << 40), B200GAT_E_SHAPE, "edge_fwd: N * Dp too large");

  EdgeFwdParams p;
  p.N = N; p.items = N * H;
  p.H = H; p.C = C; p.Cp = Cp; p.Dp = H * Cp;
  p.slope = L.negative_slope; p.act = L.logit_activation;
  p.rowptr = a->graph.rowptr; p.col = a->graph.col; p.eid = a->graph.eid;
  p.wh = a->wh; p.s_src = a->s_src; p.s_dst = a->s_dst; p.bias = a->bias; p.drop = drop;
  // bf16-stored gathered rows: only the plain configuration (LeakyReLU, no dropout) has ROW16 instantiations
  B200GAT_REQUIRE(!a->wh_bf16 || (!drop.active() && L.logit_activation == B200GAT_LOGIT_LEAKY_RELU), B200GAT_E_UNSUPPORTED,
                  "edge_fwd: wh_bf16 is offered without dropout and with LeakyReLU logits only");
  B200GAT_REQUIRE(!a->wh_bf16 || aligned16(a->wh_bf16), B200GAT_E_ALIGN, "edge_fwd: wh_bf16 must be 16-byte aligned");
  p.wh16 = a->wh_bf16;
  p.out = a->out; p.ldo = a->ldo; p.rowmax = a->rowmax; p.rowsum = a->rowsum; p.o_heads = a->o_heads;
  p.heads_mode = heads_mode ? 1 : 0;
  p.vec_out = (!heads_mode && C % 4 == 0 && a->ldo % 4 == 0 && aligned16(a->out) && aligned16(a->bias)) ? 1 : 0;
  p.out_amax = a->out_amax;
  B200GAT_REQUIRE(a->graph.num_hub_rows >= 0 && (a->graph.num_hub_rows == 0 || a->graph.hub_rows), B200GAT_E_NULL,
                  "edge_fwd: graph.hub_rows missing");
  B200GAT_REQUIRE(a->graph.num_hub_rows == 0 || a->graph.rowend, B200GAT_E_NULL, "edge_fwd: graph.rowend missing");
  p.hub = a->graph.hub_rows; p.nhub = a->graph.num_hub_rows;
  p.rowend = p.nhub > 0 ? a->graph.rowend : a->graph.rowptr + 1;
  p.max_deg = static_cast<int>(a->graph.max_in_degree);
  if (a->out_amax) {
    cudaError_t ce = cudaMemsetAsync(a->out_amax, 0, sizeof(uint32_t), stream);
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "edge_fwd: memset: %s", cudaGetErrorString(ce));
  }

  const int Q = Cp / 4;
  const bool streaming = edge_schedule_streaming(a->graph.span, int64_t(H) * Cp * 4);
  // streaming graphs only — measured on the L2-resident PPI-shaped batch (heads sweep, 50 -> H x 64) the per-(row, head)
  // schedule is faster (H = 4: 112 vs 141 us, H = 8: 221 vs 366 us); on the 2.4 M-node graph's 4 x 47 layer the row-wide
  // one wins (12.8 vs 15.6 ms)
  if (streaming && edge_fwd_row_supported(H, Q, p.act)) {
    switch (H) {
      case 2: rc = dispatch_edge_fwd_row<2>(p, H * Q, streaming, stream); break;
      case 3: rc = dispatch_edge_fwd_row<3>(p, H * Q, streaming, stream); break;
      case 4: rc = dispatch_edge_fwd_row<4>(p, H * Q, streaming, stream); break;
      case 6: rc = dispatch_edge_fwd_row<6>(p, H * Q, streaming, stream); break;
      default: rc = dispatch_edge_fwd_row<8>(p, H * Q, streaming, stream); break;
    }
  } else if (Q <= 1) rc = launch_edge_fwd<1, 1>(p, streaming, stream);
  else if (Q <= 2) rc = launch_edge_fwd<2, 1>(p, streaming, stream);
  else if (Q <= 4) rc = launch_edge_fwd<4, 1>(p, streaming, stream);
  else if (Q <= 8) rc = launch_edge_fwd<8, 1>(p, streaming, stream);
  else if (Q <= 16) rc = launch_edge_fwd<16, 1>(p, streaming, stream);
  else if (Q <= 32) rc = launch_edge_fwd<32, 1>(p, streaming, stream);
  else if (Q <= 64) rc = launch_edge_fwd<32, 2>(p, streaming, stream);
  else rc = launch_edge_fwd<32, 4>(p, streaming, stream);
  if (rc) return rc;
  if ((rc = launch_edge_fwd_hub(p, stream))) return rc;
  if (heads_mode) {
    const int64_t total = N * (Cp / 4);
    const int64_t want = ceil_div(total, 256);
    const int64_t cap = int64_t(sm_count()) * 8;
    head_mean_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, stream>>>(a->o_heads, a->bias, a->out, a->ldo,
                                                                                   N, H, C, Cp, a->out_amax);
    rc = check_launch("head_mean_kernel");
  }
  return rc;
}

// ---- the in-kernel dropout multipliers as a tensor (include/b200gat.h: b200gat_dropout_mask) ----
namespace b200gat {
__global__ void __launch_bounds__(256) dropout_mask_kernel(const DropoutSpec d, int64_t total, int H, float* __restrict__ out) {
  const DropoutKey key = dropout_key(d);
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = t / H;
    out[t] = dropout_mult(d, key, static_cast<int>(e), static_cast<int>(t - e * H), H);
  }
}
}  // namespace b200gat

extern "C" int b200gat_dropout_mask(const b200gat_dropout* d, int64_t num_edges, int64_t heads, float* mask, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(d && mask, B200GAT_E_NULL, "dropout_mask: NULL pointer");
  B200GAT_REQUIRE(num_edges >= 0 && heads > 0 && num_edges < (int64_t(1) << 31), B200GAT_E_SHAPE, "dropout_mask: bad sizes");
  DropoutSpec spec;
  int rc = make_dropout(nullptr, *d, &spec);
  if (rc) return rc;
  B200GAT_REQUIRE(spec.active(), B200GAT_E_SHAPE, "dropout_mask: p must be > 0");
  const int64_t total = num_edges * heads;
  if (total == 0) return 0;
  const int64_t want = ceil_div(total, 256), cap = int64_t(sm_count()) * 8;
  dropout_mask_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, stream>>>(spec, total, static_cast<int>(heads), mask);
  return check_launch("dropout_mask_kernel");
}
