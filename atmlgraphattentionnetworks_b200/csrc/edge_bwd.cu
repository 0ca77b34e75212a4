// K3 — fused edge backward.  The reference has no backward code: loss.backward() (run_inductive.py:84) differentiates
// GAT.py:53-67 through saved per-edge tensors.  Here nothing per-edge was saved; alpha is recomputed from
// s_src, s_dst and the forward's rowmax / rowsum, and (closed form, SURVEY.md §3D, checked by oracle/closed_form.py)
//     G[i,h,:]   = concat ? gout[i, hC:(h+1)C] : gout[i,:] / H
//     Drow[i,h]  = <G[i,h,:], O[i,h,:]>                       (= sum_k alpha~_k dalpha_k: no second pass over edges)
//     dalpha_k   = mask_k <G[i,h,:], Wh[j,h,:]>
//     dz_k       = alpha_k (dalpha_k - Drow[i,h]) * (z_k > 0 ? 1 : slope)
//     gWh[j,h,:] = sum_{k in out(j)} alpha_k mask_k G[i_k,h,:]     g_s_src[j,h] = sum_{out(j)} dz     g_s_dst[i,h] = sum_{in(i)} dz
//     gT = gWh + g_s_src (x) a1 + g_s_dst (x) a2 ;  g_bw = colsum gT ; g_a1 = colsum g_s_src*Wh ; g_a2 = colsum g_s_dst*Wh
// Passes:  bwd_prep_kernel (stream: Drow, and the per-(node, head) record {s_dst, rowmax, 1/(rowsum+1e-16), Drow} packed
// into one 16-byte gatherable word)  ->  edge_bwd_kernel (CSC: one lane group per (source row j, head h); Wh[j,h,:]
// stays in registers, G[i] rows are gathered 128 bits per lane; gWh / g_s_src are written without atomics, g_s_dst is
// the one cross-orientation reduction: [N,H] float atomics)  ->  bwd_finish_kernel (stream: gT in place + column sums).
// CSC-pass variants: edge_bwd_mean_kernel (mean-over-heads layers: one lane group per source row,
// G[i] gathered once for all heads), edge_bwd_hub_kernel (out-degree > 512: one CTA per (source row, head); > 4096: one
// CTA per 4096-edge segment, partial sums added atomically), edge_bwd_act_kernel (LogSigmoid / Tanh logits).
#include "common.cuh"
#include "split_blob.cuh"
#include "proj_tc.cuh"
#include <math.h>

namespace b200gat {

// ---- Drow + row record, and (when the upstream gradient is not directly gatherable) a padded copy of G ------------
struct PrepParams {
  int64_t N;
  int H, C, Cp;
  int concat_like;           // concat || H == 1
  int vec;                   // gout / out rows and head offsets are 16-byte aligned (C % 4 == 0 etc.)
  const float* gout; int64_t ldgo;
  const float* out; int64_t ldo;      // concat_like: O = out - bias
  const float* o_heads;               // otherwise
  const float* bias;
  const float* s_dst; const float* rowmax; const float* rowsum;
  float gscale;                       // 1 or 1/H
  int act;                            // 1: gout is d/d ELU(out) (concat_like only): G = gout * ELU'(out), needs gp
  float* gp; int64_t ldgp;            // optional padded copy: concat_like ? [N, H*Cp] : [N, Cp]
  int gp16;                           // the copy is stored as bf16 (same element layout, 2-byte elements)
  float4* rowrec;                     // [N, H] {s_dst, rowmax, 1/(rowsum + 1e-16), Drow}
};

__global__ void __launch_bounds__(256) bwd_prep_kernel(const PrepParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int64_t items = p.N * p.H;
  for (int64_t item = warp; item < items; item += nwarps) {
    const int64_t i = item / p.H;
    const int h = static_cast<int>(item - i * p.H);
    float d = 0.f;
    if (p.vec) {   // concat_like, no padded copy needed: 128-bit streaming loads
      const float* g = p.gout + i * p.ldgo + h * p.C;
      const float* o = p.out + i * p.ldo + h * p.C;
      const float* b = p.bias + h * p.C;
      for (int c = 4 * lane; c < p.C; c += 128) {
        const float4 gv = ldg4(g + c), ov = ldg4(o + c), bv = ldg4(b + c);
        d = fmaf(gv.x, ov.x - bv.x, d);
        d = fmaf(gv.y, ov.y - bv.y, d);
        d = fmaf(gv.z, ov.z - bv.z, d);
        d = fmaf(gv.w, ov.w - bv.w, d);
      }
    } else {
      const float* g = p.gout + i * p.ldgo + (p.concat_like ? h * p.C : 0);
      for (int c = lane; c < p.Cp; c += 32) {
        float gv = 0.f;
        if (c < p.C) {
          gv = __ldg(g + c) * p.gscale;
          float o;
          if (p.concat_like) {
            const float pre = __ldg(p.out + i * p.ldo + h * p.C + c);
            if (p.act) gv *= elu_grad(pre);
            o = pre - __ldg(p.bias + h * p.C + c);
          } else {
            o = __ldg(p.o_heads + (i * p.H + h) * int64_t(p.Cp) + c);
          }
          d = fmaf(gv, o, d);
        }
        if (p.gp && (p.concat_like || h == 0)) {
          const int64_t e = i * p.ldgp + (p.concat_like ? h * p.Cp : 0) + c;
          if (p.gp16) reinterpret_cast<uint16_t*>(p.gp)[e] = static_cast<uint16_t>(pack_bf16x4(gv, 0.f, 0.f, 0.f).x & 0xffffu);
          else p.gp[e] = gv;
        }
      }
    }
    d = group_sum<32>(d);
    if (lane == 0)
      p.rowrec[item] = make_float4(__ldg(p.s_dst + item), __ldg(p.rowmax + item), 1.f / (__ldg(p.rowsum + item) + 1e-16f), d);
  }
}

// ---- fast prep: one warp per destination ROW (all heads), for concat-like layers with D_out <= 1024 and a head
// width that is a power of two <= 128 or a multiple of 128 channels.  All of a row's 128-bit loads are in flight
// together; the column sums of G (g_bias) ride along in registers; with an output activation the upstream gradient is
// first multiplied by ELU'(out) and the product is written to the gatherable copy gp. -------------------------------
struct PrepRowsParams {
  int64_t N;
  int H, C, D;                        // D = H * C = D_out
  const float* gout; int64_t ldgo;
  const float* out; int64_t ldo;      // forward output BEFORE the activation (O = out - bias)
  const float* bias;
  const float* s_dst; const float* rowmax; const float* rowsum;
  float* gp;                          // [N, D] (ACT, or gp16: then ALWAYS written, as bf16)
  int gp16;
  float4* rowrec;
  float* g_bias;                      // [D], zero-initialised, accumulated atomically
};

// ---- the same pass laid out like bwd_finish_kernel: one thread per FOUR adjacent columns (TX = pow2ceil(D / 4) column
// groups x RY = 256 / TX row lanes), batches of RB rows per thread with all 2 * RB 128-bit loads issued before the first
// store, row batches interleaved over the CTAs (one moving window of rows, DRAM pages walked in address order).
// ~64 registers -> 4 CTAs / SM (a warp-per-row layout holds a row's 8 float4 x 2 arrays + 8 column-sum float4 per lane:
// 126 registers, 2 CTAs / SM, 37 % issue slots, 4.3 TB/s measured; this one: 157 -> 121 us per PPI layer).  Heads of >= 128 channels span whole warps: warp sums
// go through a double-buffered shared array (ONE barrier per batch) and the first warp writes the row records; narrower
// heads (power-of-two widths) are reduced by segmented shuffles and written by their lead lanes.
template <bool ACT, int TXS>
__global__ void __launch_bounds__(256, 4) bwd_prep_rows_wide_kernel(const PrepRowsParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int TX = 1 << TXS, RY = 256 >> TXS, RB = 4, WPR = TX / 32;    // WPR: warps per row
  const int tx = threadIdx.x & (TX - 1), ry = threadIdx.x >> TXS, lane = threadIdx.x & 31;
  const int Q = p.C >> 2, DQ = p.D >> 2;
  const bool active = tx < DQ;
  const int c = 4 * (active ? tx : 0);
  const int h = c / p.C;
  const float4 bv = ldg4(p.bias + c);
  const bool wide_heads = Q >= 32;                        // (Q % 32 == 0: checked by the host)
  const int per = Q >> 5;                                 // warps per head (wide heads)
  __shared__ float red[2][RB][8];
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  // the writer threads of the wide-head form: thread w < RY * RB * H  <->  (row lane, batch slot, head)
  const int w_h = threadIdx.x % p.H, w_u = (threadIdx.x / p.H) % RB, w_ry = threadIdx.x / (p.H * RB);
  const bool writer = wide_heads && w_ry < RY;
  int buf = 0;
  for (int64_t base = int64_t(blockIdx.x) * (RY * RB); base < p.N; base += int64_t(gridDim.x) * (RY * RB), buf ^= 1) {
    const int64_t rb = base + ry * RB;
    float4 gv[RB], ov[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int64_t r = rb + u < p.N ? rb + u : p.N - 1;    // tail rows re-read the last row and are not stored
      gv[u] = ldg4(p.gout + r * p.ldgo + c);
      ov[u] = ldg4(p.out + r * p.ldo + c);
    }
    float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t w_item = (base + w_ry * RB + w_u) * p.H + w_h;
    const bool w_live = writer && base + w_ry * RB + w_u < p.N;
    if (w_live) {                                           // in flight together with the row loads
      rec.x = __ldg(p.s_dst + w_item); rec.y = __ldg(p.rowmax + w_item); rec.z = __ldg(p.rowsum + w_item);
    }
    float pd[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      pd[u] = 0.f;
      if (!active || rb + u >= p.N) continue;
      const int64_t r = rb + u;
      if (ACT) {
        gv[u].x *= elu_grad(ov[u].x); gv[u].y *= elu_grad(ov[u].y);
        gv[u].z *= elu_grad(ov[u].z); gv[u].w *= elu_grad(ov[u].w);
      }
      if (p.gp16) *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.gp) + r * int64_t(p.D) + c) = pack_bf16x4(gv[u].x, gv[u].y, gv[u].z, gv[u].w);
      else if (ACT) *reinterpret_cast<float4*>(p.gp + r * int64_t(p.D) + c) = gv[u];
      cs.x += gv[u].x; cs.y += gv[u].y; cs.z += gv[u].z; cs.w += gv[u].w;
      pd[u] = gv[u].x * (ov[u].x - bv.x) + gv[u].y * (ov[u].y - bv.y) + gv[u].z * (ov[u].z - bv.z) + gv[u].w * (ov[u].w - bv.w);
    }
    if (wide_heads) {
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) {
#pragma unroll
        for (int u = 0; u < RB; ++u) pd[u] += __shfl_xor_sync(FULL, pd[u], o2);
      }
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < RB; ++u) red[buf][u][threadIdx.x >> 5] = pd[u];
      }
      __syncthreads();
      if (w_live) {
        float d = 0.f;
        for (int k = 0; k < per; ++k) d += red[buf][w_u][w_ry * WPR + w_h * per + k];
        p.rowrec[w_item] = make_float4(rec.x, rec.y, 1.f / (rec.z + 1e-16f), d);
      }
    } else {                                                // Q a power of two < 32: 32 / Q heads per warp
      for (int o2 = Q >> 1; o2 > 0; o2 >>= 1) {
#pragma unroll
        for (int u = 0; u < RB; ++u) pd[u] += __shfl_xor_sync(FULL, pd[u], o2);
      }
      if (active && (lane & (Q - 1)) == 0) {
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          if (rb + u >= p.N) break;
          const int64_t item = (rb + u) * p.H + h;
          p.rowrec[item] = make_float4(__ldg(p.s_dst + item), __ldg(p.rowmax + item), 1.f / (__ldg(p.rowsum + item) + 1e-16f), pd[u]);
        }
      }
    }
  }
  // column sums: the RY row lanes of the CTA through shared memory, then one atomic per column per CTA
  if constexpr (RY > 1) {
    __shared__ float4 csred[RY][TX];
    csred[ry][tx] = cs;
    __syncthreads();
    if (ry != 0) return;
#pragma unroll
    for (int y = 1; y < RY; ++y) { const float4 t = csred[y][tx]; cs.x += t.x; cs.y += t.y; cs.z += t.z; cs.w += t.w; }
  }
  if (!active) return;
  atomicAdd(p.g_bias + c + 0, cs.x); atomicAdd(p.g_bias + c + 1, cs.y);
  atomicAdd(p.g_bias + c + 2, cs.z); atomicAdd(p.g_bias + c + 3, cs.w);
}

// ---- the same for NARROW concat-like layers (D_out <= 128: the 8 x 8 layers of GATNet, the heads sweep's 1 x 64 / 2 x 64):
// one lane group of G = pow2ceil(D / 4) lanes per destination row, 32 / G rows per warp, one float4 slot per lane.
template <bool ACT, int G>
__global__ void __launch_bounds__(256) bwd_prep_rows_narrow_kernel(const PrepRowsParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int Q = p.C >> 2, DQ = p.D >> 2;                  // Q is a power of two <= G
  const bool live = gl < DQ;
  const float4 bv = live ? ldg4(p.bias + 4 * gl) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = warp * RPW; base < p.N; base += nwarps * RPW) {
    const int64_t i = base + gi;
    const bool valid = live && i < p.N;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), ov = gv;
    if (valid) {
      gv = ldg4(p.gout + i * p.ldgo + 4 * gl);
      ov = ldg4(p.out + i * p.ldo + 4 * gl);
      if (ACT) {
        gv.x *= elu_grad(ov.x); gv.y *= elu_grad(ov.y); gv.z *= elu_grad(ov.z); gv.w *= elu_grad(ov.w);
      }
      if (p.gp16) *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.gp) + i * int64_t(p.D) + 4 * gl) = pack_bf16x4(gv.x, gv.y, gv.z, gv.w);
      else if (ACT) *reinterpret_cast<float4*>(p.gp + i * int64_t(p.D) + 4 * gl) = gv;
    }
    cs.x += gv.x; cs.y += gv.y; cs.z += gv.z; cs.w += gv.w;
    float d = gv.x * (ov.x - bv.x) + gv.y * (ov.y - bv.y) + gv.z * (ov.z - bv.z) + gv.w * (ov.w - bv.w);
    for (int o2 = Q >> 1; o2 > 0; o2 >>= 1) d += __shfl_xor_sync(FULL, d, o2);
    if (valid && (gl & (Q - 1)) == 0) {
      const int64_t item = i * p.H + gl / Q;
      p.rowrec[item] = make_float4(__ldg(p.s_dst + item), __ldg(p.rowmax + item), 1.f / (__ldg(p.rowsum + item) + 1e-16f), d);
    }
  }
  // column sums: the RPW groups of a warp, then the 8 warps of the CTA, then one atomic per column per CTA
#pragma unroll
  for (int o2 = G; o2 < 32; o2 <<= 1) {
    cs.x += __shfl_xor_sync(FULL, cs.x, o2); cs.y += __shfl_xor_sync(FULL, cs.y, o2);
    cs.z += __shfl_xor_sync(FULL, cs.z, o2); cs.w += __shfl_xor_sync(FULL, cs.w, o2);
  }
  __shared__ float4 red[8][G];
  if (lane < G) red[threadIdx.x >> 5][lane] = cs;
  __syncthreads();
  if (threadIdx.x < G && threadIdx.x < DQ) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += red[w][threadIdx.x].x; t.y += red[w][threadIdx.x].y; t.z += red[w][threadIdx.x].z; t.w += red[w][threadIdx.x].w; }
    atomicAdd(p.g_bias + 4 * threadIdx.x + 0, t.x); atomicAdd(p.g_bias + 4 * threadIdx.x + 1, t.y);
    atomicAdd(p.g_bias + 4 * threadIdx.x + 2, t.z); atomicAdd(p.g_bias + 4 * threadIdx.x + 3, t.w);
  }
}

template <int G>
static void launch_prep_narrow(const PrepRowsParams& pr, bool act, int blocks, cudaStream_t stream) {
  if (act) bwd_prep_rows_narrow_kernel<true, G><<<blocks, 256, 0, stream>>>(pr);
  else bwd_prep_rows_narrow_kernel<false, G><<<blocks, 256, 0, stream>>>(pr);
}

// ---- fast prep for mean-over-heads layers (!concat, H > 1): one warp per destination row.  The row of the upstream
// gradient (C values, shared by all heads, scaled 1/H) is loaded once, written to the gatherable copy gp [N, Cp] and
// dotted with each head's aggregate O[i,h,:]; the g_bias column sums ride along. --------------------------------------
struct PrepMeanParams {
  int64_t N;
  int H, C, Cp;
  const float* gout; int64_t ldgo;
  const float* o_heads;               // [N, H, Cp], 16-byte aligned
  const float* s_dst; const float* rowmax; const float* rowsum;
  float* gp;                          // [N, Cp]
  int gp16;
  float4* rowrec;
  float* g_bias;                      // [C], zero-initialised, accumulated atomically
};

// ---- the same for head widths up to 128 channels (Cp / 4 <= 32 float4 slots: every mean layer of the BASELINE models — 6 x 121,
// 4 x 47, 8 x 3 ...): a lane GROUP of G = pow2ceil(Cp / 4) lanes per destination row and ONE slot per lane, 32 / G rows per warp,
// and the H per-head dot products reduced together.  (ncu on the 2.4 M-node graph's 4 x 47 layer, warp-per-row form: 12 of 32
// lanes busy, 1 TB/s = 13 % of DRAM peak, 2.9 ms.)
template <int G>
__global__ void __launch_bounds__(256) bwd_prep_mean_narrow_kernel(const PrepMeanParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int Q = p.Cp >> 2, c = 4 * gl;
  const bool live = gl < Q;
  const bool vec_in = (p.C % 4 == 0) && (p.ldgo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.gout) & 15u) == 0);
  const float gscale = 1.f / static_cast<float>(p.H);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = warp * RPW; base < p.N; base += nwarps * RPW) {
    const int64_t i = base + gi;
    const bool valid = live && i < p.N;
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const float* g = p.gout + i * p.ldgo;
      if (vec_in) gv = ldg4(g + c);
      else {
        if (c + 0 < p.C) gv.x = __ldg(g + c + 0);
        if (c + 1 < p.C) gv.y = __ldg(g + c + 1);
        if (c + 2 < p.C) gv.z = __ldg(g + c + 2);
        if (c + 3 < p.C) gv.w = __ldg(g + c + 3);
      }
      cs.x += gv.x; cs.y += gv.y; cs.z += gv.z; cs.w += gv.w;
      gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale;
      if (p.gp16) *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.gp) + i * int64_t(p.Cp) + c) = pack_bf16x4(gv.x, gv.y, gv.z, gv.w);
      else *reinterpret_cast<float4*>(p.gp + i * int64_t(p.Cp) + c) = gv;
    }
    const float* o = p.o_heads + (i < p.N ? i : 0) * int64_t(p.H) * p.Cp + c;
    constexpr int HB = G >= 8 ? 8 : 4;                     // heads whose loads are in flight together
    for (int h0 = 0; h0 < p.H; h0 += HB) {
      float d[HB];
#pragma unroll
      for (int u = 0; u < HB; ++u) {
        d[u] = 0.f;
        if (valid && h0 + u < p.H) {
          const float4 ov = ldg4(o + (h0 + u) * p.Cp);
          d[u] = gv.x * ov.x + gv.y * ov.y + gv.z * ov.z + gv.w * ov.w;
        }
      }
#pragma unroll
      for (int o2 = G >> 1; o2 > 0; o2 >>= 1) {
#pragma unroll
        for (int u = 0; u < HB; ++u) d[u] += __shfl_xor_sync(FULL, d[u], o2, G);
      }
      if (gl < HB && h0 + gl < p.H && i < p.N) {           // lane u of the group writes head h0 + u's record
        float dv = d[0];
#pragma unroll
        for (int u = 1; u < HB; ++u)
          if (gl == u) dv = d[u];
        const int64_t item = i * p.H + h0 + gl;
        p.rowrec[item] = make_float4(__ldg(p.s_dst + item), __ldg(p.rowmax + item), 1.f / (__ldg(p.rowsum + item) + 1e-16f), dv);
      }
    }
  }
  // g_bias column sums: the RPW groups of a warp, the 8 warps of the CTA, one atomic per column per CTA
#pragma unroll
  for (int o2 = G; o2 < 32; o2 <<= 1) {
    cs.x += __shfl_xor_sync(FULL, cs.x, o2); cs.y += __shfl_xor_sync(FULL, cs.y, o2);
    cs.z += __shfl_xor_sync(FULL, cs.z, o2); cs.w += __shfl_xor_sync(FULL, cs.w, o2);
  }
  __shared__ float4 red[8][G];
  if (lane < G) red[threadIdx.x >> 5][lane] = cs;
  __syncthreads();
  if (threadIdx.x < G && threadIdx.x < Q) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += red[w][threadIdx.x].x; t.y += red[w][threadIdx.x].y; t.z += red[w][threadIdx.x].z; t.w += red[w][threadIdx.x].w; }
    const int cc = 4 * threadIdx.x;
    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (cc + u < p.C) atomicAdd(p.g_bias + cc + u, tv[u]);
  }
}

__global__ void __launch_bounds__(256) bwd_prep_mean_rows_kernel(const PrepMeanParams p) {
  constexpr int T = 4;                                    // float4 slots per lane: Cp <= 512
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int Q = p.Cp >> 2;
  const float gscale = 1.f / static_cast<float>(p.H);
  float4 cs[T];
#pragma unroll
  for (int t = 0; t < T; ++t) cs[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = warp; i < p.N; i += nwarps) {
    const float* g = p.gout + i * p.ldgo;
    float4 gv[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int q = lane + 32 * t, c = 4 * q;
      gv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < Q) {
        if (c + 0 < p.C) gv[t].x = __ldg(g + c + 0);
        if (c + 1 < p.C) gv[t].y = __ldg(g + c + 1);
        if (c + 2 < p.C) gv[t].z = __ldg(g + c + 2);
        if (c + 3 < p.C) gv[t].w = __ldg(g + c + 3);
        cs[t].x += gv[t].x; cs[t].y += gv[t].y; cs[t].z += gv[t].z; cs[t].w += gv[t].w;
        gv[t].x *= gscale; gv[t].y *= gscale; gv[t].z *= gscale; gv[t].w *= gscale;
        if (p.gp16) *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.gp) + i * int64_t(p.Cp) + c) = pack_bf16x4(gv[t].x, gv[t].y, gv[t].z, gv[t].w);
        else *reinterpret_cast<float4*>(p.gp + i * int64_t(p.Cp) + c) = gv[t];
      }
    }
    const float* o = p.o_heads + i * int64_t(p.H) * p.Cp;
    for (int h = 0; h < p.H; ++h) {
      float d = 0.f;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int q = lane + 32 * t;
        if (q < Q) {
          const float4 ov = ldg4(o + h * p.Cp + 4 * q);
          d += gv[t].x * ov.x + gv[t].y * ov.y + gv[t].z * ov.z + gv[t].w * ov.w;
        }
      }
      d = group_sum<32>(d);
      if (lane == 0) {
        const int64_t item = i * p.H + h;
        p.rowrec[item] = make_float4(__ldg(p.s_dst + item), __ldg(p.rowmax + item), 1.f / (__ldg(p.rowsum + item) + 1e-16f), d);
      }
    }
  }
  __shared__ float4 red[8][32];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    if (32 * t >= Q) break;
    __syncthreads();
    red[threadIdx.x >> 5][lane] = cs[t];
    __syncthreads();
    if (threadIdx.x < 32) {
      float4 s4 = red[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) { s4.x += red[w][lane].x; s4.y += red[w][lane].y; s4.z += red[w][lane].z; s4.w += red[w][lane].w; }
      const int c = 4 * (lane + 32 * t);
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u < p.C) atomicAdd(p.g_bias + c + u, sv[u]);
    }
  }
}

// ---- column sums of a [N, ncols] matrix (g_bias = sum_n gout[n,:]) ------------------------------------------------
// (C, Cp): output column c reads input column (c / C) * Cp + c % C — the head-padded copy gp; C == Cp: identity
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ in, int64_t ld, int64_t N, int ncols, int C, int Cp, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int cin = (C == Cp) ? c : (c / C) * Cp + c % C;
  float s = 0.f;
  if (c < ncols)
    for (int64_t r = int64_t(blockIdx.y) * 8 + ty; r < N; r += int64_t(gridDim.y) * 8) s += __ldg(in + r * ld + cin);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < ncols) {
#pragma unroll
    for (int u = 1; u < 8; ++u) s += red[u][tx];
    atomicAdd(out + c, s);
  }
}

// ---- the CSC pass ---------------------------------------------------------------------------------------------------
struct EdgeBwdParams {
  int64_t N, items;
  int H, Cp, Dp;
  float slope; int act;
  const int32_t* colptr; const int32_t* crow; const int32_t* ceid;
  const float* wh; const float* s_src; const float4* rowrec;
  DropoutSpec drop;                        // attention dropout: mask tensor or in-kernel Philox (common.cuh)
  const float* g; int64_t ldg; int hs;     // G[i,h,c] = g[i*ldg + h*hs + c]   (hs = 0: shared by all heads)
  const void* g16;                         // the same rows stored as bf16 (ROW16 instantiations gather these instead)
  float* gwh;                              // [N, Dp]
  float* g_s_src; float* g_s_dst;          // [N, H]; g_s_dst is zero-initialised and accumulated atomically
  // scheduling by degree (b200gat_graph.hub_cols): colend[j] = colptr[j + 1] except for hub source rows, which look
  // EMPTY to the row-per-group kernels (they write zeros) and are then walked — and their outputs overwritten — by
  // edge_bwd_hub_kernel, one CTA per (source row, head).  No compare, no extra register in the hot kernels.
  const int32_t* colend; const int32_t* hub; int64_t nhub;
  int max_deg;          // largest out-degree: source rows above B200GAT_GIANT_DEGREE are split into segments (grid.y)
  float* de;            // B200GAT_LOGIT_HEAD_SOFTMAX only: [E', H] d loss / d e per CSC entry (the heads of an edge are coupled:
                        // dz is formed by head_softmax_bwd_kernel); the CSC pass then leaves g_s_src = 0 and g_s_dst untouched
  uint32_t* amax;       // optional [2]: bit patterns of max|gWh| and max|g_s_src| (atomicMax; zeroed by the host) — saves
                        // gt_amax_kernel's pass over gWh.  Not exact for giant rows (partial sums): the host ignores it then
};

// Sum U per-lane partials over the G lanes of a group and hand the total of edge u to the lane with rel == u.
// Full-warp groups use a transposing butterfly (7 shuffles for 4 edges instead of 20).
template <int G, int U>
__device__ __forceinline__ float reduce_deliver(float (&d)[U], int lane, int rel) {
  constexpr unsigned FULL = 0xffffffffu;
  float got = 0.f;
  if (G == 32 && U == 8) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float k4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                  // bit4 = 0 keeps edges 0..3, bit4 = 1 keeps edges 4..7
      const float r = __shfl_xor_sync(FULL, b4 ? d[u] : d[u + 4], 16);
      k4[u] = (b4 ? d[u + 4] : d[u]) + r;
    }
    float k2[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {                  // bit3 picks the pair
      const float r = __shfl_xor_sync(FULL, b3 ? k4[u] : k4[u + 2], 8);
      k2[u] = (b3 ? k4[u + 2] : k4[u]) + r;
    }
    const float r = __shfl_xor_sync(FULL, b2 ? k2[0] : k2[1], 4);
    float k = (b2 ? k2[1] : k2[0]) + r;            // this lane now owns edge 4*bit4 + 2*bit3 + bit2
    k += __shfl_xor_sync(FULL, k, 2);
    k += __shfl_xor_sync(FULL, k, 1);
    got = __shfl_sync(FULL, k, ((rel >> 2) & 1) * 16 + ((rel >> 1) & 1) * 8 + (rel & 1) * 4);
  } else if (G == 32 && U == 4) {
    const bool b4 = lane & 16, b3 = lane & 8;
    const float r0 = __shfl_xor_sync(FULL, b4 ? d[0] : d[2], 16);
    const float r1 = __shfl_xor_sync(FULL, b4 ? d[1] : d[3], 16);
    const float k0 = (b4 ? d[2] : d[0]) + r0;      // lanes with bit4 = 0 keep edges 0,1; bit4 = 1 keep edges 2,3
    const float k1 = (b4 ? d[3] : d[1]) + r1;
    const float r = __shfl_xor_sync(FULL, b3 ? k0 : k1, 8);
    float k = (b3 ? k1 : k0) + r;                  // this lane now owns edge 2*bit4 + bit3
    k += __shfl_xor_sync(FULL, k, 4);
    k += __shfl_xor_sync(FULL, k, 2);
    k += __shfl_xor_sync(FULL, k, 1);
    got = __shfl_sync(FULL, k, ((rel >> 1) & 1) * 16 + (rel & 1) * 8);
  } else if (G == 32 && U == 2) {
    const bool b4 = lane & 16;
    const float r = __shfl_xor_sync(FULL, b4 ? d[0] : d[1], 16);
    float k = (b4 ? d[1] : d[0]) + r;              // bit4 selects the edge
    k += __shfl_xor_sync(FULL, k, 8);
    k += __shfl_xor_sync(FULL, k, 4);
    k += __shfl_xor_sync(FULL, k, 2);
    k += __shfl_xor_sync(FULL, k, 1);
    got = __shfl_sync(FULL, k, (rel & 1) * 16);
  } else {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) d[u] += __shfl_xor_sync(FULL, d[u], o, G);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (rel == u) got = d[u];
  }
  return got;
}

template <int G, int NV, bool HAS_MASK, bool GENERIC, bool HUB, bool ROW16 = false>
__device__ __forceinline__ void edge_bwd_body(const EdgeBwdParams& p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int GPW = 32 / G;
  // edges gathered per step: bytes in flight decide the streaming-regime throughput (measured: 4 -> 8 gathers in
  // flight per lane took the large-graph forward from 2.9 to 3.7 TB/s); registers bound it for wide heads
  // (backward: 8 per step measured SLOWER than 4 on the large graph, 44.7 vs 41.0 ms — the dependent rowrec gather
  //  and the longer transposing reduce outweigh the extra loads in flight.
  //  Also measured: two gather batches sharing ONE 8-value transposing reduce (19 instead of 2 x 14 shuffle / select / add
  //  instructions per 8 edges) — the four extra live partials spill at 64 registers: PPI edge_bwd 734 -> 756 us.)
  constexpr int U = NV <= 2 ? 4 : 2;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int H = p.H, Cp = p.Cp, Q = p.Cp >> 2;
  const int64_t Dp = p.Dp;
  const uint32_t g_row_bytes = static_cast<uint32_t>(p.ldg) * (ROW16 ? 2u : 4u);
  const float slope = p.slope;
  const int act = GENERIC ? p.act : 0;
  // lanes beyond the head width gather a clamped (valid) column and are never stored; their Wh slice is zero
  int off[NV];
  bool live[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    live[v] = gl + v * G < Q;
    off[v] = 4 * (live[v] ? gl + v * G : Q - 1);
  }

  const DropoutKey dkey = HAS_MASK ? dropout_key(p.drop) : DropoutKey{0u, 0u, 0u, 0u};
  float amax_w = 0.f, amax_s = 0.f;
  for (int64_t base = warp * GPW; base < p.items; base += nwarps * GPW) {
    const int64_t item = base + gi;
    const bool valid = item < p.items;
    const int64_t j = valid ? item / H : 0;
    const int h = valid ? static_cast<int>(item - j * H) : 0;
    const int beg = valid ? __ldg(p.colptr + j) : 0;
    // HUB: hub rows are empty here (edge_bwd_hub_kernel).  A separate instantiation because even the extra pointer of
    // the colend read perturbs this kernel's code generation (measured +4 % on the PPI-shaped batch, which has no hubs)
    const int end = valid ? (HUB ? __ldg(p.colend + j) : __ldg(p.colptr + j + 1)) : 0;
    const int deg = end - beg;
    const int maxdeg = GPW == 1 ? deg : __reduce_max_sync(FULL, deg);
    const float ss = valid ? __ldg(p.s_src + j * H + h) : 0.f;

    float4 whv[NV], acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      whv[v] = (valid && live[v]) ? ldg4(p.wh + j * Dp + h * Cp + off[v]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float gsrc = 0.f;
    const char* gb[NV];                                   // this lane's column slices of row 0 of G[:, h, :]
#pragma unroll
    for (int v = 0; v < NV; ++v) gb[v] = row_base<ROW16>(p.g, p.g16, h * p.hs + off[v]);

    for (int k0 = 0; k0 < maxdeg; k0 += G) {
      const int k = beg + k0 + gl;
      const bool ok = k < end;
      int i = static_cast<int>(j);
      float alpha = 0.f, at = 0.f, mk = 1.f, dr = 0.f, dslope = 0.f;
      if (ok) {
        i = __ldg(p.crow + k);
        const float4 rr = __ldg(p.rowrec + int64_t(i) * H + h);   // {s_dst, rowmax, 1/(rowsum+eps), Drow}
        const float z = rr.x + ss;
        if (GENERIC && act == B200GAT_LOGIT_HEAD_SOFTMAX) {
          const float e = head_softmax(reinterpret_cast<const float*>(p.rowrec + int64_t(i) * H), 4, p.s_src + j * H, h, H);
          alpha = expf(e - rr.y) * rr.z;
        } else {
          dslope = logit_act_grad<GENERIC>(z, slope, act);
          alpha = expf(logit_act<GENERIC>(z, slope, act) - rr.y) * rr.z;
        }
        if (HAS_MASK && (!GENERIC || p.drop.active())) mk = dropout_mult(p.drop, dkey, __ldg(p.ceid + k), h, H);
        at = alpha * mk;
        dr = rr.w;
      }
      float dot_mine = 0.f;
      const int cnt = (maxdeg - k0) < G ? (maxdeg - k0) : G;
      for (int t = 0; t < cnt; t += U) {
        int it[U];
        float a_t[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          it[u] = __shfl_sync(FULL, i, t + u, G);
          a_t[u] = __shfl_sync(FULL, at, t + u, G);       // slots past the chunk's edges carry at = 0 ...
          if (U > G && t + u >= cnt) a_t[u] = 0.f;        // ... unless the batch is wider than the group (shuffle wraps)
        }
        float4 g4[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (HAS_MASK) {   // dropped edges (60 % under the reference's p = 0.6): dz needs no dot product, skip the gather
              g4[u][v] = a_t[u] != 0.f ? gather4<ROW16>(gb[v], it[u], g_row_bytes) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              g4[u][v] = gather4<ROW16>(gb[v], it[u], g_row_bytes);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          d[u] = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(a_t[u], g4[u][v].x, acc[v].x);
            acc[v].y = fmaf(a_t[u], g4[u][v].y, acc[v].y);
            acc[v].z = fmaf(a_t[u], g4[u][v].z, acc[v].z);
            acc[v].w = fmaf(a_t[u], g4[u][v].w, acc[v].w);
            d[u] = fmaf(g4[u][v].x, whv[v].x, d[u]);
            d[u] = fmaf(g4[u][v].y, whv[v].y, d[u]);
            d[u] = fmaf(g4[u][v].z, whv[v].z, d[u]);
            d[u] = fmaf(g4[u][v].w, whv[v].w, d[u]);
          }
        }
        const int rel = gl - t;
        const float got = reduce_deliver<G, U>(d, lane, rel);
        if (rel >= 0 && rel < U) dot_mine = got;
      }
      if (ok) {
        if (GENERIC && act == B200GAT_LOGIT_HEAD_SOFTMAX) {
          p.de[int64_t(k) * H + h] = alpha * (mk * dot_mine - dr);        // d loss / d e; dz needs the edge's other heads
        } else {
          const float dz = alpha * (mk * dot_mine - dr) * dslope;
          gsrc += dz;
          atomicAdd(p.g_s_dst + int64_t(i) * H + h, dz);
        }
      }
    }
    gsrc = group_sum<G>(gsrc);
    if (!valid) continue;
    if (gl == 0) p.g_s_src[j * H + h] = gsrc;
    amax_s = fmaxf(amax_s, fabsf(gsrc));
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (live[v]) {
        *reinterpret_cast<float4*>(p.gwh + j * Dp + h * Cp + off[v]) = acc[v];
        amax_w = fmaxf(amax_w, fmaxf(fmaxf(fabsf(acc[v].x), fabsf(acc[v].y)), fmaxf(fabsf(acc[v].z), fabsf(acc[v].w))));
      }
  }
  if (p.amax) {
    warp_atomic_amax(p.amax, amax_w);
    warp_atomic_amax(p.amax + 1, amax_s);
  }
}

// 4 CTAs per SM (64 registers) for heads up to 256 channels: without the bound ptxas picks 64 or 80 registers depending
// on details as small as the amax bookkeeping at the end of an item
template <int G, int NV, bool HAS_MASK, bool HUB>
__global__ void __launch_bounds__(256, (NV <= 2 && !HAS_MASK) ? 4 : 2) edge_bwd_kernel(const EdgeBwdParams p) {
  edge_bwd_body<G, NV, HAS_MASK, false, HUB>(p);
}
// bf16-stored gradient rows (no mask, LeakyReLU)
template <int G, int NV, bool HUB>
__global__ void __launch_bounds__(256, NV <= 2 ? 4 : 2) edge_bwd16_kernel(const EdgeBwdParams p) {
  edge_bwd_body<G, NV, false, false, HUB, true>(p);
}
// other logit activations (run_act_func_experiment.py): one mask-capable instantiation per geometry (mask may be NULL)
template <int G, int NV>
__global__ void __launch_bounds__(256) edge_bwd_act_kernel(const EdgeBwdParams p) { edge_bwd_body<G, NV, true, true, true>(p); }

// ---- the CSC pass of mean-over-heads layers (GAT.py:65-66; hs == 0): G[i,:] = gout[i,:] / H is the SAME row for every
// head, so one lane group per SOURCE ROW j walks the column once for all HH heads — G[i] is gathered once per edge
// instead of once per (edge, head), the col[] -> rowrec -> G dependent chain is paid once per chunk instead of HH
// times, and the HH row records of a destination are one contiguous 16*HH-byte read.  Each lane keeps its 4-channel
// slice of Wh[j,h,:] and of the gWh accumulator for every head in registers (Cp <= 128).
// Transposing reduce: V values per lane summed over the G lanes of the group in V - 1 + log2(G / V) shuffles; afterwards
// lane gl holds the total of value index gl / (G / V).
template <int G, int V>
__device__ __forceinline__ float transpose_reduce(float (&d)[V], int gl) {
  constexpr unsigned FULL = 0xffffffffu;
  static_assert(V <= G && (V & (V - 1)) == 0, "V must be a power of two <= G");
#pragma unroll
  for (int s = 1; (V >> s) >= 1; ++s) {
    const int half = V >> s, o = G >> s;
    const bool up = gl & o;
#pragma unroll
    for (int v = 0; v < half; ++v) {
      const float r = __shfl_xor_sync(FULL, up ? d[v] : d[v + half], o, G);
      d[v] = (up ? d[v + half] : d[v]) + r;
    }
  }
  float k = d[0];
#pragma unroll
  for (int o = G / (2 * V); o > 0; o >>= 1) k += __shfl_xor_sync(FULL, k, o, G);
  return k;
}

__host__ __device__ constexpr int pow2ceil_c(int v) { int r = 1; while (r < v) r <<= 1; return r; }

template <int G, int HH, bool HAS_MASK, bool ROW16 = false>
__global__ void __launch_bounds__(256, 2) edge_bwd_mean_kernel(const EdgeBwdParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int GPW = 32 / G;
  constexpr int HP = pow2ceil_c(HH);
  static_assert(HP <= G, "heads (padded to a power of two) must fit the lane group");
  constexpr int UMAX = HH > 6 ? 1 : (HH > 4 ? 2 : 4);
  constexpr int U = (G / HP) < UMAX ? (G / HP) : UMAX;    // edges per gather batch; V = U * HP values per reduce
  constexpr int V = U * HP;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int Cp = p.Cp, Q = p.Cp >> 2;
  const int64_t Dp = p.Dp;
  const uint32_t g_row_bytes = static_cast<uint32_t>(p.ldg) * (ROW16 ? 2u : 4u);
  const float slope = p.slope;
  const bool live = gl < Q;                               // dead lanes gather a clamped (valid) column; their Wh slice is zero
  const int off = 4 * (live ? gl : Q - 1);
  const char* gb = row_base<ROW16>(p.g, p.g16, off);

  const DropoutKey dkey = HAS_MASK ? dropout_key(p.drop) : DropoutKey{0u, 0u, 0u, 0u};
  float amax_w = 0.f, amax_s = 0.f;
  for (int64_t base = warp * GPW; base < p.N; base += nwarps * GPW) {
    const int64_t j = base + gi < p.N ? base + gi : p.N - 1;
    const bool valid = base + gi < p.N;
    const int beg = valid ? __ldg(p.colptr + j) : 0;
    const int end = valid ? __ldg(p.colend + j) : 0;            // hub rows are empty here (edge_bwd_hub_kernel)
    const int deg = end - beg;
    const int maxdeg = GPW == 1 ? deg : __reduce_max_sync(FULL, deg);
    float ss[HH], gsrc[HH];
    float4 whv[HH], acc[HH];
#pragma unroll
    for (int h = 0; h < HH; ++h) {
      ss[h] = __ldg(p.s_src + j * HH + h);
      gsrc[h] = 0.f;
      acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      whv[h] = live ? ldg4(p.wh + j * Dp + h * Cp + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (int k0 = 0; k0 < maxdeg; k0 += G) {
      const int k = beg + k0 + gl;
      const bool ok = k < end;
      int i = static_cast<int>(j);
      float at[HH], ca[HH], cb[HH];                       // alpha * mask;  dz = ca * <G[i], Wh[j,h]> - cb
#pragma unroll
      for (int h = 0; h < HH; ++h) at[h] = ca[h] = cb[h] = 0.f;
      if (ok) {
        i = __ldg(p.crow + k);
        const float4* rec = p.rowrec + int64_t(i) * HH;
        const int eid_k = HAS_MASK ? __ldg(p.ceid + k) : 0;
#pragma unroll
        for (int h = 0; h < HH; ++h) {
          const float4 rr = __ldg(rec + h);               // {s_dst, rowmax, 1/(rowsum+eps), Drow}
          const float z = rr.x + ss[h];
          const float alpha = expf(leaky(z, slope) - rr.y) * rr.z;
          const float ad = alpha * (z > 0.f ? 1.f : slope);
          const float mk = HAS_MASK ? dropout_mult(p.drop, dkey, eid_k, h, HH) : 1.f;
          at[h] = alpha * mk;
          ca[h] = ad * mk;
          cb[h] = ad * rr.w;
        }
      }
      float dot[HH];
#pragma unroll
      for (int h = 0; h < HH; ++h) dot[h] = 0.f;
      const int cnt = (maxdeg - k0) < G ? (maxdeg - k0) : G;
      for (int t = 0; t < cnt; t += U) {                  // t + u < G always; lanes past the chunk's edges carry at = 0
        float4 g4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int it = __shfl_sync(FULL, i, t + u, G);
          g4[u] = gather4<ROW16>(gb, it, g_row_bytes);
        }
        float d[V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int h = 0; h < HP; ++h) {
            if (h < HH) {
              const float a = __shfl_sync(FULL, at[h < HH ? h : 0], t + u, G);
              acc[h < HH ? h : 0].x = fmaf(a, g4[u].x, acc[h < HH ? h : 0].x);
              acc[h < HH ? h : 0].y = fmaf(a, g4[u].y, acc[h < HH ? h : 0].y);
              acc[h < HH ? h : 0].z = fmaf(a, g4[u].z, acc[h < HH ? h : 0].z);
              acc[h < HH ? h : 0].w = fmaf(a, g4[u].w, acc[h < HH ? h : 0].w);
              const float4 w = whv[h < HH ? h : 0];
              d[u * HP + h] = fmaf(g4[u].x, w.x, fmaf(g4[u].y, w.y, fmaf(g4[u].z, w.z, g4[u].w * w.w)));
            } else {
              d[u * HP + h] = 0.f;
            }
          }
        }
        const float tot = transpose_reduce<G, V>(d, gl);  // lane gl: total of value (gl / (G / V)) = edge u * HP + head h
        const int rel = gl - t;
        const bool mine = rel >= 0 && rel < U;
        const int src0 = (mine ? rel : 0) * HP * (G / V);
#pragma unroll
        for (int h = 0; h < HH; ++h) {
          const float got = __shfl_sync(FULL, tot, src0 + h * (G / V), G);
          if (mine) dot[h] = got;
        }
      }
      if (ok) {
#pragma unroll
        for (int h = 0; h < HH; ++h) {
          const float dz = ca[h] * dot[h] - cb[h];
          gsrc[h] += dz;
          atomicAdd(p.g_s_dst + int64_t(i) * HH + h, dz);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < HH; ++h) gsrc[h] = group_sum<G>(gsrc[h]);
    if (!valid) continue;
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < HH; ++h) p.g_s_src[j * HH + h] = gsrc[h];
    }
#pragma unroll
    for (int h = 0; h < HH; ++h) amax_s = fmaxf(amax_s, fabsf(gsrc[h]));
    if (live) {
#pragma unroll
      for (int h = 0; h < HH; ++h) {
        *reinterpret_cast<float4*>(p.gwh + j * Dp + h * Cp + off) = acc[h];
        amax_w = fmaxf(amax_w, fmaxf(fmaxf(fabsf(acc[h].x), fabsf(acc[h].y)), fmaxf(fabsf(acc[h].z), fabsf(acc[h].w))));
      }
    }
  }
  if (p.amax) {
    warp_atomic_amax(p.amax, amax_w);
    warp_atomic_amax(p.amax + 1, amax_s);
  }
}

template <int G, int HH>
static int launch_edge_bwd_mean(const EdgeBwdParams& p, cudaStream_t stream) {
  constexpr int GPW = 32 / G;
  const int threads = 256;
  const int64_t want = ceil_div(ceil_div(p.N, GPW), threads / 32);
  const int64_t cap = int64_t(sm_count()) * 8;
  const int blocks = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  if (p.g16) edge_bwd_mean_kernel<G, HH, false, true><<<blocks, threads, 0, stream>>>(p);
  else if (p.drop.active()) edge_bwd_mean_kernel<G, HH, true><<<blocks, threads, 0, stream>>>(p);
  else edge_bwd_mean_kernel<G, HH, false><<<blocks, threads, 0, stream>>>(p);
  return check_launch("edge_bwd_mean_kernel");
}

template <int HH>
static int dispatch_edge_bwd_mean(const EdgeBwdParams& p, int Q, cudaStream_t stream) {
  constexpr int HP = pow2ceil_c(HH);
  if constexpr (HP <= 4) {
    if (Q <= 4) return launch_edge_bwd_mean<4, HH>(p, stream);
  }
  if (Q <= 8) return launch_edge_bwd_mean<8, HH>(p, stream);
  if (Q <= 16) return launch_edge_bwd_mean<16, HH>(p, stream);
  return launch_edge_bwd_mean<32, HH>(p, stream);
}

// the heads-shared CSC pass covers H in {2, 3, 4, 6, 8}, head width <= 128 channels, LeakyReLU logits
static bool edge_bwd_mean_supported(int H, int Q, int hs, int act) {
  return hs == 0 && act == B200GAT_LOGIT_LEAKY_RELU && Q <= 32 && (H == 2 || H == 3 || H == 4 || H == 6 || H == 8);
}

// ---- hub source rows (out-degree > B200GAT_HUB_DEGREE): one CTA per (source row, head); the 8 warps walk interleaved
// 32-edge chunks of the column (same arithmetic as edge_bwd_body with a full-warp group) and merge their gWh / g_s_src
// partial sums through shared memory.  Handles every logit activation and the optional mask at run time.
template <int NV, bool ROW16 = false>
__global__ void __launch_bounds__(256) edge_bwd_hub_kernel(const EdgeBwdParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int U = NV == 1 ? 8 : (NV == 2 ? 4 : 2);
  __shared__ float sm_g[8];
  __shared__ float4 sm_acc[8][32 * NV];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int H = p.H, Cp = p.Cp, Q = p.Cp >> 2;
  const int64_t Dp = p.Dp;
  const uint32_t g_row_bytes = static_cast<uint32_t>(p.ldg) * (ROW16 ? 2u : 4u);
  const float slope = p.slope;
  const int act = p.act;
  int off[NV];
  bool live[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    live[v] = lane + v * 32 < Q;
    off[v] = 4 * (live[v] ? lane + v * 32 : Q - 1);
  }
  const DropoutKey dkey = dropout_key(p.drop);
  float amax_w = 0.f, amax_s = 0.f;
  for (int64_t item = blockIdx.x; item < p.nhub * H; item += gridDim.x) {
    const int64_t j = __ldg(p.hub + item / H);
    const int h = static_cast<int>(item % H);
    // giant source rows (out-degree > B200GAT_GIANT_DEGREE) are cut into segments, one CTA each (grid.y = segment); their
    // partial sums are ADDED to the zeros the row-per-group kernel left in gwh / g_s_src.  CTA-uniform control flow.
    const int beg0 = __ldg(p.colptr + j), end0 = __ldg(p.colptr + j + 1);
    const bool giant = end0 - beg0 > B200GAT_GIANT_DEGREE;
    if (!giant && blockIdx.y > 0) continue;
    const int beg = giant ? beg0 + static_cast<int>(blockIdx.y) * B200GAT_GIANT_DEGREE : beg0;
    const int end = giant && beg + B200GAT_GIANT_DEGREE < end0 ? beg + B200GAT_GIANT_DEGREE : end0;
    if (beg >= end) continue;
    const float ss = __ldg(p.s_src + j * H + h);
    float4 whv[NV], acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      whv[v] = live[v] ? ldg4(p.wh + j * Dp + h * Cp + off[v]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float gsrc = 0.f;
    const char* gb[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) gb[v] = row_base<ROW16>(p.g, p.g16, h * p.hs + off[v]);
    for (int k0 = beg + w * 32; k0 < end; k0 += 256) {
      const int k = k0 + lane;
      const bool ok = k < end;
      int i = static_cast<int>(j);
      float alpha = 0.f, at = 0.f, mk = 1.f, dr = 0.f, dslope = 0.f;
      if (ok) {
        i = __ldg(p.crow + k);
        const float4 rr = __ldg(p.rowrec + int64_t(i) * H + h);   // {s_dst, rowmax, 1/(rowsum+eps), Drow}
        const float z = rr.x + ss;
        if (act == B200GAT_LOGIT_HEAD_SOFTMAX) {
          const float e = head_softmax(reinterpret_cast<const float*>(p.rowrec + int64_t(i) * H), 4, p.s_src + j * H, h, H);
          alpha = expf(e - rr.y) * rr.z;
        } else {
          dslope = logit_act_grad<true>(z, slope, act);
          alpha = expf(logit_act<true>(z, slope, act) - rr.y) * rr.z;
        }
        if (p.drop.active()) mk = dropout_mult(p.drop, dkey, __ldg(p.ceid + k), h, H);
        at = alpha * mk;
        dr = rr.w;
      }
      float dot_mine = 0.f;
      const int cnt = (end - k0) < 32 ? (end - k0) : 32;
      for (int t = 0; t < cnt; t += U) {
        int it[U];
        float a_t[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          it[u] = __shfl_sync(FULL, i, t + u);
          a_t[u] = __shfl_sync(FULL, at, t + u);
        }
        float4 g4[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int v = 0; v < NV; ++v) g4[u][v] = gather4<ROW16>(gb[v], it[u], g_row_bytes);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          d[u] = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(a_t[u], g4[u][v].x, acc[v].x);
            acc[v].y = fmaf(a_t[u], g4[u][v].y, acc[v].y);
            acc[v].z = fmaf(a_t[u], g4[u][v].z, acc[v].z);
            acc[v].w = fmaf(a_t[u], g4[u][v].w, acc[v].w);
            d[u] = fmaf(g4[u][v].x, whv[v].x, d[u]);
            d[u] = fmaf(g4[u][v].y, whv[v].y, d[u]);
            d[u] = fmaf(g4[u][v].z, whv[v].z, d[u]);
            d[u] = fmaf(g4[u][v].w, whv[v].w, d[u]);
          }
        }
        const int rel = lane - t;
        const float got = reduce_deliver<32, U>(d, lane, rel);
        if (rel >= 0 && rel < U) dot_mine = got;
      }
      if (ok) {
        if (act == B200GAT_LOGIT_HEAD_SOFTMAX) {
          p.de[int64_t(k) * H + h] = alpha * (mk * dot_mine - dr);
        } else {
          const float dz = alpha * (mk * dot_mine - dr) * dslope;
          gsrc += dz;
          atomicAdd(p.g_s_dst + int64_t(i) * H + h, dz);
        }
      }
    }
    gsrc = group_sum<32>(gsrc);
    if (lane == 0) sm_g[w] = gsrc;
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
      float* dst = p.gwh + j * Dp + h * Cp + 4 * q;
      if (giant) { atomicAdd(dst, o.x); atomicAdd(dst + 1, o.y); atomicAdd(dst + 2, o.z); atomicAdd(dst + 3, o.w); }
      else *reinterpret_cast<float4*>(dst) = o;
      amax_w = fmaxf(amax_w, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
    }
    if (threadIdx.x == 0) {
      float g = sm_g[0];
#pragma unroll
      for (int w2 = 1; w2 < 8; ++w2) g += sm_g[w2];
      if (giant) atomicAdd(p.g_s_src + j * H + h, g);
      else p.g_s_src[j * H + h] = g;
      amax_s = fmaxf(amax_s, fabsf(g));
    }
    __syncthreads();
  }
  if (p.amax) {                                           // (ignored by the host when giant rows exist)
    warp_atomic_amax(p.amax, amax_w);
    warp_atomic_amax(p.amax + 1, amax_s);
  }
}

static int launch_edge_bwd_hub(const EdgeBwdParams& p, cudaStream_t stream) {
  if (!p.hub || p.nhub <= 0) return 0;
  const int64_t want = p.nhub * p.H;
  const int64_t cap = int64_t(sm_count()) * 8;
  const int64_t nseg = p.max_deg > B200GAT_GIANT_DEGREE ? ceil_div(p.max_deg, B200GAT_GIANT_DEGREE) : 1;
  B200GAT_REQUIRE(nseg <= 65535, B200GAT_E_UNSUPPORTED, "edge_bwd: a source row of %d edges is not supported", p.max_deg);
  const dim3 blocks(static_cast<unsigned>(want < cap ? want : cap), static_cast<unsigned>(nseg));
  const int Q = p.Cp / 4;
  if (p.g16) {
    if (Q <= 32) edge_bwd_hub_kernel<1, true><<<blocks, 256, 0, stream>>>(p);
    else if (Q <= 64) edge_bwd_hub_kernel<2, true><<<blocks, 256, 0, stream>>>(p);
    else edge_bwd_hub_kernel<4, true><<<blocks, 256, 0, stream>>>(p);
  }
  else if (Q <= 32) edge_bwd_hub_kernel<1><<<blocks, 256, 0, stream>>>(p);
  else if (Q <= 64) edge_bwd_hub_kernel<2><<<blocks, 256, 0, stream>>>(p);
  else edge_bwd_hub_kernel<4><<<blocks, 256, 0, stream>>>(p);
  return check_launch("edge_bwd_hub_kernel");
}

// ---- gT = gWh + g_s_src (x) a1 + g_s_dst (x) a2, plus every column sum the parameters need.  gT is written in
// place as fp32, or — when the projection backward runs on the tensor cores — directly as that GEMM's operand split
// (two fp16 planes, split_blob.cuh), which saves the fp32 write and the split pass over [N, Dp].
// The split's scale needs max|gT| BEFORE the pass: it uses the bound max|gWh| + max|g_s_src| max|a1| + max|g_s_dst| max|a2|
// (gt_amax_kernel), at most a couple of binades loose. ----
struct FinishParams {
  int64_t N;
  int H, Cp, Dp;
  const float* wh; const float* a1; const float* a2; const float* g_s_src; const float* g_s_dst;
  float* g_t;                              // in: gWh, out: gT (fp32 mode)
  float* g_bw; float* g_a1; float* g_a2;   // [Dp] zero-initialised
  float* g_b1; float* g_b2;                // [H]  zero-initialised
  const uint32_t* amax;                    // planes mode: [5] bit patterns {gWh, g_s_src, g_s_dst, a1, a2}
  __half* hi; __half* lo; int64_t ldp;     // planes mode: the split blob's planes
  float* inv_scale; uint32_t* bound_bits;  // planes mode: blob header
};

// amax[0] = max|gWh| over [n_nd], amax[1] / amax[2] = max|g_s_src| / max|g_s_dst| over [n_nh], amax[3] / amax[4] = max|a1| /
// max|a2| over [Dp] (slots zeroed by the host)
__global__ void __launch_bounds__(256)
gt_amax_kernel(const float* __restrict__ gwh, int64_t n_nd, const float* __restrict__ g_s_src, const float* __restrict__ g_s_dst,
               int64_t n_nh, const float* __restrict__ a1, const float* __restrict__ a2, int Dp, uint32_t* __restrict__ amax) {
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
  float m = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(gwh);       // 16-byte aligned, n_nd % 4 == 0 (Dp = H * c_pad)
  for (int64_t t = tid; t < (n_nd >> 2); t += nth) {
    const float4 v = __ldg(g4 + t);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  warp_atomic_amax(amax, m);
  float ms = 0.f, md = 0.f;
  for (int64_t t = tid; t < n_nh; t += nth) {
    ms = fmaxf(ms, fabsf(__ldg(g_s_src + t)));
    md = fmaxf(md, fabsf(__ldg(g_s_dst + t)));
  }
  warp_atomic_amax(amax + 1, ms);
  warp_atomic_amax(amax + 2, md);
  if (blockIdx.x == 0) {
    float m1 = 0.f, m2 = 0.f;
    for (int c = threadIdx.x; c < Dp; c += blockDim.x) {
      m1 = fmaxf(m1, fabsf(__ldg(a1 + c)));
      m2 = fmaxf(m2, fabsf(__ldg(a2 + c)));
    }
    warp_atomic_amax(amax + 3, m1);
    warp_atomic_amax(amax + 4, m2);
  }
}

// one thread per FOUR adjacent columns (same head: c_pad % 4 == 0); the 256 threads of a CTA form TX column groups x
// RY = 256 / TX row lanes (TX = min(256, pow2ceil(Dp / 4)): narrow layers — the 8 x 8 layers of GATNet have Dp = 64 — would
// otherwise leave 240 of 256 threads idle: 49 us for 15 k rows, measured), a strip of row batches per blockIdx.y
template <bool PLANES, int TXS>
__global__ void __launch_bounds__(256, 4) bwd_finish_kernel(const FinishParams p) {
  constexpr int TX = 1 << TXS, RY = 256 >> TXS;
  const int tx = threadIdx.x & (TX - 1), ry = threadIdx.x >> TXS;
  const int c = 4 * (blockIdx.x * TX + tx);
  const bool active = c < p.Dp;
  if constexpr (RY == 1) {
    if (!active) return;                                  // (no CTA-wide barrier below in this form)
  }
  const int cc = (RY == 1 || active) ? c : 0;
  const int h = cc / p.Cp;
  const bool head_lead = (cc - h * p.Cp) == 0;
  const float4 a1c = ldg4(p.a1 + cc), a2c = ldg4(p.a2 + cc);
  float scale = 1.f;
  if (PLANES) {
    const float bound = __uint_as_float(p.amax[0]) + __uint_as_float(p.amax[1]) * __uint_as_float(p.amax[3]) +
                        __uint_as_float(p.amax[2]) * __uint_as_float(p.amax[4]);
    scale = scale_from_amax(__float_as_uint(bound));
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      *p.inv_scale = 1.f / scale;
      *p.bound_bits = __float_as_uint(bound);
    }
  }
  float4 sbw = make_float4(0.f, 0.f, 0.f, 0.f), sa1 = sbw, sa2 = sbw;
  float sb1 = 0.f, sb2 = 0.f;
  // Row batches are INTERLEAVED over the CTAs (batch b goes to CTA b mod gridDim.y), so at any time the grid streams one
  // window of ~gridDim.y * RY * RB consecutive rows of each array — DRAM pages are walked in address order, as a plain copy
  // does — instead of gridDim.y separate strips (thousands of concurrently open pages).
  // rows in batches of RB: ALL loads of a batch are issued before its first store — the stores (g_t / planes) may alias
  // the loads as far as the compiler can tell, so a plain unrolled loop kept one row in flight per thread (3.9 TB/s)
  constexpr int RB = 4;
  const int64_t r1 = (RY == 1 || active) ? p.N : 0;
  for (int64_t rb = (int64_t(blockIdx.y) * RY + ry) * RB; rb < r1; rb += int64_t(gridDim.y) * RY * RB) {
    float gs[RB], gd[RB];
    float4 w[RB], t[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int64_t r = rb + u < r1 ? rb + u : r1 - 1;      // tail rows re-read the last row and are not stored
      gs[u] = __ldg(p.g_s_src + r * p.H + h);
      gd[u] = __ldg(p.g_s_dst + r * p.H + h);
      w[u] = ldg4(p.wh + r * p.Dp + c);
      t[u] = *reinterpret_cast<const float4*>(p.g_t + r * p.Dp + c);
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int64_t r = rb + u;
      if (r >= r1) break;
      t[u].x += gs[u] * a1c.x + gd[u] * a2c.x;
      t[u].y += gs[u] * a1c.y + gd[u] * a2c.y;
      t[u].z += gs[u] * a1c.z + gd[u] * a2c.z;
      t[u].w += gs[u] * a1c.w + gd[u] * a2c.w;
      if (PLANES) {
        __align__(8) __half2 hv[2];
        __align__(8) __half2 lv[2];
        split_half2(t[u].x * scale, t[u].y * scale, hv[0], lv[0]);
        split_half2(t[u].z * scale, t[u].w * scale, hv[1], lv[1]);
        *reinterpret_cast<uint2*>(p.hi + r * p.ldp + c) = *reinterpret_cast<const uint2*>(hv);
        *reinterpret_cast<uint2*>(p.lo + r * p.ldp + c) = *reinterpret_cast<const uint2*>(lv);
      } else {
        *reinterpret_cast<float4*>(p.g_t + r * p.Dp + c) = t[u];
      }
      sbw.x += t[u].x; sbw.y += t[u].y; sbw.z += t[u].z; sbw.w += t[u].w;
      sa1.x = fmaf(gs[u], w[u].x, sa1.x); sa1.y = fmaf(gs[u], w[u].y, sa1.y);
      sa1.z = fmaf(gs[u], w[u].z, sa1.z); sa1.w = fmaf(gs[u], w[u].w, sa1.w);
      sa2.x = fmaf(gd[u], w[u].x, sa2.x); sa2.y = fmaf(gd[u], w[u].y, sa2.y);
      sa2.z = fmaf(gd[u], w[u].z, sa2.z); sa2.w = fmaf(gd[u], w[u].w, sa2.w);
      sb1 += gs[u];
      sb2 += gd[u];
    }
  }
  if constexpr (RY == 1) {
    const float bw4[4] = {sbw.x, sbw.y, sbw.z, sbw.w}, a14[4] = {sa1.x, sa1.y, sa1.z, sa1.w}, a24[4] = {sa2.x, sa2.y, sa2.z, sa2.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      atomicAdd(p.g_bw + c + u, bw4[u]);
      atomicAdd(p.g_a1 + c + u, a14[u]);
      atomicAdd(p.g_a2 + c + u, a24[u]);
    }
    if (head_lead) {
      atomicAdd(p.g_b1 + h, sb1);
      atomicAdd(p.g_b2 + h, sb2);
    }
    return;
  }
  float v[14] = {sbw.x, sbw.y, sbw.z, sbw.w, sa1.x, sa1.y, sa1.z, sa1.w, sa2.x, sa2.y, sa2.z, sa2.w, sb1, sb2};
  if constexpr (RY > 1) {                                 // combine the row lanes before the atomics
    __shared__ float red[14][256 + 1];
#pragma unroll
    for (int k = 0; k < 14; ++k) red[k][threadIdx.x] = v[k];
    __syncthreads();
    if (ry != 0) return;
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      float acc = v[k];
      for (int y = 1; y < RY; ++y) acc += red[k][y * TX + tx];
      v[k] = acc;
    }
  }
  if (!active) return;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    atomicAdd(p.g_bw + c + u, v[u]);
    atomicAdd(p.g_a1 + c + u, v[4 + u]);
    atomicAdd(p.g_a2 + c + u, v[8 + u]);
  }
  if (head_lead) {
    atomicAdd(p.g_b1 + h, v[12]);
    atomicAdd(p.g_b2 + h, v[13]);
  }
}

// ---- second pass of B200GAT_LOGIT_HEAD_SOFTMAX: e_h = softmax_h(z_h) couples the heads of an edge,
//     dz_h = e_h (de_h - sum_h' e_h' de_h'),   g_s_src[j,h] += dz_h,   g_s_dst[i,h] += dz_h
// one thread per CSC entry (its source row j by binary search in colptr: robust to any degree skew), all heads in a loop.
__global__ void __launch_bounds__(256) head_softmax_bwd_kernel(const EdgeBwdParams p, int64_t num_entries) {
  const int H = p.H;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < num_entries; k += int64_t(gridDim.x) * blockDim.x) {
    int64_t lo = 0, hi = p.N;                             // largest j with colptr[j] <= k
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(p.colptr + mid) <= k) lo = mid; else hi = mid;
    }
    const int64_t j = lo;
    const int64_t i = __ldg(p.crow + k);
    const float* sd = reinterpret_cast<const float*>(p.rowrec + i * H);     // .x of every 16-byte record: stride 4
    const float* ss = p.s_src + j * H;
    const float* de = p.de + k * H;
    float mx = -INFINITY;
    for (int t = 0; t < H; ++t) mx = fmaxf(mx, __ldg(sd + 4 * t) + __ldg(ss + t));
    float sum = 0.f, dot = 0.f;
    for (int t = 0; t < H; ++t) {
      const float v = expf(__ldg(sd + 4 * t) + __ldg(ss + t) - mx);
      sum += v;
      dot = fmaf(v, de[t], dot);
    }
    const float inv = 1.f / sum;
    const float s = dot * inv;                            // sum_h e_h de_h
    for (int t = 0; t < H; ++t) {
      const float e = expf(__ldg(sd + 4 * t) + __ldg(ss + t) - mx) * inv;
      const float dz = e * (de[t] - s);
      atomicAdd(p.g_s_src + j * H + t, dz);
      atomicAdd(p.g_s_dst + i * H + t, dz);
    }
  }
}

template <int G, int NV>
static int launch_edge_bwd(const EdgeBwdParams& p, bool streaming, cudaStream_t stream) {
  constexpr int GPW = 32 / G;
  const int threads = 256;
  const int64_t want = ceil_div(ceil_div(p.items, GPW), threads / 32);
  const int64_t cap = int64_t(sm_count()) * 8;
  const int blocks = static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
  // measured on the power-law graph: the batched builds (launch_bounds (256, 2) and (256, 3)) are SLOWER here (48.2 /
  // 48.3 vs 42.6 ms): the per-batch dot-product reduce chain wants warps, not loads in flight
  (void)streaming;
  const bool hub = p.nhub > 0;
  if (p.g16) {
    if (hub) edge_bwd16_kernel<G, NV, true><<<blocks, threads, 0, stream>>>(p);
    else edge_bwd16_kernel<G, NV, false><<<blocks, threads, 0, stream>>>(p);
  }
  else if (p.act != B200GAT_LOGIT_LEAKY_RELU) edge_bwd_act_kernel<G, NV><<<blocks, threads, 0, stream>>>(p);
  else if (p.drop.active() && hub) edge_bwd_kernel<G, NV, true, true><<<blocks, threads, 0, stream>>>(p);
  else if (p.drop.active()) edge_bwd_kernel<G, NV, true, false><<<blocks, threads, 0, stream>>>(p);
  else if (hub) edge_bwd_kernel<G, NV, false, true><<<blocks, threads, 0, stream>>>(p);
  else edge_bwd_kernel<G, NV, false, false><<<blocks, threads, 0, stream>>>(p);
  return check_launch("edge_bwd_kernel");
}

struct BwdWorkspace { size_t off_rec, off_gsrc, off_gdst, off_amax, off_gp, total; };

static BwdWorkspace plan_bwd(const b200gat_layer& L, int64_t N) {
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  BwdWorkspace w;
  const size_t nh = up(size_t(N > 0 ? N : 1) * L.heads * sizeof(float));
  w.off_rec = 0;
  w.off_gsrc = 4 * nh;
  w.off_gdst = 5 * nh;
  w.off_amax = 6 * nh;
  w.off_gp = (6 * nh + 256 + (size_t(2) << 20) - 1) / (size_t(2) << 20) * (size_t(2) << 20);   // 2 MB aligned
  w.total = w.off_gp + up(size_t(N > 0 ? N : 1) * L.heads * L.c_pad * sizeof(float));
  return w;
}

// ---- the three stages, shared by the single-GPU composite and the staged (row-partitioned) entry points -------------
struct Geom { int H, C, Cp; int64_t Dp, d_out; bool concat_like; };
static Geom geom_of(const b200gat_layer& L) {
  Geom g;
  g.H = static_cast<int>(L.heads); g.C = static_cast<int>(L.out_channels); g.Cp = static_cast<int>(L.c_pad);
  g.Dp = int64_t(g.H) * g.Cp; g.concat_like = L.concat || g.H == 1; g.d_out = L.concat ? int64_t(g.H) * g.C : g.C;
  return g;
}

// can the CSC pass gather the upstream gradient rows directly (no padded / scaled copy)?
static bool gout_direct(const Geom& g, const float* gout, int64_t ldgo) {
  return g.concat_like && g.C % 4 == 0 && ldgo % 4 == 0 && aligned16(gout);
}

// stage 1: row records + Drow (+ padded / activation-scaled G copy when gp != nullptr) + g_bias column sums
static int run_prep(const b200gat_layer& L, int64_t rows, const float* gout, int64_t ldgo, const float* out, int64_t ldo,
                    const float* o_heads, const float* bias, const float* s_dst, const float* rowmax, const float* rowsum,
                    float4* rowrec, float* gp, float* g_bias, int act, cudaStream_t stream, bool gp16 = false) {
  // gp16: the gatherable copy `gp` is written as bf16 (same element layout, half the bytes) — always, activation or not
  const Geom g = geom_of(L);
  B200GAT_REQUIRE(!gp16 || gp, B200GAT_E_NULL, "edge_bwd: the bf16 gradient rows need the copy buffer");
  const int64_t cap = int64_t(sm_count()) * 8;
  B200GAT_REQUIRE(!act || (g.concat_like && gp), B200GAT_E_UNSUPPORTED,
                  "edge_bwd: out_activation needs a concat-like layer (concat or one head) and the G copy buffer");
  cudaError_t ce = cudaMemsetAsync(g_bias, 0, g.d_out * sizeof(float), stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "edge_bwd: memset: %s", cudaGetErrorString(ce));
  const bool vec = ldo % 4 == 0 && aligned16(out) && aligned16(bias);
  const int Q = g.C / 4;
  const bool fast = g.concat_like && g.C % 4 == 0 && vec && gout_direct(g, gout, ldgo) && g.d_out <= 1024 &&
                    ((Q <= 32 && (Q & (Q - 1)) == 0) || Q % 32 == 0) && (act || gp == nullptr || gp16);
  if (fast) {
    PrepRowsParams pr;
    pr.N = rows; pr.H = g.H; pr.C = g.C; pr.D = static_cast<int>(g.d_out);
    pr.gout = gout; pr.ldgo = ldgo; pr.out = out; pr.ldo = ldo; pr.bias = bias;
    pr.s_dst = s_dst; pr.rowmax = rowmax; pr.rowsum = rowsum; pr.gp = gp; pr.gp16 = gp16 ? 1 : 0; pr.rowrec = rowrec; pr.g_bias = g_bias;
    const int64_t cap_rows = int64_t(sm_count()) * 4;       // few, fat CTAs: one g_bias atomic per column per CTA
    const int DQ = static_cast<int>(g.d_out / 4);
    if (DQ <= 32 && (Q & (Q - 1)) == 0) {                   // narrow rows: a lane group per row instead of a warp
      const int64_t rows_per_cta = DQ <= 4 ? 64 : (DQ <= 8 ? 32 : (DQ <= 16 ? 16 : 8));
      const int64_t want_n = ceil_div(rows, rows_per_cta * 4);
      const int blocks_n = static_cast<int>(want_n < cap_rows ? (want_n > 0 ? want_n : 1) : cap_rows);
      if (DQ <= 4) launch_prep_narrow<4>(pr, act != 0, blocks_n, stream);
      else if (DQ <= 8) launch_prep_narrow<8>(pr, act != 0, blocks_n, stream);
      else if (DQ <= 16) launch_prep_narrow<16>(pr, act != 0, blocks_n, stream);
      else launch_prep_narrow<32>(pr, act != 0, blocks_n, stream);
      return check_launch("bwd_prep_rows_narrow_kernel");
    }
    const int txs = DQ <= 64 ? 6 : (DQ <= 128 ? 7 : 8);
    const int64_t want_w = ceil_div(rows, int64_t(256 >> txs) * 4 * 4);     // at least 4 batches per thread
    const int blocks_w = static_cast<int>(want_w < cap_rows ? (want_w > 0 ? want_w : 1) : cap_rows);
    if (txs == 6) { if (act) bwd_prep_rows_wide_kernel<true, 6><<<blocks_w, 256, 0, stream>>>(pr); else bwd_prep_rows_wide_kernel<false, 6><<<blocks_w, 256, 0, stream>>>(pr); }
    else if (txs == 7) { if (act) bwd_prep_rows_wide_kernel<true, 7><<<blocks_w, 256, 0, stream>>>(pr); else bwd_prep_rows_wide_kernel<false, 7><<<blocks_w, 256, 0, stream>>>(pr); }
    else { if (act) bwd_prep_rows_wide_kernel<true, 8><<<blocks_w, 256, 0, stream>>>(pr); else bwd_prep_rows_wide_kernel<false, 8><<<blocks_w, 256, 0, stream>>>(pr); }
    return check_launch("bwd_prep_rows_wide_kernel");
  }
  if (!g.concat_like && gp && aligned16(o_heads) && aligned16(gp)) {
    PrepMeanParams pm;
    pm.N = rows; pm.H = g.H; pm.C = g.C; pm.Cp = g.Cp;
    pm.gout = gout; pm.ldgo = ldgo; pm.o_heads = o_heads; pm.s_dst = s_dst; pm.rowmax = rowmax; pm.rowsum = rowsum;
    pm.gp = gp; pm.gp16 = gp16 ? 1 : 0; pm.rowrec = rowrec; pm.g_bias = g_bias;
    const int64_t cap_rows = int64_t(sm_count()) * 8;
    const int Qm = g.Cp / 4;
    if (Qm <= 32) {                                        // lane group per row
      const int G = Qm <= 4 ? 4 : (Qm <= 8 ? 8 : (Qm <= 16 ? 16 : 32));
      const int64_t want = ceil_div(rows, int64_t(8) * (32 / G) * 2);
      const int blocks = static_cast<int>(want < cap_rows ? (want > 0 ? want : 1) : cap_rows);
      if (G == 4) bwd_prep_mean_narrow_kernel<4><<<blocks, 256, 0, stream>>>(pm);
      else if (G == 8) bwd_prep_mean_narrow_kernel<8><<<blocks, 256, 0, stream>>>(pm);
      else if (G == 16) bwd_prep_mean_narrow_kernel<16><<<blocks, 256, 0, stream>>>(pm);
      else bwd_prep_mean_narrow_kernel<32><<<blocks, 256, 0, stream>>>(pm);
      return check_launch("bwd_prep_mean_narrow_kernel");
    }
    const int64_t want = ceil_div(rows, 8);
    bwd_prep_mean_rows_kernel<<<static_cast<int>(want < cap_rows ? want : cap_rows), 256, 0, stream>>>(pm);
    return check_launch("bwd_prep_mean_rows_kernel");
  }
  PrepParams dp;
  dp.N = rows; dp.H = g.H; dp.C = g.C; dp.Cp = g.Cp; dp.concat_like = g.concat_like ? 1 : 0;
  dp.vec = (gp == nullptr && vec) ? 1 : 0;
  B200GAT_REQUIRE(gp != nullptr || (gout_direct(g, gout, ldgo) && dp.vec), B200GAT_E_ALIGN,
                  "edge_bwd: the upstream gradient is not directly gatherable: a padded copy buffer is required");
  dp.gout = gout; dp.ldgo = ldgo; dp.out = out; dp.ldo = ldo; dp.o_heads = o_heads; dp.bias = bias;
  dp.s_dst = s_dst; dp.rowmax = rowmax; dp.rowsum = rowsum;
  dp.gscale = g.concat_like ? 1.f : 1.f / static_cast<float>(g.H);
  dp.act = act;
  dp.gp = gp;
  dp.gp16 = gp16 ? 1 : 0;
  dp.ldgp = g.concat_like ? g.Dp : g.Cp;
  dp.rowrec = rowrec;
  const int64_t want = ceil_div(rows * g.H, 8);
  bwd_prep_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, stream>>>(dp);
  int rc = check_launch("bwd_prep_kernel");
  if (rc) return rc;
  const int64_t ysplit = ceil_div(rows, 8) < 64 ? ceil_div(rows, 8) : 64;
  dim3 grid(static_cast<unsigned>(ceil_div(g.d_out, 32)), static_cast<unsigned>(ysplit));
  B200GAT_REQUIRE(!(act && gp16), B200GAT_E_UNSUPPORTED,
                  "edge_bwd: bf16 gradient rows with a deferred activation need a head width that is a multiple of 4");
  if (act) colsum_kernel<<<grid, 256, 0, stream>>>(gp, dp.ldgp, rows, static_cast<int>(g.d_out), g.C, g.Cp, g_bias);
  else colsum_kernel<<<grid, 256, 0, stream>>>(gout, ldgo, rows, static_cast<int>(g.d_out), 1, 1, g_bias);
  return check_launch("colsum_kernel");
}

// stage 2: the CSC pass over `rows` source rows.  wh / s_src / gwh / g_s_src are indexed by the LOCAL source row;
// crow holds GLOBAL destination ids indexing rowrec / g / g_s_dst (identical spaces on a single GPU).
static int run_csc(const b200gat_layer& L, int64_t rows, const int32_t* colptr, const int32_t* crow, const int32_t* ceid,
                   const float* wh, const float* s_src, const float4* rowrec, const DropoutSpec& drop, const float* gsrc_rows,
                   int64_t ldg, int hs, float* gwh, float* g_s_src, float* g_s_dst, int64_t span, const int32_t* colend,
                   const int32_t* hub, int64_t nhub, int64_t max_deg, uint32_t* amax, cudaStream_t stream,
                   float* de = nullptr, int64_t num_entries = 0, bool rows16 = false) {
  const Geom g = geom_of(L);
  EdgeBwdParams p;
  p.N = rows; p.items = rows * g.H; p.H = g.H; p.Cp = g.Cp; p.Dp = static_cast<int>(g.Dp); p.slope = L.negative_slope; p.act = L.logit_activation;
  p.colptr = colptr; p.crow = crow; p.ceid = ceid;
  p.wh = wh; p.s_src = s_src; p.rowrec = rowrec; p.drop = drop;
  p.g = gsrc_rows; p.ldg = ldg; p.hs = hs;
  p.g16 = rows16 ? static_cast<const void*>(gsrc_rows) : nullptr;      // bf16 rows: same element layout
  B200GAT_REQUIRE(!rows16 || (!drop.active() && p.act == B200GAT_LOGIT_LEAKY_RELU), B200GAT_E_UNSUPPORTED,
                  "edge_bwd: bf16 gradient rows are offered without dropout and with LeakyReLU logits only");
  p.gwh = gwh; p.g_s_src = g_s_src; p.g_s_dst = g_s_dst;
  B200GAT_REQUIRE(nhub >= 0 && (nhub == 0 || (hub && colend)), B200GAT_E_NULL, "edge_bwd: hub_cols / colend missing");
  p.hub = hub; p.nhub = nhub; p.colend = nhub > 0 ? colend : colptr + 1;
  p.max_deg = static_cast<int>(max_deg);
  p.amax = amax;
  p.de = de;
  B200GAT_REQUIRE(p.act != B200GAT_LOGIT_HEAD_SOFTMAX || de, B200GAT_E_UNSUPPORTED,
                  "edge_bwd: the across-heads softmax logit activation needs the per-edge scratch buffer (b200gat_edge_bwd only)");
  const int Q = g.Cp / 4;
  const bool streaming = edge_schedule_streaming(span, ldg * 4);
  int rc;
  if (edge_bwd_mean_supported(g.H, Q, hs, p.act)) {
    switch (g.H) {
      case 2: rc = dispatch_edge_bwd_mean<2>(p, Q, stream); break;
      case 3: rc = dispatch_edge_bwd_mean<3>(p, Q, stream); break;
      case 4: rc = dispatch_edge_bwd_mean<4>(p, Q, stream); break;
      case 6: rc = dispatch_edge_bwd_mean<6>(p, Q, stream); break;
      default: rc = dispatch_edge_bwd_mean<8>(p, Q, stream); break;
    }
  }
  else if (Q <= 1) rc = launch_edge_bwd<1, 1>(p, streaming, stream);
  else if (Q <= 2) rc = launch_edge_bwd<2, 1>(p, streaming, stream);
  else if (Q <= 4) rc = launch_edge_bwd<4, 1>(p, streaming, stream);
  else if (Q <= 8) rc = launch_edge_bwd<8, 1>(p, streaming, stream);
  else if (Q <= 16) rc = launch_edge_bwd<16, 1>(p, streaming, stream);
  else if (Q <= 32) rc = launch_edge_bwd<32, 1>(p, streaming, stream);
  else if (Q <= 64) rc = launch_edge_bwd<32, 2>(p, streaming, stream);
  else rc = launch_edge_bwd<32, 4>(p, streaming, stream);
  if (rc) return rc;
  if ((rc = launch_edge_bwd_hub(p, stream))) return rc;
  if (p.act == B200GAT_LOGIT_HEAD_SOFTMAX && num_entries > 0) {
    const int64_t want = ceil_div(num_entries, 256), cap = int64_t(sm_count()) * 8;
    head_softmax_bwd_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, stream>>>(p, num_entries);
    rc = check_launch("head_softmax_bwd_kernel");
  }
  return rc;
}

// stage 3: gT (fp32 in place, or as the operand split `gsplit` when given) + parameter column sums over `rows` rows
// (outputs are overwritten, not accumulated).  amax: the 5 slots described at FinishParams (planes mode only).
static int run_finish(const b200gat_layer& L, int64_t rows, const float* wh, const float* a1, const float* a2,
                      const float* g_s_src, const float* g_s_dst, float* g_t, float* g_bw, float* g_a1, float* g_a2,
                      float* g_b1, float* g_b2, void* gsplit, uint32_t* amax, bool amax_from_csc, cudaStream_t stream) {
  const Geom g = geom_of(L);
  cudaError_t ce = cudaMemsetAsync(g_bw, 0, g.Dp * sizeof(float), stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(g_a1, 0, g.Dp * sizeof(float), stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(g_a2, 0, g.Dp * sizeof(float), stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(g_b1, 0, g.H * sizeof(float), stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(g_b2, 0, g.H * sizeof(float), stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "edge_bwd: memset: %s", cudaGetErrorString(ce));
  if (rows == 0) return 0;
  FinishParams f{};
  f.N = rows; f.H = g.H; f.Cp = g.Cp; f.Dp = static_cast<int>(g.Dp);
  f.wh = wh; f.a1 = a1; f.a2 = a2; f.g_s_src = g_s_src; f.g_s_dst = g_s_dst; f.g_t = g_t;
  f.g_bw = g_bw; f.g_a1 = g_a1; f.g_a2 = g_a2; f.g_b1 = g_b1; f.g_b2 = g_b2;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (gsplit) {
    const int64_t n_nh = rows * g.H;
    const int64_t want = ceil_div(rows * g.Dp, 256 * 16);
    // max|gWh| (and max|g_s_src|) came out of the CSC pass when amax_from_csc: only the small [N, H] / [Dp] arrays are read
    const int64_t want_b = amax_from_csc ? ceil_div(n_nh, 256 * 4) : want;
    gt_amax_kernel<<<static_cast<int>(want_b < cap ? (want_b > 0 ? want_b : 1) : cap), 256, 0, stream>>>(
        g_t, amax_from_csc ? 0 : rows * g.Dp, g_s_src, g_s_dst, n_nh, a1, a2, f.Dp, amax);
    int rc = check_launch("gt_amax_kernel");
    if (rc) return rc;
    const Blob B = make_blob(gsplit, rows, g.Dp);
    f.amax = amax; f.hi = B.hi(); f.lo = B.lo(); f.ldp = B.ldp; f.inv_scale = B.inv_scale(); f.bound_bits = B.amax_bits();
    if (B.ldp != g.Dp) {   // pad columns of the planes must be zero
      ce = cudaMemsetAsync(B.base + BLOB_HEADER, 0, 2 * plane_bytes(rows, g.Dp), stream);
      if (ce != cudaSuccess) return fail(static_cast<int>(ce), "edge_bwd: memset: %s", cudaGetErrorString(ce));
    }
  }
  const int tx_shift = g.Dp <= 64 ? 4 : (g.Dp <= 256 ? 6 : (g.Dp <= 512 ? 7 : 8));   // TX = 16 / 64 / 128 / 256 column groups per CTA
  const int TX = 1 << tx_shift, RY = 256 >> tx_shift;
  const int xblocks = static_cast<int>(ceil_div(g.Dp, 4 * TX));
  // one resident wave (4 CTAs / SM): every CTA ends in a column-sum atomic tail, so a second wave only adds tail
  // (PPI step 5.12 -> 5.05 ms against 8 CTAs / SM; 2 and 16 per SM are both slower; the 2.4 M-node graph is unchanged)
  int64_t ysplit = ceil_div(int64_t(sm_count()) * 4, xblocks);
  const int64_t max_y = ceil_div(rows, int64_t(16) * RY);      // at least 4 batches of 4 rows per thread
  if (ysplit > max_y) ysplit = max_y;
  if (ysplit < 1) ysplit = 1;
  dim3 grid(xblocks, static_cast<unsigned>(ysplit));
  if (gsplit) {
    if (tx_shift == 4) bwd_finish_kernel<true, 4><<<grid, 256, 0, stream>>>(f);
    else if (tx_shift == 6) bwd_finish_kernel<true, 6><<<grid, 256, 0, stream>>>(f);
    else if (tx_shift == 7) bwd_finish_kernel<true, 7><<<grid, 256, 0, stream>>>(f);
    else bwd_finish_kernel<true, 8><<<grid, 256, 0, stream>>>(f);
  } else {
    if (tx_shift == 4) bwd_finish_kernel<false, 4><<<grid, 256, 0, stream>>>(f);
    else if (tx_shift == 6) bwd_finish_kernel<false, 6><<<grid, 256, 0, stream>>>(f);
    else if (tx_shift == 7) bwd_finish_kernel<false, 7><<<grid, 256, 0, stream>>>(f);
    else bwd_finish_kernel<false, 8><<<grid, 256, 0, stream>>>(f);
  }
  return check_launch("bwd_finish_kernel");
}

}  // namespace b200gat

using namespace b200gat;

extern "C" size_t b200gat_edge_bwd_workspace_bytes(const b200gat_layer* L, int64_t N) {
  if (!L || N < 0) return 0;
  return plan_bwd(*L, N).total;
}

extern "C" size_t b200gat_edge_bwd_split_bytes(const b200gat_layer* L, int64_t N) {
  if (!L || N < 0 || !proj_tc_bwd_supported(*L, N)) return 0;
  return blob_bytes(N, L->heads * L->c_pad);
}

extern "C" int b200gat_edge_bwd(const b200gat_edge_bwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "edge_bwd: NULL args");
  int rc = validate_layer(a->layer);
  if (rc) return rc;
  if ((rc = validate_graph(a->graph))) return rc;
  const b200gat_layer& L = a->layer;
  const int64_t N = a->graph.num_nodes;
  const Geom g = geom_of(L);
  B200GAT_REQUIRE(a->g_bw && a->g_a1 && a->g_a2 && a->g_b1 && a->g_b2 && (a->g_bias || a->rowrec_in), B200GAT_E_NULL,
                  "edge_bwd: NULL parameter-gradient pointer");
  if (N == 0) {
    cudaError_t ce = a->g_bias ? cudaMemsetAsync(a->g_bias, 0, g.d_out * sizeof(float), stream) : cudaSuccess;
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "edge_bwd: memset: %s", cudaGetErrorString(ce));
    return run_finish(L, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, a->g_bw, a->g_a1, a->g_a2, a->g_b1, a->g_b2,
                      nullptr, nullptr, false, stream);
  }
  B200GAT_REQUIRE(a->gout && a->wh && a->s_src && a->s_dst && a->rowmax && a->rowsum && a->a1 && a->a2 && a->g_t &&
                  a->bias && a->workspace, B200GAT_E_NULL, "edge_bwd: NULL pointer");
  B200GAT_REQUIRE(g.concat_like ? (a->out != nullptr) : (a->o_heads != nullptr), B200GAT_E_NULL,
                  "edge_bwd: forward output (out / o_heads) missing");
  DropoutSpec drop;
  if ((rc = make_dropout(a->mask, a->dropout, &drop))) return rc;
  B200GAT_REQUIRE(!drop.active() || a->graph.ceid, B200GAT_E_NULL, "edge_bwd: dropout needs graph.ceid");
  B200GAT_REQUIRE(a->ldgo >= g.d_out && (!g.concat_like || a->ldo >= g.d_out), B200GAT_E_SHAPE, "edge_bwd: leading dimension < D_out");
  B200GAT_REQUIRE(aligned16(a->wh) && aligned16(a->g_t) && aligned16(a->a1) && aligned16(a->a2), B200GAT_E_ALIGN,
                  "edge_bwd: wh / g_t / a1 / a2 must be 16-byte aligned");
  B200GAT_REQUIRE(L.logit_activation != B200GAT_LOGIT_HEAD_SOFTMAX ||
                  (a->edge_scratch && a->edge_scratch_bytes >= size_t(a->graph.num_edges) * g.H * sizeof(float)),
                  B200GAT_E_WORKSPACE, "edge_bwd: edge_scratch of E' * H * 4 bytes is required for the across-heads softmax logits");
  const BwdWorkspace w = plan_bwd(L, N);
  B200GAT_REQUIRE(a->workspace_bytes >= w.total, B200GAT_E_WORKSPACE, "edge_bwd: workspace %zu < %zu bytes",
                  a->workspace_bytes, w.total);
  B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255u) == 0, B200GAT_E_ALIGN,
                  "edge_bwd: workspace must be 256-byte aligned");
  char* base = static_cast<char*>(a->workspace);
  float4* rowrec = reinterpret_cast<float4*>(base + w.off_rec);
  float* g_s_src = reinterpret_cast<float*>(base + w.off_gsrc);
  float* g_s_dst = reinterpret_cast<float*>(base + w.off_gdst);
  float* gp = reinterpret_cast<float*>(base + w.off_gp);
  uint32_t* amax = reinterpret_cast<uint32_t*>(base + w.off_amax);
  // zero g_s_dst (accumulated atomically) and, right behind it, the amax slots
  cudaError_t ce = cudaMemsetAsync(g_s_dst, 0, w.off_gp - w.off_gdst, stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "edge_bwd: memset: %s", cudaGetErrorString(ce));
  // rowrec_in: the prep pass already ran inside the consuming layer's gX GEMM — gout IS G (activation applied)
  const bool prep_done = a->rowrec_in != nullptr;
  const int act = prep_done ? ACT_NONE : a->out_activation;
  B200GAT_REQUIRE(act == ACT_NONE || act == ACT_ELU, B200GAT_E_UNSUPPORTED, "edge_bwd: unknown out_activation %d", act);
  if (prep_done) {
    B200GAT_REQUIRE(aligned16(a->rowrec_in) && gout_direct(g, a->gout, a->ldgo) && !a->gather_bf16, B200GAT_E_ALIGN,
                    "edge_bwd: rowrec_in needs a directly gatherable fp32 gout (concat-like layer, C %% 4 == 0, aligned rows)");
    rowrec = const_cast<float4*>(reinterpret_cast<const float4*>(a->rowrec_in));
  }
  void* gsplit = a->g_t_split;
  if (gsplit) {
    const size_t need = blob_bytes(N, g.Dp);
    B200GAT_REQUIRE(a->g_t_split_bytes >= need, B200GAT_E_WORKSPACE, "edge_bwd: g_t_split %zu < %zu bytes", a->g_t_split_bytes, need);
    B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(gsplit) & 255u) == 0, B200GAT_E_ALIGN, "edge_bwd: g_t_split must be 256-byte aligned");
  }
  // with an output activation the gathered rows are gout * ELU'(out): always the copy
  const bool rows16 = a->gather_bf16 != 0;
  const bool direct = prep_done ||
                      (!act && !rows16 && gout_direct(g, a->gout, a->ldgo) && (!g.concat_like || (a->ldo % 4 == 0 && aligned16(a->out))) &&
                       aligned16(a->bias));
  if (!prep_done &&
      (rc = run_prep(L, N, a->gout, a->ldgo, a->out, a->ldo, a->o_heads, a->bias, a->s_dst, a->rowmax, a->rowsum, rowrec,
                     direct ? nullptr : gp, a->g_bias, act, stream, rows16)))
    return rc;
  const float* grows = direct ? a->gout : gp;
  const int64_t ldg = direct ? a->ldgo : (g.concat_like ? g.Dp : g.Cp);
  const int hs = direct ? g.C : (g.concat_like ? g.Cp : 0);
  if ((rc = run_csc(L, N, a->graph.colptr, a->graph.crow, a->graph.ceid, a->wh, a->s_src, rowrec, drop, grows, ldg, hs,
                    a->g_t, g_s_src, g_s_dst, a->graph.span, a->graph.colend, a->graph.hub_cols, a->graph.num_hub_cols,
                    a->graph.max_out_degree, gsplit ? amax : nullptr, stream, a->edge_scratch, a->graph.num_edges, rows16)))
    return rc;
  // giant source rows are accumulated from per-segment partial sums: their maxima are not the maxima of the sums
  // (and with the across-heads softmax g_s_src is only complete after the second pass)
  const bool amax_from_csc = gsplit && a->graph.max_out_degree <= B200GAT_GIANT_DEGREE && L.logit_activation != B200GAT_LOGIT_HEAD_SOFTMAX;
  return run_finish(L, N, a->wh, a->a1, a->a2, g_s_src, g_s_dst, a->g_t, a->g_bw, a->g_a1, a->g_a2, a->g_b1, a->g_b2,
                    gsplit, amax, amax_from_csc, stream);
}

// ---- staged entry points (destination-row partitioned multi-GPU execution: the caller runs the collectives between
//      the stages; see include/b200gat.h) ------------------------------------------------------------------------------
extern "C" int b200gat_edge_bwd_prep(const b200gat_edge_bwd_prep_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "edge_bwd_prep: NULL args");
  int rc = validate_layer(a->layer);
  if (rc) return rc;
  const Geom g = geom_of(a->layer);
  B200GAT_REQUIRE(a->num_rows >= 0, B200GAT_E_SHAPE, "edge_bwd_prep: negative num_rows");
  B200GAT_REQUIRE(a->g_bias, B200GAT_E_NULL, "edge_bwd_prep: NULL g_bias");
  if (a->num_rows == 0) {
    cudaError_t ce = cudaMemsetAsync(a->g_bias, 0, g.d_out * sizeof(float), stream);
    return ce == cudaSuccess ? 0 : fail(static_cast<int>(ce), "edge_bwd_prep: memset: %s", cudaGetErrorString(ce));
  }
  B200GAT_REQUIRE(a->gout && a->bias && a->s_dst && a->rowmax && a->rowsum && a->rowrec, B200GAT_E_NULL, "edge_bwd_prep: NULL pointer");
  B200GAT_REQUIRE(g.concat_like ? (a->out != nullptr) : (a->o_heads != nullptr), B200GAT_E_NULL,
                  "edge_bwd_prep: forward output (out / o_heads) missing");
  B200GAT_REQUIRE(a->ldgo >= g.d_out && (!g.concat_like || a->ldo >= g.d_out), B200GAT_E_SHAPE, "edge_bwd_prep: leading dimension < D_out");
  B200GAT_REQUIRE(aligned16(a->rowrec), B200GAT_E_ALIGN, "edge_bwd_prep: rowrec must be 16-byte aligned");
  B200GAT_REQUIRE(a->out_activation == ACT_NONE || a->out_activation == ACT_ELU, B200GAT_E_UNSUPPORTED,
                  "edge_bwd_prep: unknown out_activation %d", a->out_activation);
  B200GAT_REQUIRE(!(a->g_pad && a->g_pad_bf16), B200GAT_E_SHAPE, "edge_bwd_prep: give g_pad OR g_pad_bf16");
  return run_prep(a->layer, a->num_rows, a->gout, a->ldgo, a->out, a->ldo, a->o_heads, a->bias, a->s_dst, a->rowmax, a->rowsum,
                  reinterpret_cast<float4*>(a->rowrec), a->g_pad_bf16 ? static_cast<float*>(a->g_pad_bf16) : a->g_pad, a->g_bias,
                  a->out_activation, stream, a->g_pad_bf16 != nullptr);
}

extern "C" int b200gat_edge_bwd_csc(const b200gat_edge_bwd_csc_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "edge_bwd_csc: NULL args");
  int rc = validate_layer(a->layer);
  if (rc) return rc;
  B200GAT_REQUIRE(a->num_rows >= 0, B200GAT_E_SHAPE, "edge_bwd_csc: negative num_rows");
  if (a->num_rows == 0) return 0;
  const float* grows = a->g_bf16 ? static_cast<const float*>(a->g_bf16) : a->g;
  B200GAT_REQUIRE(a->colptr && a->crow && a->wh && a->s_src && a->rowrec && grows && a->g_wh && a->g_s_src && a->g_s_dst,
                  B200GAT_E_NULL, "edge_bwd_csc: NULL pointer");
  DropoutSpec drop;
  if ((rc = make_dropout(a->mask, a->dropout, &drop))) return rc;
  B200GAT_REQUIRE(!drop.active() || a->ceid, B200GAT_E_NULL, "edge_bwd_csc: dropout needs ceid");
  B200GAT_REQUIRE(aligned16(a->wh) && aligned16(a->g_wh) && aligned16(grows) && aligned16(a->rowrec) && a->ldg % 4 == 0 &&
                  a->g_head_stride % 4 == 0, B200GAT_E_ALIGN, "edge_bwd_csc: wh / g / g_wh / rowrec must be 16-byte aligned");
  return run_csc(a->layer, a->num_rows, a->colptr, a->crow, a->ceid, a->wh, a->s_src,
                 reinterpret_cast<const float4*>(a->rowrec), drop, grows, a->ldg, static_cast<int>(a->g_head_stride),
                 a->g_wh, a->g_s_src, a->g_s_dst, a->span, a->colend, a->hub_cols, a->num_hub_cols, a->max_out_degree, nullptr, stream,
                 nullptr, 0, a->g_bf16 != nullptr);
}

extern "C" int b200gat_edge_bwd_finish(const b200gat_edge_bwd_finish_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "edge_bwd_finish: NULL args");
  int rc = validate_layer(a->layer);
  if (rc) return rc;
  B200GAT_REQUIRE(a->num_rows >= 0, B200GAT_E_SHAPE, "edge_bwd_finish: negative num_rows");
  B200GAT_REQUIRE(a->g_bw && a->g_a1 && a->g_a2 && a->g_b1 && a->g_b2, B200GAT_E_NULL, "edge_bwd_finish: NULL output");
  B200GAT_REQUIRE(a->num_rows == 0 || (a->wh && a->a1 && a->a2 && a->g_s_src && a->g_s_dst && a->g_t), B200GAT_E_NULL,
                  "edge_bwd_finish: NULL pointer");
  B200GAT_REQUIRE(aligned16(a->wh) && aligned16(a->g_t) && aligned16(a->a1) && aligned16(a->a2), B200GAT_E_ALIGN,
                  "edge_bwd_finish: wh / g_t / a1 / a2 must be 16-byte aligned");
  uint32_t* amax = nullptr;
  if (a->g_t_split && a->num_rows > 0) {
    const Geom g = geom_of(a->layer);
    const size_t need = blob_bytes(a->num_rows, g.Dp);
    B200GAT_REQUIRE(a->g_t_split_bytes >= need, B200GAT_E_WORKSPACE, "edge_bwd_finish: g_t_split %zu < %zu bytes", a->g_t_split_bytes, need);
    B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(a->g_t_split) & 255u) == 0, B200GAT_E_ALIGN, "edge_bwd_finish: g_t_split must be 256-byte aligned");
    // the five max-magnitude slots of the scale bound live in the spare words of the blob header
    amax = reinterpret_cast<uint32_t*>(static_cast<char*>(a->g_t_split) + 16);
    cudaError_t ce = cudaMemsetAsync(amax, 0, 5 * sizeof(uint32_t), stream);
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "edge_bwd_finish: memset: %s", cudaGetErrorString(ce));
  }
  return run_finish(a->layer, a->num_rows, a->wh, a->a1, a->a2, a->g_s_src, a->g_s_dst, a->g_t, a->g_bw, a->g_a1, a->g_a2,
                    a->g_b1, a->g_b2, a->num_rows > 0 ? a->g_t_split : nullptr, amax, false, stream);
}
