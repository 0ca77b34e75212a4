// Graph-level readout head of GATNet (GATNet.py:72-75), fused:
//     x = elu(conv2(...))  ->  pooled = scatter_mean(x, batch)  ->  hid = relu(lin1(pooled))  ->  log_softmax(lin2(hid))
// The reference issues ~12 library kernels forward and ~15 backward for this (index_add x2, clamp, div, two addmm, relu,
// log_softmax and their autograd mirrors) on tensors of a few KB; here: a pooling pass over the nodes (float atomics into
// the [G, F] sums — no assumption that `batch` is sorted) and ONE CTA per graph for the two small layers, and three kernels
// for the backward.  fp32 throughout.  The ELU of the last GAT layer is applied while x is loaded (x_activation), so the
// producing layer hands over its pre-activation output (layer-boundary fusion, include/b200gat.h B200GAT_ACT_*).
#include "common.cuh"
#include "split_blob.cuh"
#include <math.h>

namespace b200gat {

// sums[g, :] += act(x[n, :]) ; counts[g] += 1        (sums / counts zero-initialised)
__global__ void __launch_bounds__(256)
readout_pool_kernel(const float* __restrict__ x, int64_t ldx, const int64_t* __restrict__ batch, int64_t N, int F, int64_t G,
                    int act, float* __restrict__ sums, float* __restrict__ counts, int32_t* __restrict__ status) {
  const int64_t total = N * F;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t n = t / F;
    const int f = static_cast<int>(t - n * F);
    const int64_t g = __ldg(batch + n);
    if (g < 0 || g >= G) {                               // reported lazily (status), never written out of bounds
      if (f == 0) atomicAdd(status, 1);
      continue;
    }
    float v = __ldg(x + n * ldx + f);
    if (act) v = elu_fwd(v);
    atomicAdd(sums + g * F + f, v);
    if (f == 0) atomicAdd(counts + g, 1.f);
  }
}

struct HeadParams {
  int64_t G;
  int F, Hd, K;                                          // in_channels, hidden, classes
  const float* w1; const float* b1; const float* w2; const float* b2;
  float* pooled;                                         // in: sums, out: means [G, F]
  const float* counts;                                   // [G]
  float* hid;                                            // out [G, Hd] (after relu)
  float* logp;                                           // out [G, K]
};

// one CTA per graph; dynamic shared memory: F + Hd + K floats
__global__ void __launch_bounds__(128) readout_head_fwd_kernel(const HeadParams p) {
  extern __shared__ float sm[];
  float* s_pool = sm; float* s_hid = sm + p.F; float* s_log = s_hid + p.Hd;
  __shared__ float s_red[2];
  for (int64_t g = blockIdx.x; g < p.G; g += gridDim.x) {
    const float inv = 1.f / fmaxf(__ldg(p.counts + g), 1.f);        // scatter_mean: empty groups -> 0 (count clamped to 1)
    for (int f = threadIdx.x; f < p.F; f += blockDim.x) {
      const float m = p.pooled[g * p.F + f] * inv;
      s_pool[f] = m;
      p.pooled[g * p.F + f] = m;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.Hd; j += blockDim.x) {
      float a = __ldg(p.b1 + j);
      const float* w = p.w1 + int64_t(j) * p.F;
      for (int f = 0; f < p.F; ++f) a = fmaf(__ldg(w + f), s_pool[f], a);
      a = a > 0.f ? a : 0.f;
      s_hid[j] = a;
      p.hid[g * p.Hd + j] = a;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
      float a = __ldg(p.b2 + k);
      const float* w = p.w2 + int64_t(k) * p.Hd;
      for (int j = 0; j < p.Hd; ++j) a = fmaf(__ldg(w + j), s_hid[j], a);
      s_log[k] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                              // log_softmax over a handful of classes
      float mx = -INFINITY;
      for (int k = 0; k < p.K; ++k) mx = fmaxf(mx, s_log[k]);
      float sum = 0.f;
      for (int k = 0; k < p.K; ++k) sum += expf(s_log[k] - mx);
      s_red[0] = mx; s_red[1] = logf(sum);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) p.logp[g * p.K + k] = s_log[k] - s_red[0] - s_red[1];
    __syncthreads();
  }
}

struct HeadBwdParams {
  int64_t G;
  int F, Hd, K;
  const float* w1; const float* w2;
  const float* counts; const float* hid; const float* logp; const float* g_logp;
  float* g_logits;                                       // out [G, K]
  float* g_hid;                                          // out [G, Hd]
  float* g_pool;                                         // out [G, F]: d loss / d pooled, already divided by the node count
};

// one CTA per graph: log_softmax backward, lin2 backward (data), relu, lin1 backward (data)
__global__ void __launch_bounds__(128) readout_head_bwd_kernel(const HeadBwdParams p) {
  extern __shared__ float sm[];
  float* s_gl = sm; float* s_gh = sm + p.K;
  __shared__ float s_sum;
  for (int64_t g = blockIdx.x; g < p.G; g += gridDim.x) {
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int k = 0; k < p.K; ++k) s += __ldg(p.g_logp + g * p.K + k);
      s_sum = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
      const float v = __ldg(p.g_logp + g * p.K + k) - expf(__ldg(p.logp + g * p.K + k)) * s_sum;
      s_gl[k] = v;
      p.g_logits[g * p.K + k] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.Hd; j += blockDim.x) {
      float a = 0.f;
      for (int k = 0; k < p.K; ++k) a = fmaf(__ldg(p.w2 + int64_t(k) * p.Hd + j), s_gl[k], a);
      a = __ldg(p.hid + g * p.Hd + j) > 0.f ? a : 0.f;
      s_gh[j] = a;
      p.g_hid[g * p.Hd + j] = a;
    }
    __syncthreads();
    const float inv = 1.f / fmaxf(__ldg(p.counts + g), 1.f);
    for (int f = threadIdx.x; f < p.F; f += blockDim.x) {
      float a = 0.f;
      for (int j = 0; j < p.Hd; ++j) a = fmaf(__ldg(p.w1 + int64_t(j) * p.F + f), s_gh[j], a);
      p.g_pool[g * p.F + f] = a * inv;
    }
    __syncthreads();
  }
}

// all four parameter gradients in ONE launch (a first version with one serial loop over the graphs per output element and
// four launches cost 4 x 40 us on a 128-graph batch — 35 % of the whole CIFAR-shaped step under ncu):
//     g_w1[j, f] = sum_g g_hid[g, j] pooled[g, f]     g_b1[j] = sum_g g_hid[g, j]
//     g_w2[k, j] = sum_g g_logits[g, k] hid[g, j]     g_b2[k] = sum_g g_logits[g, k]
// One output element per group of 8 lanes; the lanes split the graphs, 4 independent loads in flight each.
struct WgradParams {
  int64_t G;
  int F, Hd, K;
  const float* g_hid; const float* pooled; const float* g_logits; const float* hid;
  float* g_w1; float* g_b1; float* g_w2; float* g_b2;
};

__global__ void __launch_bounds__(256) readout_wgrad_kernel(const WgradParams p) {
  const int64_t n1 = int64_t(p.Hd) * p.F, n2 = n1 + p.Hd, n3 = n2 + int64_t(p.K) * p.Hd, n4 = n3 + p.K;
  const int sub = threadIdx.x & 7, grp = (threadIdx.x & 31) >> 3;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t base = warp * 4; base < n4; base += nwarps * 4) {      // warp-uniform trip count (full-mask shuffles below)
    const bool valid = base + grp < n4;
    const int64_t t = valid ? base + grp : n4 - 1;
    const float* A; const float* B; int R, C, r, c; float* out;
    if (t < n1) { A = p.g_hid; R = p.Hd; B = p.pooled; C = p.F; r = static_cast<int>(t / p.F); c = static_cast<int>(t - int64_t(r) * p.F); out = p.g_w1 + t; }
    else if (t < n2) { A = p.g_hid; R = p.Hd; B = nullptr; C = 1; r = static_cast<int>(t - n1); c = 0; out = p.g_b1 + r; }
    else if (t < n3) { const int64_t u = t - n2; A = p.g_logits; R = p.K; B = p.hid; C = p.Hd; r = static_cast<int>(u / p.Hd); c = static_cast<int>(u - int64_t(r) * p.Hd); out = p.g_w2 + u; }
    else { A = p.g_logits; R = p.K; B = nullptr; C = 1; r = static_cast<int>(t - n3); c = 0; out = p.g_b2 + r; }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t g0 = sub; g0 < p.G; g0 += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t g = g0 + 8 * u;
        if (g < p.G) acc[u] = fmaf(__ldg(A + g * R + r), B ? __ldg(B + g * C + c) : 1.f, acc[u]);
      }
    }
    float a = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (valid && sub == 0) *out = a;
  }
}

// g_x[n, :] = g_pool[batch[n], :]  — the gradient w.r.t. act(x); the producing layer applies act' (B200GAT_ACT_* contract)
__global__ void __launch_bounds__(256)
readout_scatter_bwd_kernel(const float* __restrict__ g_pool, const int64_t* __restrict__ batch, int64_t N, int F, int64_t G,
                           float* __restrict__ g_x, int64_t ldgx) {
  const int64_t total = N * F;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t n = t / F;
    const int f = static_cast<int>(t - n * F);
    const int64_t g = __ldg(batch + n);
    g_x[n * ldgx + f] = (g >= 0 && g < G) ? __ldg(g_pool + g * F + f) : 0.f;
  }
}

static int grid_for(int64_t work, int threads) {
  const int64_t want = ceil_div(work, threads), cap = int64_t(sm_count()) * 8;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace b200gat

using namespace b200gat;

static int check_readout_geom(const b200gat_readout_geom& g, const char* what) {
  B200GAT_REQUIRE(g.num_nodes >= 0 && g.num_graphs >= 0, B200GAT_E_SHAPE, "%s: negative sizes", what);
  B200GAT_REQUIRE(g.in_channels > 0 && g.hidden > 0 && g.classes > 0 && g.in_channels <= 4096 && g.hidden <= 4096 &&
                  g.classes <= 4096, B200GAT_E_UNSUPPORTED, "%s: in_channels / hidden / classes must be in [1, 4096]", what);
  return 0;
}

extern "C" int b200gat_readout_fwd(const b200gat_readout_fwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "readout_fwd: NULL args");
  int rc = check_readout_geom(a->geom, "readout_fwd");
  if (rc) return rc;
  const b200gat_readout_geom& g = a->geom;
  if (g.num_graphs == 0) return 0;
  B200GAT_REQUIRE(a->w1 && a->b1 && a->w2 && a->b2 && a->pooled && a->counts && a->hidden_out && a->logp && a->status,
                  B200GAT_E_NULL, "readout_fwd: NULL pointer");
  B200GAT_REQUIRE(g.num_nodes == 0 || (a->x && a->batch), B200GAT_E_NULL, "readout_fwd: NULL x / batch");
  B200GAT_REQUIRE(a->ldx >= g.in_channels, B200GAT_E_SHAPE, "readout_fwd: ldx < in_channels");
  B200GAT_REQUIRE(a->x_activation == ACT_NONE || a->x_activation == ACT_ELU, B200GAT_E_UNSUPPORTED,
                  "readout_fwd: unknown x_activation %d", a->x_activation);
  const int F = static_cast<int>(g.in_channels), Hd = static_cast<int>(g.hidden), K = static_cast<int>(g.classes);
  cudaError_t ce = cudaMemsetAsync(a->pooled, 0, size_t(g.num_graphs) * F * sizeof(float), stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(a->counts, 0, size_t(g.num_graphs) * sizeof(float), stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(a->status, 0, sizeof(int32_t), stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "readout_fwd: memset: %s", cudaGetErrorString(ce));
  if (g.num_nodes > 0) {
    readout_pool_kernel<<<grid_for(g.num_nodes * F, 256), 256, 0, stream>>>(a->x, a->ldx, a->batch, g.num_nodes, F, g.num_graphs,
                                                                              a->x_activation, a->pooled, a->counts, a->status);
    if ((rc = check_launch("readout_pool_kernel"))) return rc;
  }
  HeadParams p{g.num_graphs, F, Hd, K, a->w1, a->b1, a->w2, a->b2, a->pooled, a->counts, a->hidden_out, a->logp};
  const int64_t cap = int64_t(sm_count()) * 8;
  const int blocks = static_cast<int>(g.num_graphs < cap ? g.num_graphs : cap);
  readout_head_fwd_kernel<<<blocks, 128, size_t(F + Hd + K) * sizeof(float), stream>>>(p);
  return check_launch("readout_head_fwd_kernel");
}

extern "C" int b200gat_readout_bwd(const b200gat_readout_bwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(a, B200GAT_E_NULL, "readout_bwd: NULL args");
  int rc = check_readout_geom(a->geom, "readout_bwd");
  if (rc) return rc;
  const b200gat_readout_geom& g = a->geom;
  const int F = static_cast<int>(g.in_channels), Hd = static_cast<int>(g.hidden), K = static_cast<int>(g.classes);
  B200GAT_REQUIRE(a->g_w1 && a->g_b1 && a->g_w2 && a->g_b2, B200GAT_E_NULL, "readout_bwd: NULL parameter-gradient pointer");
  if (g.num_graphs == 0) {
    cudaError_t ce = cudaMemsetAsync(a->g_w1, 0, size_t(Hd) * F * sizeof(float), stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(a->g_b1, 0, size_t(Hd) * sizeof(float), stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(a->g_w2, 0, size_t(K) * Hd * sizeof(float), stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(a->g_b2, 0, size_t(K) * sizeof(float), stream);
    return ce == cudaSuccess ? 0 : fail(static_cast<int>(ce), "readout_bwd: memset: %s", cudaGetErrorString(ce));
  }
  B200GAT_REQUIRE(a->w1 && a->w2 && a->counts && a->pooled && a->hidden_out && a->logp && a->g_logp && a->workspace,
                  B200GAT_E_NULL, "readout_bwd: NULL pointer");
  const size_t need = size_t(g.num_graphs) * (K + Hd + F) * sizeof(float);
  B200GAT_REQUIRE(a->workspace_bytes >= need, B200GAT_E_WORKSPACE, "readout_bwd: workspace %zu < %zu bytes", a->workspace_bytes, need);
  B200GAT_REQUIRE(!a->g_x || (a->batch && a->ldgx >= F), B200GAT_E_SHAPE, "readout_bwd: g_x needs batch and ldgx >= in_channels");
  float* g_logits = static_cast<float*>(a->workspace);
  float* g_hid = g_logits + g.num_graphs * K;
  float* g_pool = g_hid + g.num_graphs * Hd;
  HeadBwdParams p{g.num_graphs, F, Hd, K, a->w1, a->w2, a->counts, a->hidden_out, a->logp, a->g_logp, g_logits, g_hid, g_pool};
  const int64_t cap = int64_t(sm_count()) * 8;
  const int blocks = static_cast<int>(g.num_graphs < cap ? g.num_graphs : cap);
  readout_head_bwd_kernel<<<blocks, 128, size_t(K + Hd) * sizeof(float), stream>>>(p);
  if ((rc = check_launch("readout_head_bwd_kernel"))) return rc;
  WgradParams wp{g.num_graphs, F, Hd, K, g_hid, a->pooled, g_logits, a->hidden_out, a->g_w1, a->g_b1, a->g_w2, a->g_b2};
  const int64_t outputs = int64_t(Hd) * F + Hd + int64_t(K) * Hd + K;
  readout_wgrad_kernel<<<grid_for(outputs * 8, 256), 256, 0, stream>>>(wp);
  if ((rc = check_launch("readout_wgrad_kernel"))) return rc;
  if (a->g_x && g.num_nodes > 0) {
    readout_scatter_bwd_kernel<<<grid_for(g.num_nodes * F, 256), 256, 0, stream>>>(g_pool, a->batch, g.num_nodes, F, g.num_graphs,
                                                                                    a->g_x, a->ldgx);
    rc = check_launch("readout_scatter_bwd_kernel");
  }
  return rc;
}
