// Micro-benchmark (go / no-go for the TMA-gather edge kernels): weighted neighbour aggregation
//     out[i, :] = sum_k w * T[col[k], :]      over a destination-sorted CSR
// (A) per-lane 128-bit gathers (ld.global.nc.v4), as edge_fwd.cu does, one warp per (row, 128-float4 slice)
// (B) cp.async.bulk (TMA, UBLKCP) of whole source rows into a per-warp shared-memory ring + LDS consumption
// on (1) a PPI-shaped block-diagonal batch (cached regime) and (2) a random graph whose table is far larger than L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- (A) LDG gathers: one warp per (row, slice of 128 floats... generalised: lane owns NV float4 of the row)
template <int NV, int U>
__global__ void __launch_bounds__(256) gather_ldg(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                  const float* __restrict__ T, float* __restrict__ out, int64_t N, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int slices = D / (128 * NV);
  for (int64_t item = warp; item < N * slices; item += nwarps) {
    const int64_t i = item / slices; const int s = item % slices;
    const int beg = rowptr[i], end = rowptr[i + 1];
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0, 0, 0, 0);
    const float w = 1.f / (end - beg);
    for (int k0 = beg; k0 < end; k0 += 32) {
      const int k = k0 + lane;
      const int j = k < end ? col[k] : (int)i;
      const int cnt = min(32, end - k0);
      for (int t = 0; t < cnt; t += U) {
        float4 g[U][NV]; float wt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jt = __shfl_sync(0xffffffffu, j, (t + u) & 31);
          wt[u] = (t + u < cnt) ? w : 0.f;
          const float* src = T + int64_t(jt) * D + s * 128 * NV;
#pragma unroll
          for (int v = 0; v < NV; ++v) g[u][v] = __ldg(reinterpret_cast<const float4*>(src) + lane + 32 * v);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(wt[u], g[u][v].x, acc[v].x); acc[v].y = fmaf(wt[u], g[u][v].y, acc[v].y);
            acc[v].z = fmaf(wt[u], g[u][v].z, acc[v].z); acc[v].w = fmaf(wt[u], g[u][v].w, acc[v].w);
          }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) reinterpret_cast<float4*>(out + i * D + s * 128 * NV)[lane + 32 * v] = acc[v];
  }
}

// ---- (A') the same loop with the real per-edge weight: p = exp(leaky(s_dst[i,h] + s_src[j,h]) - rowmax) (online),
// MODE 0: weights first, then gathers (edge_fwd.cu as of r1e)   MODE 1: first gather batch issued before the weight math
// MODE 2: two passes — pass 1 computes max/sum only (4-byte gathers), pass 2 recomputes p and gathers
template <int NV, int U, int MODE>
__global__ void __launch_bounds__(256) gather_softmax(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                      const float* __restrict__ T, const float* __restrict__ s_src,
                                                      const float* __restrict__ s_dst, float* __restrict__ out,
                                                      float* __restrict__ stats, int64_t N, int D) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int H = D / (128 * NV);
  for (int64_t item = warp; item < N * H; item += nwarps) {
    const int64_t i = item / H; const int h = item % H;
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
    const float sd = __ldg(s_dst + item);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0, 0, 0, 0);
    float m = -INFINITY, l = 0.f;
    const float* Th = T + h * 128 * NV;
    for (int k0 = beg; k0 < end; k0 += 32) {
      const int k = k0 + lane;
      const bool ok = k < end;
      const int j = ok ? __ldg(col + k) : (int)i;
      const int cnt = min(32, end - k0);
      float4 g[U][NV];
      auto load_batch = [&](int t) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jt = __shfl_sync(FULL, j, (t + u) & 31);
          const float* src = Th + int64_t(jt) * D;
#pragma unroll
          for (int v = 0; v < NV; ++v) g[u][v] = __ldg(reinterpret_cast<const float4*>(src) + lane + 32 * v);
        }
      };
      if (MODE == 1) load_batch(0);
      float e = -INFINITY;
      if (ok) { const float z = sd + __ldg(s_src + int64_t(j) * H + h); e = z > 0.f ? z : 0.2f * z; }
      float mx = e;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
      const float m_new = fmaxf(m, mx);
      if (k0 > beg && m_new != m) {
        const float sc = expf(m - m_new); l *= sc;
#pragma unroll
        for (int v = 0; v < NV; ++v) { acc[v].x *= sc; acc[v].y *= sc; acc[v].z *= sc; acc[v].w *= sc; }
      }
      m = m_new;
      const float pp = ok ? expf(e - m) : 0.f;
      l += pp;
      for (int t = 0; t < cnt; t += U) {
        if (MODE != 1 || t > 0) load_batch(t);
        float wt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { wt[u] = __shfl_sync(FULL, pp, (t + u) & 31); if (t + u >= cnt) wt[u] = 0.f; }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            acc[v].x = fmaf(wt[u], g[u][v].x, acc[v].x); acc[v].y = fmaf(wt[u], g[u][v].y, acc[v].y);
            acc[v].z = fmaf(wt[u], g[u][v].z, acc[v].z); acc[v].w = fmaf(wt[u], g[u][v].w, acc[v].w);
          }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(FULL, l, o);
    const float inv = 1.f / (l + 1e-16f);
    if (lane == 0) { stats[2 * item] = m; stats[2 * item + 1] = l; }
#pragma unroll
    for (int v = 0; v < NV; ++v)
      reinterpret_cast<float4*>(out + i * D + h * 128 * NV)[lane + 32 * v] = make_float4(acc[v].x * inv, acc[v].y * inv, acc[v].z * inv, acc[v].w * inv);
  }
}

// ---- (B) TMA bulk copies of whole rows into a per-warp ring of S stages
template <int S, int DV>   // DV = D / 128 float4 slots per lane
__global__ void __launch_bounds__(256) gather_tma(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                  const float* __restrict__ T, float* __restrict__ out, int64_t N, int D) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t row_bytes = D * 4;
  uint8_t* ring = smem + size_t(wid) * S * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(nw) * S * row_bytes) + wid * S;
  if (lane == 0)
    for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t warp = int64_t(blockIdx.x) * nw + wid, nwarps = int64_t(gridDim.x) * nw;
  uint32_t issued = 0, consumed = 0;     // running counters over the whole kernel -> stage = n % S, parity = (n / S) & 1
  for (int64_t i = warp; i < N; i += nwarps) {
    const int beg = rowptr[i], end = rowptr[i + 1];
    const float w = 1.f / (end - beg);
    float4 acc[DV];
#pragma unroll
    for (int v = 0; v < DV; ++v) acc[v] = make_float4(0, 0, 0, 0);
    int kissue = beg;
    // prologue: fill the ring
    if (lane == 0) {
      for (; kissue < end && kissue - beg < S; ++kissue, ++issued) {
        const uint32_t st = issued % S;
        const uint32_t bar = smem_u32(&bars[st]), dst = smem_u32(ring + st * row_bytes);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(T + int64_t(col[kissue]) * D), "r"(row_bytes), "r"(bar) : "memory");
      }
    }
    for (int k = beg; k < end; ++k, ++consumed) {
      const uint32_t st = consumed % S, ph = (consumed / S) & 1;
      const uint32_t bar = smem_u32(&bars[st]);
      asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(ph) : "memory");
      const float4* src = reinterpret_cast<const float4*>(ring + st * row_bytes);
#pragma unroll
      for (int v = 0; v < DV; ++v) {
        const float4 g = src[lane + 32 * v];
        acc[v].x = fmaf(w, g.x, acc[v].x); acc[v].y = fmaf(w, g.y, acc[v].y);
        acc[v].z = fmaf(w, g.z, acc[v].z); acc[v].w = fmaf(w, g.w, acc[v].w);
      }
      __syncwarp();
      if (lane == 0) {
        const int kn = k + S;
        if (kn < end) {
          const uint32_t dst = smem_u32(ring + st * row_bytes);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(T + int64_t(col[kn]) * D), "r"(row_bytes), "r"(bar) : "memory");
          ++issued;
        }
      }
      if (lane != 0 && k + S < end) ++issued;   // keep the counter warp-uniform
    }
    // lanes != 0 never ran the prologue: re-sync the issue counter
    issued = __shfl_sync(0xffffffffu, issued, 0);
#pragma unroll
    for (int v = 0; v < DV; ++v) reinterpret_cast<float4*>(out + i * D)[lane + 32 * v] = acc[v];
  }
}

struct Graph { std::vector<int> rowptr, col; int64_t N; };
static Graph make_graph(int64_t N, int blocks, int deg, unsigned seed) {
  Graph g; g.N = N; g.rowptr.resize(N + 1); g.col.reserve(N * (deg + 1));
  srand(seed);
  const int64_t bs = (N + blocks - 1) / blocks;
  uint64_t x = 88172645463325252ull + seed;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (int64_t i = 0; i < N; ++i) {
    g.rowptr[i] = (int)g.col.size();
    const int64_t b0 = (i / bs) * bs, bn = std::min(bs, N - b0);
    const int d = 1 + (int)(rnd() % (2 * deg - 1));
    for (int k = 0; k < d; ++k) g.col.push_back((int)(b0 + rnd() % bn));
    g.col.push_back((int)i);
  }
  g.rowptr[N] = (int)g.col.size();
  return g;
}

// both endpoints = perm[floor(N * U^2)] (the BASELINE power-law graph: in-degree max ~40k, median 19), dst-sorted CSR
static Graph make_powerlaw(int64_t N, int64_t E, unsigned seed) {
  Graph g; g.N = N;
  uint64_t x = 88172645463325252ull + seed;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  std::vector<int> perm(N);
  for (int64_t i = 0; i < N; ++i) perm[i] = (int)i;
  for (int64_t i = N - 1; i > 0; --i) std::swap(perm[i], perm[rnd() % (i + 1)]);
  auto pick = [&]() { double u = (rnd() >> 11) * (1.0 / 9007199254740992.0); int64_t k = (int64_t)(u * u * N); return perm[k < N ? k : N - 1]; };
  std::vector<int> src(E), dst(E);
  std::vector<int> cnt(N + 1, 0);
  for (int64_t e = 0; e < E; ++e) { src[e] = pick(); dst[e] = pick(); cnt[dst[e] + 1]++; }
  for (int64_t i = 0; i < N; ++i) cnt[i + 1] += 1;          // self loops
  g.rowptr.assign(N + 1, 0);
  for (int64_t i = 0; i < N; ++i) g.rowptr[i + 1] = g.rowptr[i] + cnt[i + 1];
  g.col.resize(g.rowptr[N]);
  std::vector<int> fill(g.rowptr.begin(), g.rowptr.end() - 1);
  for (int64_t e = 0; e < E; ++e) g.col[fill[dst[e]]++] = src[e];
  for (int64_t i = 0; i < N; ++i) g.col[fill[i]++] = (int)i;
  int mx = 0; for (int64_t i = 0; i < N; ++i) mx = std::max(mx, g.rowptr[i + 1] - g.rowptr[i]);
  printf("power-law graph: max in-degree %d\n", mx);
  return g;
}

template <typename F> static float time_ms(F f, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int r = 0; r < reps; ++r) f();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms / reps;
}

template <int DV> static void run_case(const char* name, int64_t N, int blocks, int deg, int reps, int64_t powerlaw_edges = 0) {
  const int D = DV * 128;
  Graph g = powerlaw_edges ? make_powerlaw(N, powerlaw_edges, 1) : make_graph(N, blocks, deg, 1);
  const int64_t E = g.col.size();
  int *rowptr, *col; float *T, *out;
  CK(cudaMalloc(&rowptr, (N + 1) * 4)); CK(cudaMalloc(&col, E * 4));
  CK(cudaMalloc(&T, size_t(N) * D * 4)); CK(cudaMalloc(&out, size_t(N) * D * 4));
  CK(cudaMemcpy(rowptr, g.rowptr.data(), (N + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(col, g.col.data(), E * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(T, 0, size_t(N) * D * 4));
  float *s_src, *s_dst, *stats;
  CK(cudaMalloc(&s_src, size_t(N) * 8 * 4)); CK(cudaMalloc(&s_dst, size_t(N) * 8 * 4)); CK(cudaMalloc(&stats, size_t(N) * 8 * 8));
  CK(cudaMemset(s_src, 0, size_t(N) * 8 * 4)); CK(cudaMemset(s_dst, 0, size_t(N) * 8 * 4));
  const double gather_gb = double(E) * D * 4 / 1e9;
  printf("== %s: N=%lld E'=%lld D=%d  gathered %.2f GB, table %.2f GB\n", name, (long long)N, (long long)E, D, gather_gb, double(N) * D * 4 / 1e9);
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  {
    float ms = time_ms([&] { gather_ldg<1, 8><<<sms * 8, 256>>>(rowptr, col, T, out, N, D); }, reps);
    printf("  LDG  NV=1 U=8 (warp per 128-float4 slice)   %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
    if (DV % 2 == 0) {
      ms = time_ms([&] { gather_ldg<2, 4><<<sms * 8, 256>>>(rowptr, col, T, out, N, D); }, reps);
      printf("  LDG  NV=2 U=4                                %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
    }
  }
  if (DV % 2 == 0) {
    float ms = time_ms([&] { gather_softmax<2, 4, 0><<<sms * 8, 256>>>(rowptr, col, T, s_src, s_dst, out, stats, N, D); }, reps);
    printf("  LDG+softmax NV=2 U=4 weights first           %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
    ms = time_ms([&] { gather_softmax<2, 4, 1><<<sms * 8, 256>>>(rowptr, col, T, s_src, s_dst, out, stats, N, D); }, reps);
    printf("  LDG+softmax NV=2 U=4 first batch early       %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
    ms = time_ms([&] { gather_softmax<2, 2, 1><<<sms * 8, 256>>>(rowptr, col, T, s_src, s_dst, out, stats, N, D); }, reps);
    printf("  LDG+softmax NV=2 U=2 first batch early       %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
  }
  {
    float ms = time_ms([&] { gather_softmax<1, 8, 0><<<sms * 8, 256>>>(rowptr, col, T, s_src, s_dst, out, stats, N, D); }, reps);
    printf("  LDG+softmax NV=1 U=8 weights first           %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
    ms = time_ms([&] { gather_softmax<1, 8, 1><<<sms * 8, 256>>>(rowptr, col, T, s_src, s_dst, out, stats, N, D); }, reps);
    printf("  LDG+softmax NV=1 U=8 first batch early       %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
    ms = time_ms([&] { gather_softmax<1, 4, 1><<<sms * 8, 256>>>(rowptr, col, T, s_src, s_dst, out, stats, N, D); }, reps);
    printf("  LDG+softmax NV=1 U=4 first batch early       %8.3f ms  %7.1f GB/s gathered\n", ms, gather_gb / ms * 1e3);
  }
  auto run_tma = [&](auto kern, int S, int warps) {
    const size_t sm = size_t(warps) * S * D * 4 + size_t(warps) * S * 8;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, sm));
    float ms = time_ms([&] { kern<<<sms * occ, warps * 32, sm>>>(rowptr, col, T, out, N, D); }, reps);
    printf("  TMA  S=%d warps/CTA=%d CTAs/SM=%d (%3zu KB smem/CTA) %8.3f ms  %7.1f GB/s gathered\n", S, warps, occ, sm >> 10, ms, gather_gb / ms * 1e3);
  };
  run_tma(gather_tma<2, DV>, 2, 8);
  run_tma(gather_tma<4, DV>, 4, 8);
  run_tma(gather_tma<4, DV>, 4, 4);
  run_tma(gather_tma<8, DV>, 8, 4);
  if (DV <= 4) run_tma(gather_tma<8, DV>, 8, 8);
  CK(cudaGetLastError());
  CK(cudaFree(rowptr)); CK(cudaFree(col)); CK(cudaFree(T)); CK(cudaFree(out));
}

int main() {
  run_case<8>("PPI-shaped (24 blocks, cached regime), D=1024", 56944, 24, 14, 20);
  run_case<4>("large random (streaming regime), D=512", 1200000, 1, 26, 3);
  run_case<4>("power-law 2.4M nodes / 62M edges (streaming regime, skewed degrees), D=512", 2400000, 1, 0, 2, 62000000);
  return 0;
}
