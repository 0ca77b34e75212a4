// Plain fp32 CUDA-core GEMM used for the shapes the tcgen05 projection does not take (tiny K or tiny N: F_in = 3/5,
// D = 7, ...) and as the reduction engine for gW (split-K over the node dimension).
//   C[m, n] (+)= sum_k A(m, k) * B(k, n) (+ bias[n])
//   A(m,k) = A_KMAJOR ? A[m*lda + k] : A[k*lda + m]        B(k,n) = B_KMAJOR ? B[n*ldb + k] : B[k*ldb + n]
//   act_a / act_b = ACT_ELU applies the fused inter-layer activation to the operand as it is loaded (the operand is x)
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, true fp32 FMA (matches torch's allow_tf32=False default).
#pragma once
#include "split_blob.cuh"

namespace b200gat {

constexpr int GM = 64, GN = 64, GK = 16;

template <bool A_KMAJOR, bool B_KMAJOR, bool ATOMIC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                 float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int64_t M, int64_t N, int64_t K,
                 int64_t k_per_split, int act_a, int act_b) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = int64_t(blockIdx.x) * GM, n0 = int64_t(blockIdx.y) * GN;
  const int64_t kbeg = int64_t(blockIdx.z) * k_per_split;
  const int64_t kend = kbeg + k_per_split < K ? kbeg + k_per_split : K;
  const int tx = tid & 15, ty = tid >> 4;   // thread computes rows ty*4.., cols tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += GK) {
    // ---- stage A tile (GM x GK) ----
    if (A_KMAJOR) {
      const int m = tid >> 2, kk = (tid & 3) * 4;
      const int64_t gm = m0 + m;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t gk = k0 + kk + u;
        float v = (gm < M && gk < kend) ? A[gm * lda + gk] : 0.f;
        As[kk + u][m] = act_a ? elu_fwd(v) : v;
      }
    } else {
      const int kk = tid >> 4, m = (tid & 15) * 4;
      const int64_t gk = k0 + kk;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t gm = m0 + m + u;
        float v = (gm < M && gk < kend) ? A[gk * lda + gm] : 0.f;
        As[kk][m + u] = act_a ? elu_fwd(v) : v;
      }
    }
    // ---- stage B tile (GK x GN) ----
    if (B_KMAJOR) {
      const int n = tid >> 2, kk = (tid & 3) * 4;
      const int64_t gn = n0 + n;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t gk = k0 + kk + u;
        float v = (gn < N && gk < kend) ? B[gn * ldb + gk] : 0.f;
        Bs[kk + u][n] = act_b ? elu_fwd(v) : v;
      }
    } else {
      const int kk = tid >> 4, n = (tid & 15) * 4;
      const int64_t gk = k0 + kk;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t gn = n0 + n + u;
        float v = (gn < N && gk < kend) ? B[gk * ldb + gn] : 0.f;
        Bs[kk][n + u] = act_b ? elu_fwd(v) : v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (ATOMIC) {
        atomicAdd(&C[gm * ldc + gn], v);
      } else {
        if (bias) v += bias[gn];
        C[gm * ldc + gn] = v;
      }
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR>
inline int gemm_simt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                     const float* bias, int64_t M, int64_t N, int64_t K, int splits, cudaStream_t stream,
                     int act_a = ACT_NONE, int act_b = ACT_NONE) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid(static_cast<unsigned>(ceil_div(M, GM)), static_cast<unsigned>(ceil_div(N, GN)), 1);
  if (splits <= 1) {
    gemm_simt_kernel<A_KMAJOR, B_KMAJOR, false><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, bias, M, N, K, K, act_a, act_b);
  } else {
    int64_t kps = ceil_div(ceil_div(K, splits), GK) * GK;
    grid.z = static_cast<unsigned>(ceil_div(K, kps));
    cudaError_t e = cudaMemsetAsync(C, 0, size_t(M) * ldc * sizeof(float), stream);   // caller passes ldc == N here
    if (e != cudaSuccess) return fail(static_cast<int>(e), "gemm_simt: memset: %s", cudaGetErrorString(e));
    gemm_simt_kernel<A_KMAJOR, B_KMAJOR, true><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, nullptr, M, N, K, kps, act_a, act_b);
  }
  return check_launch("gemm_simt");
}

}  // namespace b200gat
