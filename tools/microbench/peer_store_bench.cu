// Micro-benchmark for the projection kernel's fused all-gather (b200gat_proj_fwd_args.wh_peers): how fast
// can ONE GPU write a [rows, 512] fp32 block into a PEER GPU's memory over NVLink, as a function of the store pattern?
//   (A) 16-byte stores in 64-byte runs per row   — what gemm_tc_kernel's store epilogue does today (measured through the
//       kernel: ~430 GB/s per GPU at 4 GPUs, where NCCL's all-gather sustains ~600 GB/s)
//   (B) 16-byte stores in 128-byte runs per row  — a 32-column staging chunk
//   (C) 16-byte stores, fully coalesced 512-byte runs (a plain copy kernel: the upper bound of the LD/ST path)
//   (D) cp.async.bulk shared -> global(peer) of 2 KB rows (TMA bulk store; one elected lane per row)
// Single process, two GPUs with peer access.  Run on a multi-GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_store_bench peer_store_bench.cu && ./peer_store_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int COLS = 512;                 // floats per row (Dp of the 4 x 128 layers)

// RUN = floats per contiguous run written by one group of RUN/4 lanes; a warp covers 32 / (RUN/4) rows per instruction
template <int RUN>
__global__ void __launch_bounds__(256) store_runs(float* __restrict__ dst, int64_t rows, float v) {
  constexpr int LPR = RUN / 4;            // lanes per run
  constexpr int RPW = 32 / LPR;           // rows per warp instruction
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int r_in = lane / LPR, c4 = (lane % LPR) * 4;
  const float4 val = make_float4(v, v + 1.f, v + 2.f, v + 3.f);
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    const int64_t r = r0 + r_in;
    if (r >= rows) continue;
#pragma unroll 4
    for (int c = 0; c < COLS; c += RUN) *reinterpret_cast<float4*>(dst + r * COLS + c + c4) = val;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one warp stages a 2 KB row in shared memory and one lane issues a bulk shared -> global store of it
__global__ void __launch_bounds__(256) store_bulk(float* __restrict__ dst, int64_t rows, float v) {
  __shared__ __align__(128) float stage[8][COLS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    for (int c = lane * 4; c < COLS; c += 128) *reinterpret_cast<float4*>(&stage[w][c]) = make_float4(v, v + 1.f, v + 2.f, v + 3.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(dst + r * COLS), "r"(smem_u32(&stage[w][0])), "r"(COLS * 4) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging row may be overwritten again
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float time_ms(F&& launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const int64_t rows = 1 << 20;           // 2 GiB block
  const size_t bytes = size_t(rows) * COLS * 4;
  float* local = nullptr; float* remote = nullptr;
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&local, bytes));
  bool have_peer = false;
  if (ndev >= 2) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (can) {
      CK(cudaSetDevice(1));
      CK(cudaMalloc(&remote, bytes));
      CK(cudaSetDevice(0));
      CK(cudaDeviceEnablePeerAccess(1, 0));
      have_peer = true;
    }
  }
  if (!have_peer) printf("no peer GPU: local-memory numbers only\n");
  const int grid = 148 * 8, reps = 5;
  struct Target { const char* name; float* ptr; } targets[2] = {{"local HBM", local}, {"peer over NVLink", remote}};
  for (const Target& t : targets) {
    if (!t.ptr) continue;
    float* p = t.ptr;
    const float a = time_ms([&] { store_runs<16><<<grid, 256>>>(p, rows, 1.f); }, reps);
    const float b = time_ms([&] { store_runs<32><<<grid, 256>>>(p, rows, 2.f); }, reps);
    const float c = time_ms([&] { store_runs<128><<<grid, 256>>>(p, rows, 3.f); }, reps);
    const float d = time_ms([&] { store_bulk<<<grid, 256>>>(p, rows, 4.f); }, reps);
    CK(cudaGetLastError());
    printf("%-18s  (A) 64 B runs %7.1f GB/s   (B) 128 B runs %7.1f GB/s   (C) 512 B runs %7.1f GB/s   (D) 2 KB bulk stores %7.1f GB/s\n",
           t.name, bytes / a / 1e6, bytes / b / 1e6, bytes / c / 1e6, bytes / d / 1e6);
  }
  return 0;
}
