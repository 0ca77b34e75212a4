// tcgen05 projection kernels (sm_100a): the only dense contraction of the path (GAT.py:43, per-head Linear) and its
// two backward GEMMs, on the 5th-gen tensor cores with TMA-fed shared-memory operands and TMEM accumulators.
//
// fp32 parity on tensor cores: the reference's Linear is a true-fp32 sgemm (torch allow_tf32=False).  One TF32 MMA
// lands ~1e-4 from it (fails the 1e-5 bar), so every operand is split  a = a_hi + a_lo  with a_hi = a & 0xffffe000
// (exactly TF32-representable) and a_lo = (a - a_hi) & 0xffffe000, and three kind::tf32 MMAs accumulate
//     a_lo*b_hi + a_hi*b_lo + a_hi*b_hi                                   ("3xTF32", error ~2^-21 relative)
// The split is a streaming pre-pass that also pads K to a multiple of 32 (TMA needs 16-byte row strides; F_in = 50,
// 1433, ... do not have them).
//
// Tensor-core fp32 accumulation TRUNCATES (measured on B200: the error of one long TMEM accumulation chain grows
// ~0.5 ulp per MMA, 9e-6 relative at K = 1024), so a TMEM accumulator only ever holds a SHORT chain: two 128 x BN
// accumulators alternate every TC_KC k-blocks (8 k-steps x 3 MMAs) and the epilogue warps promote each finished
// chunk into fp32 REGISTER accumulators with round-to-nearest adds while the tensor core fills the other buffer.
// The result is as accurate as an fp32 CUDA-core GEMM for any K (the split-K gW reduction runs K = 57k nodes).
//
// Kernel anatomy (one 128 x BN output tile per CTA):
//   warp 0    : TMA producer — cp.async.bulk.tensor.2d of the four operand tiles (A_hi, A_lo, B_hi, B_lo) of one
//               32-deep k-block into a 128B-swizzled stage, mbarrier complete_tx
//   warp 1    : TMEM allocator + MMA issuer — one elected lane issues 4 k-steps x 3 tcgen05.mma.kind::tf32
//               (M=128, N=BN, K=8) per stage; tcgen05.commit frees the stage / hands a finished chunk to the epilogue
//   warps 2.. : epilogue, 4 warps per 128 output columns — tcgen05.ld 32x32b.x32 of each finished chunk (TMEM lane
//               quarter = warp_id % 4) into 128 register accumulators; at the end + bias, fused attention-logit
//               reductions s_src = <Wh_h, a1_h> + b1_h, s_dst = <Wh_h, a2_h> + b2_h, fp32 store (or red.global.add
//               for the split-K gW reduction)
// All operands are K-major ([rows, K] row-major, K contiguous): X·W^T directly, gT·W through a transposed copy of the
// small W, and gT^T·X through transposing split pre-passes of gT and X (K = nodes).
#include "proj_tc.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace b200gat {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;   // fp32 elements per k-block row = 128 bytes = one SWIZZLE_128B atom row
constexpr int TC_KC = 2;    // k-blocks per TMEM accumulation chunk (2 x 4 k-steps x 3 MMAs = 24 MMAs per chain)

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major shared-memory matrix descriptor: 128-byte rows, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;            // LBO (unused for swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;    // SBO
  d |= uint64_t(1) << 46;            // descriptor version (sm_100)
  d |= uint64_t(2) << 61;            // LayoutType::SWIZZLE_128B
  return d;
}

// instruction descriptor: kind::tf32 (a/b format 2), fp32 accumulate (c format 1), K-major A and B, M=128, N=BN
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(TC_BM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ the GEMM
struct TcGemmParams {
  int64_t M, N;            // output extent
  int k_blocks;            // total k-blocks (K padded / 32)
  int k_blocks_per_split;  // k-blocks per blockIdx.z
  float* C; int64_t ldc;
  const float* bias;       // [N] or null
  // fused attention-logit epilogue (EPI_LOGITS)
  const float* a1; const float* a2; const float* b1; const float* b2; float* s_src; float* s_dst; int H, Cp;
};

enum { EPI_STORE = 0, EPI_LOGITS = 1, EPI_ATOMIC = 2 };

template <int BN>
struct TcCfg {
  static constexpr int STAGES = BN == 256 ? 2 : 3;
  static constexpr int EPI_WARPS = 4 * (BN / 128);
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;            // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int VEC_BYTES = 3 * BN * 4;                 // bias / a1 / a2 slices of the tile
  static constexpr int RED_BYTES = 2 * TC_BM * 4;              // cross-half logit partials (BN = 256, Cp = 256)
  static constexpr int BAR_BYTES = 128;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + VEC_BYTES + RED_BYTES + BAR_BYTES + 1024;   // + alignment slack
};

template <int BN, int EPI>
__global__ void __launch_bounds__(TcCfg<BN>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
               const TcGemmParams p) {
  using S = TcCfg<BN>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* vec = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
  float* red = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::VEC_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::VEC_BYTES + S::RED_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] chunk finished in TMEM buffer b
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] buffer b drained by every epilogue warp
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // n-tiles vary fastest: the CTAs resident together share one A (row) tile and sweep the small B operand, so A is
  // streamed from HBM once instead of once per n-tile (ncu, r1b: 1.9 GB -> the 4 n-tiles of Dp = 1024 re-read it)
  const int64_t n0 = int64_t(blockIdx.x) * BN;
  const int64_t m0 = int64_t(blockIdx.y) * TC_BM;
  const int kb0 = blockIdx.z * p.k_blocks_per_split;
  const int kb1 = (kb0 + p.k_blocks_per_split < p.k_blocks) ? kb0 + p.k_blocks_per_split : p.k_blocks;
  const int nkb = kb1 - kb0;
  const int nchunks = (nkb + TC_KC - 1) / TC_KC;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ah)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_al)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bl)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], S::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_ptr);
  if (warp >= 2) {   // epilogue warps stage the per-column vectors of this tile
    for (int c = threadIdx.x - 64; c < BN; c += 32 * S::EPI_WARPS) {
      const int64_t col = n0 + c;
      const bool ok = col < p.N;
      vec[c] = (ok && p.bias) ? __ldg(p.bias + col) : 0.f;
      if (EPI == EPI_LOGITS) {
        vec[BN + c] = ok ? __ldg(p.a1 + col) : 0.f;
        vec[2 * BN + c] = ok ? __ldg(p.a2 + col) : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = stage_base + s * S::STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        const int k0 = (kb0 + i) * TC_BK;
        tma_load_2d(st, &map_ah, &full_bar[s], k0, static_cast<int>(m0));
        tma_load_2d(st + S::A_BYTES, &map_al, &full_bar[s], k0, static_cast<int>(m0));
        tma_load_2d(st + 2 * S::A_BYTES, &map_bh, &full_bar[s], k0, static_cast<int>(n0));
        tma_load_2d(st + 2 * S::A_BYTES + S::B_BYTES, &map_bl, &full_bar[s], k0, static_cast<int>(n0));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        const int chunk = i / TC_KC, b = chunk & 1;
        const bool chunk_first = (i % TC_KC) == 0;
        const bool chunk_last = (i % TC_KC) == TC_KC - 1 || i == nkb - 1;
        if (chunk_first) {
          mbar_wait(&tempty_bar[b], ((chunk >> 1) & 1) ^ 1);   // the epilogue has drained this buffer
          tc_fence_after();
        }
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(b * BN);
        const uint32_t st = smem_u32(stage_base + s * S::STAGE_BYTES);
        const uint32_t a_hi = st, a_lo = st + S::A_BYTES, b_hi = st + 2 * S::A_BYTES, b_lo = b_hi + S::B_BYTES;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          // 8 fp32 = 32 B along the swizzled 128-B row
          const uint64_t dah = make_sdesc(a_hi + ks * 32), dal = make_sdesc(a_lo + ks * 32);
          const uint64_t dbh = make_sdesc(b_hi + ks * 32), dbl = make_sdesc(b_lo + ks * 32);
          mma_tf32(d_tmem, dal, dbh, idesc, (!chunk_first || ks > 0) ? 1u : 0u);
          mma_tf32(d_tmem, dah, dbl, idesc, 1u);
          mma_tf32(d_tmem, dah, dbh, idesc, 1u);
        }
        mma_commit(&empty_bar[s]);                  // frees the stage once these MMAs have read it
        if (chunk_last) mma_commit(&tfull_bar[b]);  // chunk complete in TMEM buffer b
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = ew >> 2;                       // which 128-column half of the tile
    const int64_t row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const int cbase = half * 128;                   // first tile column of this thread
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int b = c & 1;
      mbar_wait(&tfull_bar[b], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * BN + cbase);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[ch * 32 + j] += __uint_as_float(v[j]);   // round-to-nearest promotion
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[b]);
    }

    const int64_t col0 = n0 + cbase;                // first global column of this thread
    float* dst = p.C + row * p.ldc + col0;
    if (EPI == EPI_ATOMIC) {
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (col0 + j < p.N) atomicAdd(dst + j, acc[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] += vec[cbase + j];
      if (EPI == EPI_LOGITS) {
        float d1 = 0.f, d2 = 0.f;
        if (p.Cp <= 128) {
          // heads tile this thread's 128 columns (128 % Cp == 0, checked on the host)
#pragma unroll
          for (int j = 0; j < 128; ++j) {
            d1 = fmaf(acc[j], vec[BN + cbase + j], d1);
            d2 = fmaf(acc[j], vec[2 * BN + cbase + j], d2);
            if ((j + 1) % p.Cp == 0) {
              const int64_t col = col0 + j;
              if (row_ok && col < p.N) {
                const int h = static_cast<int>(col / p.Cp);
                p.s_src[row * p.H + h] = d1 + __ldg(p.b1 + h);
                p.s_dst[row * p.H + h] = d2 + __ldg(p.b2 + h);
              }
              d1 = 0.f;
              d2 = 0.f;
            }
          }
        } else {
          // Cp == 256 == BN: one head per tile, its two halves live in two warps -> combine through shared memory
#pragma unroll
          for (int j = 0; j < 128; ++j) {
            d1 = fmaf(acc[j], vec[BN + cbase + j], d1);
            d2 = fmaf(acc[j], vec[2 * BN + cbase + j], d2);
          }
          const int r = q * 32 + lane;
          if (half == 1) {
            red[r] = d1;
            red[TC_BM + r] = d2;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * S::EPI_WARPS) : "memory");
          if (half == 0 && row_ok && n0 < p.N) {
            const int h = static_cast<int>(n0 / p.Cp);
            p.s_src[row * p.H + h] = d1 + red[r] + __ldg(p.b1 + h);
            p.s_dst[row * p.H + h] = d2 + red[TC_BM + r] + __ldg(p.b2 + h);
          }
        }
      }
      if (row_ok) {
        const bool vec_ok = (col0 + 128 <= p.N) && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
        if (vec_ok) {
#pragma unroll
          for (int j = 0; j < 128; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 128; ++j)
            if (col0 + j < p.N) dst[j] = acc[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ split pre-passes
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
}

// src [rows, cols] (row stride ld) -> hi / lo [rows, ldp] with zero padding of columns [cols, ldp)
__global__ void __launch_bounds__(256)
split_pad_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols, float* __restrict__ hi,
                 float* __restrict__ lo, int64_t ldp) {
  const int64_t total = rows * ldp;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = t / ldp, c = t - r * ldp;
    float h = 0.f, l = 0.f;
    if (c < cols) split_tf32(__ldg(src + r * ld + c), h, l);
    hi[t] = h;
    lo[t] = l;
  }
}

// src [rows, cols] -> TRANSPOSED hi / lo [cols, ldp] (ldp = pad32(rows)), zero padded
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols, float* __restrict__ hi,
                       float* __restrict__ lo, int64_t ldp) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = int64_t(blockIdx.x) * 32, c0 = int64_t(blockIdx.y) * 32;   // x: the long (node) dimension
  for (int u = ty; u < 32; u += 8) {
    const int64_t r = r0 + u, c = c0 + tx;
    tile[u][tx] = (r < rows && c < cols) ? __ldg(src + r * ld + c) : 0.f;
  }
  __syncthreads();
  for (int u = ty; u < 32; u += 8) {
    const int64_t c = c0 + u, r = r0 + tx;       // output row = source column
    if (c < cols && r < ldp) {
      float h, l;
      split_tf32(tile[tx][u], h, l);
      hi[c * ldp + r] = h;
      lo[c * ldp + r] = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch `ld` elements; box = {32 cols (128 B, inner), box_rows}, SWIZZLE_128B
static int make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200GAT_E_UNSUPPORTED, "proj_tc: cuTensorMapEncodeTiled is not available");
  cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t gstr[1] = {cuuint64_t(ld) * 4};
  cuuint32_t box[2] = {cuuint32_t(TC_BK), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200GAT_E_SHAPE, "proj_tc: cuTensorMapEncodeTiled failed (%d)", int(r));
  return 0;
}

static inline int64_t pad32(int64_t v) { return (v + 31) / 32 * 32; }
static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

template <int BN, int EPI>
static int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                     const TcGemmParams& p, int splits, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BN, EPI>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::TOTAL);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "proj_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  dim3 grid(static_cast<unsigned>(ceil_div(p.N, BN)), static_cast<unsigned>(ceil_div(p.M, TC_BM)), static_cast<unsigned>(splits));
  kern<<<grid, TcCfg<BN>::THREADS, TcCfg<BN>::TOTAL, stream>>>(ah, al, bh, bl, p);
  return check_launch("gemm_tc_kernel");
}

static int launch_split(const float* src, int64_t ld, int64_t rows, int64_t cols, float* hi, float* lo, int64_t ldp,
                        cudaStream_t stream) {
  const int64_t total = rows * ldp;
  const int64_t want = ceil_div(total, 256);
  const int64_t cap = int64_t(sm_count()) * 16;
  split_pad_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, stream>>>(src, ld, rows, cols, hi, lo, ldp);
  return check_launch("split_pad_kernel");
}

static int launch_split_transpose(const float* src, int64_t ld, int64_t rows, int64_t cols, float* hi, float* lo,
                                  int64_t ldp, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>(ceil_div(ldp, 32)), static_cast<unsigned>(ceil_div(cols, 32)));
  split_transpose_kernel<<<grid, 256, 0, stream>>>(src, ld, rows, cols, hi, lo, ldp);
  return check_launch("split_transpose_kernel");
}

// C[M,N] (=, +bias | +=) A[M,K] · B[N,K]^T with pre-split K-major operands of pitch Kp; `splits` > 1 => EPI_ATOMIC
template <int EPI>
static int gemm_nt(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int64_t M, int64_t N,
                   int64_t Kp, TcGemmParams p, int splits, cudaStream_t stream) {
  p.M = M; p.N = N; p.k_blocks = static_cast<int>(Kp / TC_BK);
  p.k_blocks_per_split = static_cast<int>(ceil_div(p.k_blocks, splits));
  splits = static_cast<int>(ceil_div(p.k_blocks, p.k_blocks_per_split));
  CUtensorMap ah, al, bh, bl;
  int rc;
  const int bn = N > 128 ? 256 : 128;
  if ((rc = make_map(&ah, a_hi, M, Kp, Kp, TC_BM))) return rc;
  if ((rc = make_map(&al, a_lo, M, Kp, Kp, TC_BM))) return rc;
  if ((rc = make_map(&bh, b_hi, N, Kp, Kp, bn))) return rc;
  if ((rc = make_map(&bl, b_lo, N, Kp, Kp, bn))) return rc;
  if (bn == 256) return launch_tc<256, EPI>(ah, al, bh, bl, p, splits, stream);
  return launch_tc<128, EPI>(ah, al, bh, bl, p, splits, stream);
}

// ---- shape gates: the tensor-core path takes the projections that are worth a 128-row tile pipeline -----------------
static bool tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200GAT_DISABLE_TC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

bool proj_tc_fwd_supported(const b200gat_layer& L, int64_t N) {
  const int64_t Dp = L.heads * L.c_pad;
  return !tc_disabled() && N >= 512 && Dp >= 64 && L.in_channels >= 16 && N < (int64_t(1) << 31) - 256;
}
bool proj_tc_bwd_supported(const b200gat_layer& L, int64_t N) { return proj_tc_fwd_supported(L, N); }

struct FwdWs { size_t x_hi, x_lo, w_hi, w_lo, total; };
static FwdWs plan_fwd(const b200gat_layer& L, int64_t N) {
  const int64_t Kp = pad32(L.in_channels), Dp = L.heads * L.c_pad;
  FwdWs w;
  const size_t xs = up256(size_t(N) * Kp * 4), wsz = up256(size_t(Dp) * Kp * 4);
  w.x_hi = 0; w.x_lo = xs; w.w_hi = 2 * xs; w.w_lo = 2 * xs + wsz; w.total = 2 * xs + 2 * wsz;
  return w;
}
size_t proj_tc_fwd_workspace_bytes(const b200gat_layer& L, int64_t N) {
  return proj_tc_fwd_supported(L, N) ? plan_fwd(L, N).total : 0;
}

int proj_tc_fwd(const b200gat_proj_fwd_args& a, cudaStream_t stream) {
  const b200gat_layer& L = a.layer;
  const int64_t N = a.num_nodes, F = L.in_channels, H = L.heads, Cp = L.c_pad, Dp = H * Cp, Kp = pad32(F);
  const FwdWs w = plan_fwd(L, N);
  B200GAT_REQUIRE(a.workspace && a.workspace_bytes >= w.total, B200GAT_E_WORKSPACE, "proj_fwd: workspace %zu < %zu bytes",
                  a.workspace_bytes, w.total);
  B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(a.workspace) & 255u) == 0, B200GAT_E_ALIGN, "proj_fwd: workspace must be 256-byte aligned");
  char* base = static_cast<char*>(a.workspace);
  float* x_hi = reinterpret_cast<float*>(base + w.x_hi);
  float* x_lo = reinterpret_cast<float*>(base + w.x_lo);
  float* w_hi = reinterpret_cast<float*>(base + w.w_hi);
  float* w_lo = reinterpret_cast<float*>(base + w.w_lo);
  int rc;
  if ((rc = launch_split(a.x, a.ldx, N, F, x_hi, x_lo, Kp, stream))) return rc;
  if ((rc = launch_split(a.w, F, Dp, F, w_hi, w_lo, Kp, stream))) return rc;
  TcGemmParams p{};
  p.C = a.wh; p.ldc = Dp; p.bias = a.bw;
  p.a1 = a.a1; p.a2 = a.a2; p.b1 = a.b1; p.b2 = a.b2; p.s_src = a.s_src; p.s_dst = a.s_dst;
  p.H = static_cast<int>(H); p.Cp = static_cast<int>(Cp);
  const int bn = Dp > 128 ? 256 : 128;
  // heads must not straddle an epilogue thread's 128 columns, or be exactly one 256-wide tile
  const bool fuse_logits = (Cp <= 128 && 128 % Cp == 0) || (Cp == 256 && bn == 256);
  if (fuse_logits) return gemm_nt<EPI_LOGITS>(x_hi, x_lo, w_hi, w_lo, N, Dp, Kp, p, 1, stream);
  if ((rc = gemm_nt<EPI_STORE>(x_hi, x_lo, w_hi, w_lo, N, Dp, Kp, p, 1, stream))) return rc;
  return launch_logits(a, stream);
}

struct BwdWs { size_t g_hi, g_lo, wt_hi, wt_lo, gt_hi, gt_lo, xt_hi, xt_lo, total; };
static BwdWs plan_bwd_ws(const b200gat_layer& L, int64_t N) {
  const int64_t F = L.in_channels, Dp = L.heads * L.c_pad, Dk = pad32(Dp), Np = pad32(N);
  BwdWs w;
  const size_t gs = up256(size_t(N) * Dk * 4), ws = up256(size_t(F) * Dk * 4);
  const size_t gts = up256(size_t(Dp) * Np * 4), xts = up256(size_t(F) * Np * 4);
  size_t o = 0;
  w.g_hi = o; o += gs; w.g_lo = o; o += gs;
  w.wt_hi = o; o += ws; w.wt_lo = o; o += ws;
  w.gt_hi = o; o += gts; w.gt_lo = o; o += gts;
  w.xt_hi = o; o += xts; w.xt_lo = o; o += xts;
  w.total = o;
  return w;
}
size_t proj_tc_bwd_workspace_bytes(const b200gat_layer& L, int64_t N) {
  return proj_tc_bwd_supported(L, N) ? plan_bwd_ws(L, N).total : 0;
}

int proj_tc_bwd(const b200gat_proj_bwd_args& a, cudaStream_t stream) {
  const b200gat_layer& L = a.layer;
  const int64_t N = a.num_nodes, F = L.in_channels, Dp = L.heads * L.c_pad, Dk = pad32(Dp), Np = pad32(N);
  const BwdWs w = plan_bwd_ws(L, N);
  B200GAT_REQUIRE(a.workspace && a.workspace_bytes >= w.total, B200GAT_E_WORKSPACE, "proj_bwd: workspace %zu < %zu bytes",
                  a.workspace_bytes, w.total);
  B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(a.workspace) & 255u) == 0, B200GAT_E_ALIGN, "proj_bwd: workspace must be 256-byte aligned");
  char* base = static_cast<char*>(a.workspace);
  auto at = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  int rc;
  if (a.g_x) {
    // gX[N,F] = gT[N,Dp] · (W^T)[F,Dp]^T : K = Dp
    if ((rc = launch_split(a.g_t, Dp, N, Dp, at(w.g_hi), at(w.g_lo), Dk, stream))) return rc;
    if ((rc = launch_split_transpose(a.w, F, Dp, F, at(w.wt_hi), at(w.wt_lo), Dk, stream))) return rc;
    TcGemmParams p{};
    p.C = a.g_x; p.ldc = a.ldgx;
    if ((rc = gemm_nt<EPI_STORE>(at(w.g_hi), at(w.g_lo), at(w.wt_hi), at(w.wt_lo), N, F, Dk, p, 1, stream))) return rc;
  }
  // gW[Dp,F] = (gT^T)[Dp,N] · (X^T)[F,N]^T : K = nodes, split across CTAs, red.global.add epilogue
  if ((rc = launch_split_transpose(a.g_t, Dp, N, Dp, at(w.gt_hi), at(w.gt_lo), Np, stream))) return rc;
  if ((rc = launch_split_transpose(a.x, a.ldx, N, F, at(w.xt_hi), at(w.xt_lo), Np, stream))) return rc;
  cudaError_t ce = cudaMemsetAsync(a.g_w, 0, size_t(Dp) * F * sizeof(float), stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "proj_bwd: memset: %s", cudaGetErrorString(ce));
  const int bn = F > 128 ? 256 : 128;
  const int64_t tiles = ceil_div(Dp, TC_BM) * ceil_div(F, bn);
  int64_t splits = ceil_div(int64_t(sm_count()) * 2, tiles);
  const int64_t kblocks = Np / TC_BK;
  if (splits > kblocks / (2 * TC_KC)) splits = kblocks / (2 * TC_KC);
  if (splits < 1) splits = 1;
  TcGemmParams p{};
  p.C = a.g_w; p.ldc = F;
  return gemm_nt<EPI_ATOMIC>(at(w.gt_hi), at(w.gt_lo), at(w.xt_hi), at(w.xt_lo), Dp, F, Np, p, static_cast<int>(splits), stream);
}

}  // namespace b200gat
