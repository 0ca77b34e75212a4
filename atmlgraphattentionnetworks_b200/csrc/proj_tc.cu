// tcgen05 projection kernels (sm_100a): the only dense contraction of the path (GAT.py:43, per-head Linear) and its
// two backward GEMMs, on the 5th-gen tensor cores with TMA-fed shared-memory operands and TMEM accumulators.
//
// fp32 parity on tensor cores ("3xFP16 split"): the reference's Linear is a true-fp32 sgemm (torch allow_tf32=False);
// one TF32 or one 16-bit MMA lands 1e-4..1e-3 from it (fails the 1e-5 bar).  Every operand tensor T is therefore
// stored as TWO fp16 planes of  T * s  (s = a power of two chosen from max|T| so that max|T*s| is in [2^14, 2^15)):
//     hi = fp16(T*s)            (11 significant bits)
//     lo = fp16(T*s - hi)       (the next 11 bits; exact residual, unscaled, so one accumulator serves all products)
// and three kind::f16 MMAs accumulate  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  in fp32 — every product is exact, the
// dropped a_lo*b_lo term is 2^-22 relative.  Per element the representation error is max(2^-22 |t|, 2^-40 max|T|):
// elements within 2^-18 of the tensor's largest magnitude carry 22 bits, smaller ones an absolute 2^-40 max|T|.
// kind::f16 runs at twice the kind::tf32 rate and the planes are half the bytes of a TF32 hi/lo pair, so this is 2x the
// 3xTF32 ceiling in both the tensor pipe and shared-memory / L2 operand traffic.  The result is multiplied by
// 1/(s_a s_b) (exact) in the epilogue.
//
// No transposes: tcgen05 takes either major-ness from shared memory (instruction-descriptor bits 15/16), so the same
// row-major planes feed   X·W^T (A, B K-major),   gT·W (A K-major, B = W planes MN-major)   and   gT^T·X (A = gT planes
// MN-major, B = X planes MN-major, K = nodes, split-K) — X and W are split once in the forward and reused.
//
// Tensor-core fp32 accumulation TRUNCATES (measured on B200: the error of one long TMEM accumulation chain grows
// ~0.5 ulp per MMA), so a TMEM accumulator only ever holds a SHORT chain: two 128 x BN accumulators alternate every
// TC_KC k-blocks (16 k-steps x 3 MMAs) and the epilogue warps promote each finished chunk into fp32 REGISTER
// accumulators with round-to-nearest adds while the tensor core fills the other buffer.
//
// Kernel anatomy (persistent: one CTA per SM walks a static list of (split, m-tile, n-tile) work items, n fastest):
//   warp 0    : TMA producer — cp.async.bulk.tensor.2d of the four operand planes (A_hi, A_lo, B_hi, B_lo) of one
//               64-deep k-block into a 128B-swizzled stage, mbarrier complete_tx
//   warp 1    : TMEM allocator + MMA issuer — one elected lane issues 4 k-steps x 3 tcgen05.mma.kind::f16
//               (M=128, N=BN, K=16) per stage; tcgen05.commit frees the stage / hands a finished chunk to the epilogue
//   warps 2.. : epilogue, 4 warps per 128 output columns — tcgen05.ld 32x32b.x16 of each finished chunk (TMEM lane
//               quarter = warp_id % 4) into 128 register accumulators; at the end of a work item: * 1/(s_a s_b), + bias,
//               fused attention-logit reductions s_src = <Wh_h, a1_h> + b1_h, s_dst = <Wh_h, a2_h> + b2_h, fp32 store
//               (or red.global.add for the split-K gW reduction) — overlapping the next item's first chunks
#include "proj_tc.cuh"
#include "split_blob.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace b200gat {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;   // fp16 elements per k-block row = 128 bytes = one SWIZZLE_128B atom row
// k-blocks per TMEM accumulation chunk (4 x 4 k-steps x 3 MMAs = 48 MMAs per chain).  Measured (tools/gemm_error.py, normalised
// max error of out / gX / gW against f64, independent of K beyond one chunk): 2 -> 6..8e-7, 4 -> 1.0..1.3e-6, 8 -> 2.0..2.7e-6,
// 16 -> 3.5..5.2e-6 — the truncation bias grows linearly with the chain.  4 keeps the GEMMs at the error of a plain fp32
// GEMM (8x inside the 1e-5 bar) and halves the chunk hand-offs (MMA -> epilogue drain -> MMA), which at 2 left the tensor
// pipe idle 39 % of the time: PPI-shaped step 4.94 -> 4.75 ms.  B200GAT_TC_KC=<n> overrides (2 = the tighter setting).
constexpr int TC_KC = 4;
constexpr int TC_MN_CHUNK_BYTES = 64 * TC_BK * 2;   // one MN-major TMA box: 64 k-rows x 64 fp16 (128 B) = 8 KB

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS, int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int COLS, int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// ---- CTA-pair (cta_group::2) helpers.  Shared-memory addresses in the shared::cluster window carry the CTA rank in bit 24;
// clearing it addresses the same offset in the pair's leader (even) CTA.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a pair member: the bytes land in the executing CTA, the transaction completes on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the pair leader's copy of `bar`
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(uint16_t(3)) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_128B, descriptor version 1 (sm_100).
//   K-major  operand (rows of 64 fp16 = 128 B along K): 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major operand (rows of 64 fp16 = 128 B along M/N, one row per k): 8-k groups 1024 B apart (SBO), the next
//            64 M/N elements TC_MN_CHUNK_BYTES further (LBO) — each chunk is one 64 x 64 TMA box.
template <bool MN>
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(MN ? (TC_MN_CHUNK_BYTES >> 4) : 1) << 16;   // LBO
  d |= uint64_t(1024 >> 4) << 32;                           // SBO
  d |= uint64_t(1) << 46;                                   // descriptor version (sm_100)
  d |= uint64_t(2) << 61;                                   // LayoutType::SWIZZLE_128B
  return d;
}
// bytes to advance an operand's start address per 16-deep k-step
template <bool MN>
__device__ __forceinline__ constexpr uint32_t kstep_bytes() { return MN ? 16 * 128 : 16 * 2; }

// instruction descriptor: kind::f16 with fp16 A/B (format 0), fp32 accumulate (c format 1), M=128, N=BN
template <int BN, bool A_MN, bool B_MN, int CG>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (uint32_t(A_MN) << 15) | (uint32_t(B_MN) << 16) | (uint32_t(BN >> 3) << 17) |
         (uint32_t((TC_BM * CG) >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ the GEMM
struct TcGemmParams {
  int64_t M, N;            // output extent
  int k_blocks;            // total k-blocks (ceil(K / 64))
  int k_blocks_per_split;  // k-blocks per split
  int m_tiles, n_tiles, splits;
  float* C; int64_t ldc;
  const float* bias;       // [N] or null
  const float* inv_a; const float* inv_b;   // device scalars: 1/s of the two operand splits
  // fused attention-logit epilogue (EPI_LOGITS)
  const float* a1; const float* a2; const float* b1; const float* b2; float* s_src; float* s_dst; int H, Cp;
  // projection fused with the all-gather of its output (b200gat_proj_fwd_args.wh_peers): every stored tile is also written
  // to the same (row, column) of n_peer peer-mapped copies of C
  float* peer_c[B200GAT_MAX_PEERS]; int n_peer;
  uint16_t* c16;           // optional bf16 copy of C (same ldc): the gathered-row storage of the bf16 mode (wh_bf16)
  // EPI_PREP: the gX GEMM of layer k+1 also runs the PREP pass of layer k's edge backward (its output gX IS layer k's upstream
  // gradient): C receives G = gX * act'(out_k) — the gatherable gradient rows — and the row records / g_bias are produced here
  const float* pr_out; int64_t pr_ldo;        // layer k's pre-activation output [M, N]
  const float* pr_bias;                       // [N]
  const float* pr_s_dst; const float* pr_rowmax; const float* pr_rowsum;   // [M, pr_H]
  float4* pr_rowrec;                          // out [M, pr_H] {s_dst, rowmax, 1/(rowsum + 1e-16), Drow}
  float* pr_g_bias;                           // out [N], zero-initialised by the host, accumulated atomically
  int pr_H, pr_C, pr_act;                     // layer k's heads / channels per head (N = pr_H * pr_C), 1: multiply by ELU'(out)
  int kc;                                     // k-blocks per TMEM accumulation chunk
};

enum { EPI_STORE = 0, EPI_LOGITS = 1, EPI_ATOMIC = 2, EPI_PREP = 3 };

// CG = CTAs per MMA (cta_group): 1 = one 128 x BN tile per CTA; 2 = a CTA PAIR (cluster of two SMs) computes one 256 x BN
// tile: each CTA stages its own 128 A rows and HALF of the B tile (BN/2 rows), one tcgen05.mma.cta_group::2 issued by the
// leader multiplies both, and each CTA's TMEM receives its own 128 rows of the result.  B is fetched once per pair (the
// 1-CTA kernel is bound by operand bytes entering the SM: 96 KB per 1536-cycle k-block) and a stage shrinks to 64 KB, so
// three stages fit.
template <int BN, int CG = 1>
struct TcCfg {
  static constexpr int STAGES = (BN == 256 && CG == 1) ? 2 : 3;
  static constexpr int EPI_WARPS = 4 * (BN / 128);
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;            // 16 KB per plane
  static constexpr int B_ROWS = BN / CG;                       // B rows staged by one CTA
  static constexpr int B_BYTES = B_ROWS * TC_BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int VEC_BYTES = 3 * BN * 4;                 // bias / a1 / a2 slices of the current tile
  static constexpr int RED_BYTES = 2 * TC_BM * 4;              // cross-half logit partials (BN = 256, Cp = 256)
  static constexpr int BAR_BYTES = 128;
  static constexpr int STG_ROW = 20;                           // floats per staged row: 16 columns + 4 pad (conflict-free 128-bit access)
  static constexpr int STG_BYTES = EPI_WARPS * 32 * STG_ROW * 4;   // per-warp transposing buffer of the store epilogue
  static constexpr int TOTAL = STAGES * STAGE_BYTES + VEC_BYTES + RED_BYTES + BAR_BYTES + STG_BYTES + 1024;   // + alignment slack
};

// one operand plane of one k-block: K-major = one box {64 k, ROWS}; MN-major = ROWS/64 boxes {64 mn, 64 k}
template <bool MN, int ROWS, int CG>
__device__ __forceinline__ void load_plane(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int k0, int mn0) {
  if (!MN) {
    if (CG == 1) tma_load_2d(dst, map, bar, k0, mn0);
    else tma_load_2d_pair(dst, map, bar, k0, mn0);
  } else {
#pragma unroll
    for (int c = 0; c < ROWS / 64; ++c) {
      if (CG == 1) tma_load_2d(dst + c * TC_MN_CHUNK_BYTES, map, bar, mn0 + 64 * c, k0);
      else tma_load_2d_pair(dst + c * TC_MN_CHUNK_BYTES, map, bar, mn0 + 64 * c, k0);
    }
  }
}

// s_src / s_dst of the 128 / CP heads inside one epilogue thread's 128 columns (acc already holds Wh + bw)
template <int CP>
__device__ __forceinline__ void logits_heads(const float (&acc)[128], const float* __restrict__ va1,
                                             const float* __restrict__ va2, const TcGemmParams& p, int64_t row, bool row_ok,
                                             int64_t col0) {
#pragma unroll
  for (int hh = 0; hh < 128 / CP; ++hh) {
    float d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      d1 = fmaf(acc[hh * CP + c], va1[hh * CP + c], d1);
      d2 = fmaf(acc[hh * CP + c], va2[hh * CP + c], d2);
    }
    const int64_t col = col0 + hh * CP;
    if (row_ok && col < p.N) {
      const int h = static_cast<int>(col / CP);
      p.s_src[row * p.H + h] = d1 + __ldg(p.b1 + h);
      p.s_dst[row * p.H + h] = d2 + __ldg(p.b2 + h);
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CG>
__global__ void __launch_bounds__(TcCfg<BN, CG>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
               const TcGemmParams p) {
  using S = TcCfg<BN, CG>;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;      // position in the CTA pair; 0 = leader (issues the MMAs)
  const int unit = CG == 2 ? int(blockIdx.x >> 1) : int(blockIdx.x);          // the pair / CTA walking the item list
  const int nunits = CG == 2 ? int(gridDim.x >> 1) : int(gridDim.x);
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the __shared__ array: an integer round trip would
  // lose the address space and turn every epilogue read of vec[] / red[] into a generic LD (ncu: 39 % of all stalls)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  float* vec = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
  float* red = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::VEC_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::VEC_BYTES + S::RED_BYTES);
  float* stgbuf = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::VEC_BYTES + S::RED_BYTES + S::BAR_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] chunk finished in TMEM buffer b
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] buffer b drained by every epilogue warp
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ah)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_al)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bl)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                   // pair: only the leader's copy is used (both CTAs' TMA bytes land on it)
      mbar_init(&empty_bar[s], 1);                  // one (multicast) tcgen05.commit per use
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], CG * S::EPI_WARPS); // pair: the leader's copy collects both CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<2 * BN, CG>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();                      // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // work item w -> (split, m-tile, n-tile); n-tiles vary fastest so that the CTAs resident together share A (row)
  // tiles and sweep the small B operand: A is streamed from HBM once instead of once per n-tile
  auto item_coords = [&](int w, int& m0, int& n0, int& kb0, int& nkb) {
    const int nt = w % p.n_tiles;
    const int r = w / p.n_tiles;
    const int mt = r % p.m_tiles, sp = r / p.m_tiles;   // m_tiles counts 128*CG-row tiles
    m0 = mt * (TC_BM * CG) + int(rank) * TC_BM;          // this CTA's 128 rows
    n0 = nt * BN;
    kb0 = sp * p.k_blocks_per_split;
    const int kb1 = (kb0 + p.k_blocks_per_split < p.k_blocks) ? kb0 + p.k_blocks_per_split : p.k_blocks;
    nkb = kb1 - kb0;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = unit; w < total_items; w += nunits) {
        int m0, n0, kb0, nkb;
        item_coords(w, m0, n0, kb0, nkb);
        const int nb0 = n0 + int(rank) * S::B_ROWS;     // pair: this CTA stages its half of the B tile
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t st = smem_u32(stage_base + s * S::STAGE_BYTES);
          if (rank == 0) mbar_expect_tx(&full_bar[s], CG * S::STAGE_BYTES);   // the bytes of BOTH CTAs
          const int k0 = (kb0 + i) * TC_BK;
          load_plane<A_MN, TC_BM, CG>(st, &map_ah, &full_bar[s], k0, m0);
          load_plane<A_MN, TC_BM, CG>(st + S::A_BYTES, &map_al, &full_bar[s], k0, m0);
          load_plane<B_MN, S::B_ROWS, CG>(st + 2 * S::A_BYTES, &map_bh, &full_bar[s], k0, nb0);
          load_plane<B_MN, S::B_ROWS, CG>(st + 2 * S::A_BYTES + S::B_BYTES, &map_bl, &full_bar[s], k0, nb0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc<BN, A_MN, B_MN, CG>();
      uint32_t it = 0, chunk = 0;
      for (int w = unit; w < total_items; w += nunits) {
        int m0, n0, kb0, nkb;
        item_coords(w, m0, n0, kb0, nkb);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          const int b = chunk & 1;
          const bool chunk_first = (i % p.kc) == 0;
          const bool chunk_last = (i % p.kc) == p.kc - 1 || i == nkb - 1;
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
          for (int ks = 0; ks < TC_BK / 16; ++ks) {
            const uint32_t ao = ks * kstep_bytes<A_MN>(), bo = ks * kstep_bytes<B_MN>();
            const uint64_t dah = make_sdesc<A_MN>(a_hi + ao), dal = make_sdesc<A_MN>(a_lo + ao);
            const uint64_t dbh = make_sdesc<B_MN>(b_hi + bo), dbl = make_sdesc<B_MN>(b_lo + bo);
            if (CG == 1) {
              mma_f16(d_tmem, dal, dbh, idesc, (!chunk_first || ks > 0) ? 1u : 0u);
              mma_f16(d_tmem, dah, dbl, idesc, 1u);
              mma_f16(d_tmem, dah, dbh, idesc, 1u);
            } else {
              mma_f16_pair(d_tmem, dal, dbh, idesc, (!chunk_first || ks > 0) ? 1u : 0u);
              mma_f16_pair(d_tmem, dah, dbl, idesc, 1u);
              mma_f16_pair(d_tmem, dah, dbh, idesc, 1u);
            }
          }
          // frees the stage (in both CTAs of a pair) once these MMAs have read it
          if (CG == 1) mma_commit(&empty_bar[s]); else mma_commit_pair(&empty_bar[s]);
          if (chunk_last) {                           // chunk complete in TMEM buffer b (of both CTAs)
            if (CG == 1) mma_commit(&tfull_bar[b]); else mma_commit_pair(&tfull_bar[b]);
            ++chunk;
          }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = ew >> 2;                       // which 128-column half of the tile
    const int cbase = half * 128;                   // first tile column of this thread
    const float inv = __ldg(p.inv_a) * __ldg(p.inv_b);
    uint32_t chunk = 0;
    for (int w = unit; w < total_items; w += nunits) {
      int m0, n0, kb0, nkb;
      item_coords(w, m0, n0, kb0, nkb);
      const int nchunks = (nkb + p.kc - 1) / p.kc;
      const int64_t row = int64_t(m0) + q * 32 + lane;
      const bool row_ok = row < p.M;
      if (EPI != EPI_ATOMIC) {
        // stage the per-column vectors of this tile (the previous item's readers are past the barrier)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * S::EPI_WARPS) : "memory");
        for (int c = threadIdx.x - 64; c < BN; c += 32 * S::EPI_WARPS) {
          const int64_t col = int64_t(n0) + c;
          const bool ok = col < p.N;
          vec[c] = (ok && p.bias) ? __ldg(p.bias + col) : 0.f;
          if (EPI == EPI_PREP) vec[BN + c] = ok ? __ldg(p.pr_bias + col) : 0.f;      // layer k's bias (O = out - bias)
          if (EPI == EPI_LOGITS) {
            vec[BN + c] = ok ? __ldg(p.a1 + col) : 0.f;
            vec[2 * BN + c] = ok ? __ldg(p.a2 + col) : 0.f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * S::EPI_WARPS) : "memory");
      }
      if (EPI == EPI_PREP) {
        // the epilogue has no slack for DRAM round trips (the MMA warp stalls once both TMEM buffers are full): pull this
        // warp's 32 x 128 block of layer k's output into L2 now, while the tensor core works on the tile
        const int64_t wr0 = int64_t(m0) + q * 32, c0 = int64_t(n0) + cbase;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int line = lane + 32 * i;                 // 32 rows x 4 lines of 128 bytes
          const int64_t rr = wr0 + (line >> 2), col = c0 + ((line & 3) << 5);
          if (rr < p.M && col < p.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pr_out + rr * p.pr_ldo + col));
        }
      }
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++chunk) {
        const int b = chunk & 1;
        mbar_wait(&tfull_bar[b], (chunk >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * BN + cbase);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {   // x16 loads: 128 accumulators + 32 in-flight values would not fit 168 registers
          uint32_t v[16];
          tmem_ld16(taddr + ch * 16, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[ch * 16 + j] += __uint_as_float(v[j]);   // round-to-nearest promotion
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1) mbar_arrive(&tempty_bar[b]); else mbar_arrive_leader(&tempty_bar[b]);
        }
      }

      const int64_t col0 = int64_t(n0) + cbase;       // first global column of this thread
      float* dst = p.C + row * p.ldc + col0;
      const bool vec_ok = (col0 + 128 <= p.N) && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
      if (EPI == EPI_ATOMIC) {
        if (row_ok) {
          if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 128; j += 4)
              red_add_v4(dst + j, acc[j] * inv, acc[j + 1] * inv, acc[j + 2] * inv, acc[j + 3] * inv);
          } else {
#pragma unroll
            for (int j = 0; j < 128; ++j)
              if (col0 + j < p.N) atomicAdd(dst + j, acc[j] * inv);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 128; j += 4) {                // (128-bit broadcast reads of the staged bias slice)
          const float4 bv = *reinterpret_cast<const float4*>(vec + cbase + j);
          acc[j] = fmaf(acc[j], inv, bv.x); acc[j + 1] = fmaf(acc[j + 1], inv, bv.y);
          acc[j + 2] = fmaf(acc[j + 2], inv, bv.z); acc[j + 3] = fmaf(acc[j + 3], inv, bv.w);
        }
        if (EPI == EPI_LOGITS) {
          float d1 = 0.f, d2 = 0.f;
          if (p.Cp <= 128) {
            // heads tile this thread's 128 columns: Cp divides 128 (checked on the host), hence is a power of two.  The head
            // width is a compile-time constant of each case (one store sequence per head instead of one per column: the
            // fully unrolled run-time version was 11k SASS instructions and stalled on instruction fetch)
            const float* va1 = vec + BN + cbase;
            const float* va2 = vec + 2 * BN + cbase;
            switch (p.Cp) {
              case 128: logits_heads<128>(acc, va1, va2, p, row, row_ok, col0); break;
              case 64: logits_heads<64>(acc, va1, va2, p, row, row_ok, col0); break;
              case 32: logits_heads<32>(acc, va1, va2, p, row, row_ok, col0); break;
              case 16: logits_heads<16>(acc, va1, va2, p, row, row_ok, col0); break;
              case 8: logits_heads<8>(acc, va1, va2, p, row, row_ok, col0); break;
              default: logits_heads<4>(acc, va1, va2, p, row, row_ok, col0); break;
            }
          } else {
            // Cp == 256 == BN: one head per tile, its two halves live in two warps -> combine through shared memory
            float e1[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {0.f, 0.f, 0.f, 0.f};   // four chains each instead of one of 128 FMAs
#pragma unroll
            for (int j = 0; j < 128; j += 4) {
              const float4 v1 = *reinterpret_cast<const float4*>(vec + BN + cbase + j);
              const float4 v2 = *reinterpret_cast<const float4*>(vec + 2 * BN + cbase + j);
              e1[0] = fmaf(acc[j], v1.x, e1[0]); e1[1] = fmaf(acc[j + 1], v1.y, e1[1]);
              e1[2] = fmaf(acc[j + 2], v1.z, e1[2]); e1[3] = fmaf(acc[j + 3], v1.w, e1[3]);
              e2[0] = fmaf(acc[j], v2.x, e2[0]); e2[1] = fmaf(acc[j + 1], v2.y, e2[1]);
              e2[2] = fmaf(acc[j + 2], v2.z, e2[2]); e2[3] = fmaf(acc[j + 3], v2.w, e2[3]);
            }
            d1 = (e1[0] + e1[1]) + (e1[2] + e1[3]);
            d2 = (e2[0] + e2[1]) + (e2[2] + e2[3]);
            const int r = q * 32 + lane;
            if (half == 1) {
              red[r] = d1;
              red[TC_BM + r] = d2;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * S::EPI_WARPS) : "memory");
            if (half == 0 && row_ok && n0 < p.N) {
              const int h = n0 / p.Cp;
              p.s_src[row * p.H + h] = d1 + red[r] + __ldg(p.b1 + h);
              p.s_dst[row * p.H + h] = d2 + red[TC_BM + r] + __ldg(p.b2 + h);
            }
          }
        }
        if (EPI == EPI_PREP) {
          // ---- layer k's prep, fused: acc holds gX = d loss / d act(out_k) of this thread's row and 128 columns.  The
          // matching out_k tile comes in through the warp's transposing buffer (coalesced 64-byte runs, 16 columns at a
          // time); G = gX * ELU'(out) replaces acc, Drow[row, h] = <G, out - bias> is summed per head of layer k. ----
          float* stg = stgbuf + ew * (32 * S::STG_ROW);
          const int64_t wrow0 = int64_t(m0) + q * 32;
          const int CPk = p.pr_C;                       // 8, 16, 32, 64, 128 or 256 (== BN): checked on the host
          float dh[16];                                 // per-head partial dot products (up to 128 / 8 heads per thread)
#pragma unroll
          for (int t = 0; t < 16; ++t) dh[t] = 0.f;
          float4 nxt[4];                                // chunk cc + 16 travels while chunk cc is transposed and consumed
          auto load_chunk = [&](int cc) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int idx = lane + 32 * i, r = idx >> 2, c4 = (idx & 3) << 2;
              const int64_t rr = wrow0 + r, col = col0 + cc + c4;
              nxt[i] = (rr < p.M && col + 4 <= p.N) ? ldg4(p.pr_out + rr * p.pr_ldo + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          };
          load_chunk(0);
#pragma unroll
          for (int cc = 0; cc < 128; cc += 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int idx = lane + 32 * i, r = idx >> 2, c4 = (idx & 3) << 2;
              *reinterpret_cast<float4*>(stg + r * S::STG_ROW + c4) = nxt[i];
            }
            if (cc + 16 < 128) load_chunk(cc + 16);
            __syncwarp();
            float dc = 0.f;                             // this 16-column chunk's contribution (CPk >= 16: one head)
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 o = *reinterpret_cast<const float4*>(stg + lane * S::STG_ROW + j);
              const float ov[4] = {o.x, o.y, o.z, o.w};
              float d4 = 0.f;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float g = acc[cc + j + t];
                if (p.pr_act) g *= elu_grad(ov[t]);
                acc[cc + j + t] = g;
                d4 = fmaf(g, ov[t] - vec[BN + cbase + cc + j + t], d4);
              }
              if (CPk == 8) dh[(cc + j) >> 3] += d4;    // two 4-column groups per head
              else dc += d4;
            }
            if (CPk >= 128) dh[0] += dc;
            else if (CPk == 64) dh[cc >> 6] += dc;
            else if (CPk == 32) dh[cc >> 5] += dc;
            else if (CPk == 16) dh[cc >> 4] += dc;
            __syncwarp();
          }
          if (CPk == 256) {                             // one head per 256-wide tile: combine the two column halves
            const int r = q * 32 + lane;
            if (half == 1) red[r] = dh[0];
            asm volatile("bar.sync 1, %0;" ::"n"(32 * S::EPI_WARPS) : "memory");
            if (half == 0 && row_ok && n0 < p.N) {
              const int64_t item = row * p.pr_H + n0 / 256;
              p.pr_rowrec[item] = make_float4(__ldg(p.pr_s_dst + item), __ldg(p.pr_rowmax + item),
                                              1.f / (__ldg(p.pr_rowsum + item) + 1e-16f), dh[0] + red[r]);
            }
          } else if (row_ok) {
            const int nh = CPk >= 128 ? 1 : 128 / CPk;  // heads inside this thread's 128 columns
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const int64_t col = col0 + int64_t(t) * CPk;
              if (t < nh && col < p.N) {
                const int64_t item = row * p.pr_H + col / CPk;
                p.pr_rowrec[item] = make_float4(__ldg(p.pr_s_dst + item), __ldg(p.pr_rowmax + item),
                                                1.f / (__ldg(p.pr_rowsum + item) + 1e-16f), dh[t]);
              }
            }
          }
        }
        const bool al_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
        if (al_ok) {
          // Coalesced stores.  A thread owns one ROW (TMEM lane) and 128 of its columns, so a direct 128-bit store hits 32
          // different lines per instruction, 16 bytes each (ncu on the 2.4 M-node graph's 100 -> 512 projection: l1tex at
          // 68 %, the tensor pipe at 15 %, 4.9 GB written in 3 ms).  The warp's 32 x 128 block goes through a per-warp
          // shared buffer 16 columns at a time and leaves as 64-byte runs: 8 lines per instruction, whole sectors.
          float* stg = stgbuf + ew * (32 * S::STG_ROW);
          const int64_t wrow0 = int64_t(m0) + q * 32;
          // the common case — whole 128 columns inside C, no bf16 copy, no peers: the lane's four row pointers (rows
          // lane / 4 + 8 i, its 16-byte piece of a 64-byte run) and row predicates are computed ONCE per tile and every store
          // is pointer + immediate.  (The general loop below recomputed a 64-bit address and three predicates per store: ncu
          // showed the tile-end phase spread over IMAD.X / spill reloads, 11 us per 128 x 256 tile.)
          if (EPI != EPI_PREP && col0 + 128 <= p.N && p.c16 == nullptr && p.n_peer == 0) {
            const int r0 = lane >> 2, c4 = (lane & 3) << 2;
            float* d0 = p.C + (wrow0 + r0) * p.ldc + col0 + c4;
            const int64_t step = 8 * p.ldc;
            const float* sl = stg + r0 * S::STG_ROW + c4;
            const int64_t left = p.M - wrow0 - r0;               // rows r0 + 8 i exist while 8 i < left
#pragma unroll
            for (int cc = 0; cc < 128; cc += 16) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(stg + lane * S::STG_ROW + j) =
                    make_float4(acc[cc + j], acc[cc + j + 1], acc[cc + j + 2], acc[cc + j + 3]);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(sl + 8 * i * S::STG_ROW);
                if (8 * i < left) *reinterpret_cast<float4*>(d0 + i * step + cc) = v;
              }
              __syncwarp();
            }
            continue;
          }
#pragma unroll
          for (int cc = 0; cc < 128; cc += 16) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(stg + lane * S::STG_ROW + j) =
                  make_float4(acc[cc + j], acc[cc + j + 1], acc[cc + j + 2], acc[cc + j + 3]);
            __syncwarp();
            float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);        // EPI_PREP: column sums of G (layer k's g_bias)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int idx = lane + 32 * i, r = idx >> 2, c4 = (idx & 3) << 2;
              const float4 v = *reinterpret_cast<const float4*>(stg + r * S::STG_ROW + c4);
              const int64_t rr = wrow0 + r, col = col0 + cc + c4;
              if (EPI == EPI_PREP && rr < p.M) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
              if (rr < p.M) {
                float* d = p.C + rr * p.ldc + col;
                if (col + 4 <= p.N) {
                  *reinterpret_cast<float4*>(d) = v;
                  if (p.c16) *reinterpret_cast<uint2*>(p.c16 + rr * p.ldc + col) = pack_bf16x4(v.x, v.y, v.z, v.w);
                  for (int k = 0; k < p.n_peer; ++k)    // fused all-gather: the same 64-byte runs over NVLink
                    *reinterpret_cast<float4*>(p.peer_c[k] + rr * p.ldc + col) = v;
                } else if (col < p.N) {                 // the (rare) ragged group of a column tail
                  d[0] = v.x;
                  if (col + 1 < p.N) d[1] = v.y;
                  if (col + 2 < p.N) d[2] = v.z;
                  for (int k = 0; k < p.n_peer; ++k) {
                    float* dk = p.peer_c[k] + rr * p.ldc + col;
                    dk[0] = v.x;
                    if (col + 1 < p.N) dk[1] = v.y;
                    if (col + 2 < p.N) dk[2] = v.z;
                  }
                }
              }
            }
            if (EPI == EPI_PREP) {                      // lanes with equal (lane & 3) hold the same 4 columns: combine, then
#pragma unroll
              for (int o2 = 4; o2 < 32; o2 <<= 1) {     // lanes 0..3 add the warp's 32-row partial sums to g_bias
                cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o2); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o2);
                cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o2); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o2);
              }
              const int64_t col = col0 + cc + (lane << 2);
              if (lane < 4 && col + 4 <= p.N) {
                atomicAdd(p.pr_g_bias + col, cs.x); atomicAdd(p.pr_g_bias + col + 1, cs.y);
                atomicAdd(p.pr_g_bias + col + 2, cs.z); atomicAdd(p.pr_g_bias + col + 3, cs.w);
              }
            }
            __syncwarp();
          }
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 128; ++j)
            if (col0 + j < p.N) dst[j] = acc[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();                      // the peer may still be signalling this CTA's barriers / reading its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN, CG>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ split pre-passes
// max |src| as a uint bit pattern (non-negative floats order like unsigned integers; NaN sorts above inf and is kept)
// (no activation variant: |ELU(x)| <= |x|, so max|x| is a valid — at most slightly loose — bound for the scale)
__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols, uint32_t* __restrict__ out) {
  uint32_t m = 0;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = int64_t(gridDim.x) * blockDim.x;
  if (ld == cols && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const int64_t total = rows * cols, n4 = total >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int64_t t = tid; t < n4; t += nth) {
      const float4 v = __ldg(s4 + t);
      m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
    }
    for (int64_t t = (n4 << 2) + tid; t < total; t += nth) m = max(m, __float_as_uint(fabsf(__ldg(src + t))));
  } else {
    const int64_t total = rows * cols;
    for (int64_t t = tid; t < total; t += nth) {
      const int64_t r = t / cols, c = t - r * cols;
      m = max(m, __float_as_uint(fabsf(__ldg(src + r * ld + c))));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ uint32_t sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = max(m, sm[i]);
    if (m) atomicMax(out, m);
  }
}

// src [rows, cols] (row stride ld) -> hi / lo fp16 planes [rows, ldp]; pad columns [cols, ldp) are zero
template <int ACT>
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols, __half* __restrict__ hi,
             __half* __restrict__ lo, int64_t ldp, const uint32_t* __restrict__ amax_bits, float* __restrict__ inv_scale) {
  const float s = scale_from_amax(*amax_bits);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = 1.f / s;
  const int64_t groups = ldp >> 3, total = rows * groups;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
  // (row, column group) of item t are stepped incrementally: the 64-bit division per item made this pass issue-bound
  // (84 % issue slots busy at half the copy rate, ncu)
  const int64_t nth = int64_t(gridDim.x) * blockDim.x, t_first = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t dr = nth / groups, dg = nth - dr * groups;
  int64_t r = t_first / groups, g = t_first - r * groups;
  for (int64_t t = t_first; t < total; t += nth, r += dr, g += dg) {
    if (g >= groups) { g -= groups; ++r; }
    const int64_t c0 = g << 3;
    float v[8];
    const float* p = src + r * ld + c0;
    if (vec && c0 + 8 <= cols) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(p)), x1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? __ldg(p + j) : 0.f;
    }
    __align__(16) __half2 h[4];
    __align__(16) __half2 l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      split_half2((ACT == ACT_ELU ? elu_fwd(v[2 * j]) : v[2 * j]) * s, (ACT == ACT_ELU ? elu_fwd(v[2 * j + 1]) : v[2 * j + 1]) * s,
                  h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + r * ldp + c0) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + r * ldp + c0) = *reinterpret_cast<const uint4*>(l);
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

// 2-D fp16 plane [rows, cols] with row pitch `ld` elements; box = {64 cols (128 B, inner), box_rows}, SWIZZLE_128B;
// out-of-range parts of a box are zero-filled (K / M / N tails need no padding in memory)
static int make_map(CUtensorMap* m, const __half* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200GAT_E_UNSUPPORTED, "proj_tc: cuTensorMapEncodeTiled is not available");
  cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t gstr[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {cuuint32_t(TC_BK), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200GAT_E_SHAPE, "proj_tc: cuTensorMapEncodeTiled failed (%d)", int(r));
  return 0;
}

// split `src` (through the activation `act`) into blob b; amax_hint: device bound of max|src| (skips the amax pass)
static int launch_split(const float* src, int64_t ld, const Blob& b, cudaStream_t stream, int act = ACT_NONE,
                        const uint32_t* amax_hint = nullptr) {
  const int64_t cap = int64_t(sm_count()) * 8;
  int rc;
  if (!amax_hint) {
    cudaError_t ce = cudaMemsetAsync(b.base, 0, 8, stream);
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "proj_tc: memset: %s", cudaGetErrorString(ce));
    const int64_t want_a = ceil_div(b.rows * b.cols, 256 * 4);
    amax_kernel<<<static_cast<int>(want_a < cap ? (want_a > 0 ? want_a : 1) : cap), 256, 0, stream>>>(src, ld, b.rows, b.cols, b.amax_bits());
    if ((rc = check_launch("amax_kernel"))) return rc;
    amax_hint = b.amax_bits();
  }
  const int64_t want_s = ceil_div(b.rows * (b.ldp >> 3), 256);
  const int blocks = static_cast<int>(want_s < 2 * cap ? (want_s > 0 ? want_s : 1) : 2 * cap);
  if (act == ACT_ELU)
    split_kernel<ACT_ELU><<<blocks, 256, 0, stream>>>(src, ld, b.rows, b.cols, b.hi(), b.lo(), b.ldp, amax_hint, b.inv_scale());
  else
    split_kernel<ACT_NONE><<<blocks, 256, 0, stream>>>(src, ld, b.rows, b.cols, b.hi(), b.lo(), b.ldp, amax_hint, b.inv_scale());
  return check_launch("split_kernel");
}

// co-resident CTA pairs of the cta_group::2 kernel (cached per instantiation; its shared-memory attribute must be set first)
template <int BN, bool A_MN, bool B_MN, int EPI>
static int max_active_pairs() {
  using S = TcCfg<BN, 2>;
  static int max_pairs = 0;
  static std::once_flag once2;
  std::call_once(once2, [&] {
    auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI, 2>;
    (void)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(2 * sm_count());
    q.blockDim = dim3(S::THREADS);
    q.dynamicSmemBytes = S::TOTAL;
    cudaLaunchAttribute qa;
    qa.id = cudaLaunchAttributeClusterDimension;
    qa.val.clusterDim.x = 2; qa.val.clusterDim.y = 1; qa.val.clusterDim.z = 1;
    q.attrs = &qa;
    q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) n = sm_count() / 2 - 4;
    max_pairs = n;
    (void)cudaGetLastError();
  });
  return max_pairs;
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CG>
static int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                     TcGemmParams p, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, EPI, CG>;
  using S = TcCfg<BN, CG>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  });
  if (attr_err != cudaSuccess)
    return fail(static_cast<int>(attr_err), "proj_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  p.m_tiles = static_cast<int>(ceil_div(p.M, TC_BM * CG));
  p.n_tiles = static_cast<int>(ceil_div(p.N, BN));
  const int64_t items = int64_t(p.m_tiles) * p.n_tiles * p.splits;
  if constexpr (CG == 1) {
    const int grid = static_cast<int>(items < sm_count() ? items : sm_count());
    kern<<<grid, S::THREADS, S::TOTAL, stream>>>(ah, al, bh, bl, p);
  } else {
    // persistent pairs: as many clusters as can be co-resident (GPCs with an odd SM count leave an SM without a partner)
    const int max_pairs = max_active_pairs<BN, A_MN, B_MN, EPI>();
    const int64_t units = max_pairs;
    const int grid = static_cast<int>((items < units ? items : units) * CG);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(S::THREADS);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ah, al, bh, bl, p);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "proj_tc: cluster launch: %s", cudaGetErrorString(e));
  }
  return check_launch("gemm_tc_kernel");
}

// The CTA-pair kernel is OPT-IN (B200GAT_GEMM_PAIR=1): it passes every parity test, but measured on the PPI-shaped
// GEMMs it is no faster than the single-CTA kernel (forward 229.9 vs 227 us, gX 298 vs 303, gW 271 vs 296) — at
// ~1.56 PFLOP/s of fp16 MMA work those kernels already run at 92 % of the pool's measured cuBLAS bf16 burst rate
// (1.69 PFLOP/s; the part is power-limited well below the nominal 2.25), so halving the operand bytes buys nothing.
// Re-measured after the store epilogue was coalesced (it had hidden the difference): with K >= 1024 the pair is ahead in
// the backward GEMMs of the PPI-shaped layers (gX + gW 629 -> 597 us, 519 -> 475 us) and level in the forward; with K = 512
// (2.4 M-node graph) it is 1-2 % behind.  Default: pairs for the backward GEMMs with K >= 1024; B200GAT_GEMM_PAIR=0|1 forces
// either for every GEMM.
static bool use_pair(int64_t M, int bn, int64_t K, bool backward) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B200GAT_GEMM_PAIR");
    on = !e ? 2 : ((e[0] == '1') ? 1 : 0);
  }
  if (!on || bn != 256 || M < 256) return false;
  return on == 1 || (backward && K >= 1024);   // `backward`: the gX / gW GEMMs (B operand MN-major)
}

// C[M,N] (=, +bias | +=) A · B^T.  A is the blob of a [M,K] (K-major) or [K,M] (A_MN) tensor, B of a [N,K] (K-major)
// or [K,N] (B_MN) tensor; `splits` > 1 requires EPI_ATOMIC on a zeroed C.
template <bool A_MN, bool B_MN, int EPI>
static int gemm_blobs(const Blob& A, const Blob& B, int64_t M, int64_t N, int64_t K, TcGemmParams p, int splits,
                      cudaStream_t stream) {
  p.M = M; p.N = N;
  p.k_blocks = static_cast<int>(ceil_div(K, TC_BK));
  if (splits < 1) splits = 1;
  p.k_blocks_per_split = static_cast<int>(ceil_div(p.k_blocks, splits));
  p.splits = static_cast<int>(ceil_div(p.k_blocks, p.k_blocks_per_split));
  { static int kc = 0; if (!kc) { const char* e = getenv("B200GAT_TC_KC"); kc = e ? atoi(e) : TC_KC; if (kc < 1) kc = TC_KC; } p.kc = kc; }
  p.inv_a = A.inv_scale();
  p.inv_b = B.inv_scale();
  const int bn = N > 128 ? 256 : 128;
  CUtensorMap ah, al, bh, bl;
  int rc;
  const bool pair = use_pair(M, bn, K, B_MN);
  const int a_box = A_MN ? 64 : TC_BM, b_box = B_MN ? 64 : (pair ? bn / 2 : bn);
  if ((rc = make_map(&ah, A.hi(), A.rows, A.cols, A.ldp, a_box))) return rc;
  if ((rc = make_map(&al, A.lo(), A.rows, A.cols, A.ldp, a_box))) return rc;
  if ((rc = make_map(&bh, B.hi(), B.rows, B.cols, B.ldp, b_box))) return rc;
  if ((rc = make_map(&bl, B.lo(), B.rows, B.cols, B.ldp, b_box))) return rc;
  if (pair) return launch_tc<256, A_MN, B_MN, EPI, 2>(ah, al, bh, bl, p, stream);
  if (bn == 256) return launch_tc<256, A_MN, B_MN, EPI, 1>(ah, al, bh, bl, p, stream);
  return launch_tc<128, A_MN, B_MN, EPI, 1>(ah, al, bh, bl, p, stream);
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

// Can the gX GEMM of `consumer` (layer k+1) run the prep pass of `producer` (layer k) in its epilogue (EPI_PREP)?  The
// producer must be concat-like with a head width the epilogue's per-thread 128-column block can sum (8 .. 128, or one whole
// 256-wide tile) and its output must be exactly the consumer's input.
bool proj_tc_can_fuse_prep(const b200gat_layer& consumer, int64_t N, const b200gat_layer& producer) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("B200GAT_NO_FUSE_PREP"); off = (e && e[0] == '1') ? 1 : 0; }
  if (off || !proj_tc_bwd_supported(consumer, N)) return false;
  const int64_t C = producer.out_channels, H = producer.heads, F = consumer.in_channels;
  if (!(producer.concat || H == 1) || H * C != F || F % 4 != 0) return false;
  if (C == 256) return F > 128;                          // BN == 256: one head per tile
  return C == 8 || C == 16 || C == 32 || C == 64 || C == 128;
}

size_t proj_tc_split_bytes(const b200gat_layer& L, int64_t N) {
  // [x blob][W blob]: the backward's gX GEMM reads the same W planes as the forward (MN-major), so they are kept too
  return proj_tc_fwd_supported(L, N) ? blob_bytes(N, L.in_channels) + blob_bytes(L.heads * L.c_pad, L.in_channels) : 0;
}

size_t proj_tc_fwd_workspace_bytes(const b200gat_layer& L, int64_t N) {
  if (!proj_tc_fwd_supported(L, N)) return 0;
  return blob_bytes(N, L.in_channels) + blob_bytes(L.heads * L.c_pad, L.in_channels);
}

static int check_ws(const void* ws, size_t have, size_t need, const char* what) {
  B200GAT_REQUIRE(ws && have >= need, B200GAT_E_WORKSPACE, "%s: workspace %zu < %zu bytes", what, have, need);
  B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, B200GAT_E_ALIGN, "%s: workspace must be 256-byte aligned", what);
  return 0;
}

int proj_tc_fwd(const b200gat_proj_fwd_args& a, cudaStream_t stream) {
  const b200gat_layer& L = a.layer;
  const int64_t N = a.num_nodes, F = L.in_channels, H = L.heads, Cp = L.c_pad, Dp = H * Cp;
  const size_t xb = blob_bytes(N, F), wb = blob_bytes(Dp, F);
  int rc;
  // the splits of x and W are written into the caller's x_split buffer when given (kept for b200gat_proj_bwd), else into
  // the workspace
  const bool keep = a.x_split != nullptr;
  if (keep) {
    B200GAT_REQUIRE(a.x_split_bytes >= xb + wb, B200GAT_E_WORKSPACE, "proj_fwd: x_split %zu < %zu bytes", a.x_split_bytes, xb + wb);
    B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(a.x_split) & 255u) == 0, B200GAT_E_ALIGN, "proj_fwd: x_split must be 256-byte aligned");
  }
  if (!keep && (rc = check_ws(a.workspace, a.workspace_bytes, xb + wb, "proj_fwd"))) return rc;
  char* base = static_cast<char*>(keep ? a.x_split : a.workspace);
  const Blob X = make_blob(base, N, F);
  const Blob W = make_blob(base + xb, Dp, F);
  if ((rc = launch_split(a.x, a.ldx, X, stream, a.x_activation, a.x_amax))) return rc;
  if ((rc = launch_split(a.w, F, W, stream))) return rc;
  TcGemmParams p{};
  p.C = a.wh; p.ldc = Dp; p.bias = a.bw;
  p.a1 = a.a1; p.a2 = a.a2; p.b1 = a.b1; p.b2 = a.b2; p.s_src = a.s_src; p.s_dst = a.s_dst;
  p.H = static_cast<int>(H); p.Cp = static_cast<int>(Cp);
  B200GAT_REQUIRE(a.num_peers >= 0 && a.num_peers <= B200GAT_MAX_PEERS, B200GAT_E_SHAPE, "proj_fwd: num_peers out of range");
  p.n_peer = a.num_peers;
  B200GAT_REQUIRE(!a.wh_bf16 || (reinterpret_cast<uintptr_t>(a.wh_bf16) & 7u) == 0, B200GAT_E_ALIGN, "proj_fwd: wh_bf16 must be 8-byte aligned");
  p.c16 = static_cast<uint16_t*>(a.wh_bf16);
  for (int k = 0; k < a.num_peers; ++k) {
    B200GAT_REQUIRE(a.wh_peers[k] && aligned16(a.wh_peers[k]), B200GAT_E_ALIGN, "proj_fwd: wh_peers[%d] NULL or not 16-byte aligned", k);
    p.peer_c[k] = a.wh_peers[k];
  }
  const int bn = Dp > 128 ? 256 : 128;
  // heads must not straddle an epilogue thread's 128 columns, or be exactly one 256-wide tile
  const bool fuse_logits = (Cp <= 128 && 128 % Cp == 0) || (Cp == 256 && bn == 256);
  if (fuse_logits) return gemm_blobs<false, false, EPI_LOGITS>(X, W, N, Dp, F, p, 1, stream);
  if ((rc = gemm_blobs<false, false, EPI_STORE>(X, W, N, Dp, F, p, 1, stream))) return rc;
  return launch_logits(a, stream);
}

size_t proj_tc_bwd_workspace_bytes(const b200gat_layer& L, int64_t N) {
  if (!proj_tc_bwd_supported(L, N)) return 0;
  const int64_t F = L.in_channels, Dp = L.heads * L.c_pad;
  return blob_bytes(N, Dp) + blob_bytes(Dp, F) + blob_bytes(N, F);
}

int proj_tc_bwd(const b200gat_proj_bwd_args& a, cudaStream_t stream) {
  const b200gat_layer& L = a.layer;
  const int64_t N = a.num_nodes, F = L.in_channels, Dp = L.heads * L.c_pad;
  const size_t gb = blob_bytes(N, Dp), wb = blob_bytes(Dp, F), xb = blob_bytes(N, F);
  const bool have_x = a.x_split != nullptr;
  if (have_x) {
    B200GAT_REQUIRE(a.x_split_bytes >= xb + wb, B200GAT_E_WORKSPACE, "proj_bwd: x_split %zu < %zu bytes", a.x_split_bytes, xb + wb);
    B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(a.x_split) & 255u) == 0, B200GAT_E_ALIGN, "proj_bwd: x_split must be 256-byte aligned");
  }
  const bool have_g = a.g_t_split != nullptr;
  if (have_g) {
    B200GAT_REQUIRE(a.g_t_split_bytes >= gb, B200GAT_E_WORKSPACE, "proj_bwd: g_t_split %zu < %zu bytes", a.g_t_split_bytes, gb);
    B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(a.g_t_split) & 255u) == 0, B200GAT_E_ALIGN, "proj_bwd: g_t_split must be 256-byte aligned");
  }
  B200GAT_REQUIRE((have_g || a.g_t) && (have_x || a.x), B200GAT_E_NULL, "proj_bwd: NULL pointer");
  int rc;
  // workspace layout: [W blob][gT blob unless g_t_split][x blob unless x_split]
  if ((rc = check_ws(a.workspace, a.workspace_bytes, wb + (have_g ? 0 : gb) + (have_x ? 0 : xb), "proj_bwd"))) return rc;
  char* base = static_cast<char*>(a.workspace);
  // (W planes: the forward's, behind the x blob in x_split — same W, same layout, read MN-major here — else split again)
  const Blob W = make_blob(have_x ? static_cast<char*>(const_cast<void*>(a.x_split)) + xb : base, Dp, F);
  const Blob G = make_blob(have_g ? const_cast<void*>(a.g_t_split) : static_cast<void*>(base + wb), N, Dp);
  const Blob X = make_blob(have_x ? const_cast<void*>(a.x_split) : static_cast<void*>(base + wb + (have_g ? 0 : gb)), N, F);
  if (!have_g && (rc = launch_split(a.g_t, Dp, G, stream))) return rc;
  if (!have_x && (rc = launch_split(a.x, a.ldx, X, stream, a.x_activation))) return rc;
  const bool want_gx = a.g_x && a.parts != B200GAT_PROJ_BWD_GW, want_gw = a.parts != B200GAT_PROJ_BWD_GX;
  if (want_gx) {
    // gX[N,F] = gT[N,Dp] · W[Dp,F] : K = Dp; B = the W planes read MN-major (F contiguous)
    if (!have_x && (rc = launch_split(a.w, F, W, stream))) return rc;
    TcGemmParams p{};
    p.C = a.g_x; p.ldc = a.ldgx;
    if (a.fuse_prep) {
      // the producer layer's prep pass rides in this GEMM's epilogue (b200gat_proj_bwd_args.fuse_prep)
      const b200gat_edge_bwd_prep_args& f = *a.fuse_prep;
      B200GAT_REQUIRE(proj_tc_can_fuse_prep(L, N, f.layer), B200GAT_E_UNSUPPORTED,
                      "proj_bwd: fuse_prep is not offered for this pair of layers (b200gat_proj_bwd_can_fuse_prep)");
      B200GAT_REQUIRE(f.num_rows == N && f.out && f.bias && f.s_dst && f.rowmax && f.rowsum && f.rowrec && f.g_bias,
                      B200GAT_E_NULL, "proj_bwd: fuse_prep: NULL pointer or row count mismatch");
      B200GAT_REQUIRE(f.ldo >= F && f.ldo % 4 == 0 && aligned16(f.out) && aligned16(f.rowrec) && a.ldgx % 4 == 0 && aligned16(a.g_x),
                      B200GAT_E_ALIGN, "proj_bwd: fuse_prep needs 16-byte aligned out / g_x rows");
      B200GAT_REQUIRE(f.out_activation == ACT_NONE || f.out_activation == ACT_ELU, B200GAT_E_UNSUPPORTED,
                      "proj_bwd: fuse_prep: unknown out_activation %d", f.out_activation);
      cudaError_t ce0 = cudaMemsetAsync(f.g_bias, 0, size_t(F) * sizeof(float), stream);
      if (ce0 != cudaSuccess) return fail(static_cast<int>(ce0), "proj_bwd: memset: %s", cudaGetErrorString(ce0));
      p.pr_out = f.out; p.pr_ldo = f.ldo; p.pr_bias = f.bias;
      p.pr_s_dst = f.s_dst; p.pr_rowmax = f.rowmax; p.pr_rowsum = f.rowsum;
      p.pr_rowrec = reinterpret_cast<float4*>(f.rowrec); p.pr_g_bias = f.g_bias;
      p.pr_H = static_cast<int>(f.layer.heads); p.pr_C = static_cast<int>(f.layer.out_channels); p.pr_act = f.out_activation;
      if ((rc = gemm_blobs<false, true, EPI_PREP>(G, W, N, F, Dp, p, 1, stream))) return rc;
    } else if ((rc = gemm_blobs<false, true, EPI_STORE>(G, W, N, F, Dp, p, 1, stream))) return rc;
  }
  if (!want_gw) return 0;
  // gW[Dp,F] = gT^T · X : K = nodes; both operands MN-major straight from the row-major planes; split-K across CTAs
  // with a red.global.add epilogue
  cudaError_t ce = cudaMemsetAsync(a.g_w, 0, size_t(Dp) * F * sizeof(float), stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "proj_bwd: memset: %s", cudaGetErrorString(ce));
  const int bn = F > 128 ? 256 : 128;
  const int64_t tiles = ceil_div(Dp, TC_BM) * ceil_div(F, bn);
  const int64_t kblocks = ceil_div(N, TC_BK);
  // split-K factor: the persistent units (CTAs, or CTA pairs) walk tiles x splits items in waves; pick the number of waves
  // w (splits = floor(units * w / tiles): the last wave is full) that minimises  w * (k-blocks per item + epilogue), the
  // red.global.add epilogue of an item costed at ~4 k-blocks.  (It was "~2 items per SM": 19 splits x 16 tiles = 304 items
  // on 74 pairs = 4.1 -> 5 waves on the PPI-shaped layers, the last one nearly empty.)
  const bool pair = use_pair(Dp, bn, N, true);
  const int64_t units = pair ? (bn == 256 ? max_active_pairs<256, true, true, EPI_ATOMIC>() : 1) : sm_count();
  const int64_t work_tiles = pair ? ceil_div(Dp, 2 * TC_BM) * ceil_div(F, bn) : tiles;
  const int64_t max_splits = kblocks / (2 * TC_KC) > 1 ? kblocks / (2 * TC_KC) : 1;
  int64_t splits = 1, best = -1;
  for (int64_t w = 1; w <= 8; ++w) {
    int64_t sp = units * w / work_tiles;
    if (sp < 1) continue;
    if (sp > max_splits) sp = max_splits;
    const int64_t waves = ceil_div(work_tiles * sp, units);
    const int64_t cost = waves * (ceil_div(kblocks, sp) + 4);
    if (best < 0 || cost < best) { best = cost; splits = sp; }
  }
  TcGemmParams p{};
  p.C = a.g_w; p.ldc = F;
  return gemm_blobs<true, true, EPI_ATOMIC>(G, X, Dp, F, N, p, static_cast<int>(splits), stream);
}

}  // namespace b200gat
