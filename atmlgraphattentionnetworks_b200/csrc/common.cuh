// Shared helpers for the b200gat kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200gat.h"

namespace b200gat {

// ---- error reporting (thread-local message, C-ABI return codes) ----
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define B200GAT_REQUIRE(cond, code, ...) \
  do { if (!(cond)) return ::b200gat::fail((code), __VA_ARGS__); } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();   // cached multiprocessor count of the current device (148 on B200)
int64_t l2_bytes();   // cached L2 size of the current device (126 MB on B200)

// Edge-kernel schedule (see b200gat_graph.span): `row_bytes` = bytes of one gathered row.  Streaming when the rows an
// edge can reach from the row being processed do not fit half of L2; B200GAT_EDGE_SCHEDULE=cached|stream overrides.
bool edge_schedule_streaming(int64_t span, int64_t row_bytes);

inline int validate_layer(const b200gat_layer& L) {
  B200GAT_REQUIRE(L.in_channels > 0 && L.out_channels > 0 && L.heads > 0, B200GAT_E_SHAPE,
                  "layer: in_channels/out_channels/heads must be positive");
  B200GAT_REQUIRE(L.c_pad == ((L.out_channels + 3) / 4) * 4, B200GAT_E_SHAPE,
                  "layer: c_pad must be round_up(out_channels, 4)");
  B200GAT_REQUIRE(L.c_pad <= 512, B200GAT_E_UNSUPPORTED, "layer: out_channels > 512 per head is not supported");
  B200GAT_REQUIRE(L.heads <= 1024, B200GAT_E_UNSUPPORTED, "layer: more than 1024 heads is not supported");
  B200GAT_REQUIRE(L.logit_activation >= B200GAT_LOGIT_LEAKY_RELU && L.logit_activation <= B200GAT_LOGIT_HEAD_SOFTMAX,
                  B200GAT_E_UNSUPPORTED, "layer: unknown logit_activation %d", L.logit_activation);
  return 0;
}

inline int validate_graph(const b200gat_graph& g) {
  B200GAT_REQUIRE(g.num_nodes >= 0 && g.num_edges >= g.num_nodes, B200GAT_E_SHAPE,
                  "graph: num_edges must include the N self loops");
  B200GAT_REQUIRE(g.num_edges < (int64_t(1) << 31), B200GAT_E_SHAPE, "graph: E + N must be < 2^31");
  if (g.num_nodes == 0) return 0;
  B200GAT_REQUIRE(g.rowptr && g.col && g.colptr && g.crow, B200GAT_E_NULL, "graph: NULL array");
  return 0;
}

// ---- device helpers ----
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// One gathered 128-bit slice: `base` = this lane's byte pointer into row 0 (column slice already applied), `row` a
// non-negative row id, `row_bytes` < 2^32.  unsigned 32 x 32 -> 64 multiply-add = ONE IMAD.WIDE.U32 per gather; the
// int64 index arithmetic it replaces was 11-13 integer instructions per gathered row (ncu source view: 55 % of the
// warp instructions of edge_fwd were integer / address work at a 66 % issue-slot utilisation).
__device__ __forceinline__ float4 ldg4_row(const char* base, int row, uint32_t row_bytes) {
  return __ldg(reinterpret_cast<const float4*>(base + static_cast<uint64_t>(static_cast<uint32_t>(row)) * row_bytes));
}

// ---- optional bf16 storage of the GATHERED rows (Wh in the forward, the gradient rows G in the backward): halves the bytes
// every edge pulls through L2 / HBM (and over NVLink in the row-partitioned mode) at a separately stated tolerance
// (DESIGN.md §4.11); all arithmetic stays fp32.  ROW16 = false is the default fp32 path, untouched.
template <bool ROW16>
__device__ __forceinline__ const char* row_base(const float* rows32, const void* rows16, int64_t elem) {
  return ROW16 ? static_cast<const char*>(rows16) + elem * 2 : reinterpret_cast<const char*>(rows32 + elem);
}
template <bool ROW16>
__device__ __forceinline__ float4 gather4(const char* base, int row, uint32_t row_bytes) {
  if (!ROW16) return ldg4_row(base, row, row_bytes);
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(base + static_cast<uint64_t>(static_cast<uint32_t>(row)) * row_bytes));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
// four fp32 values -> four bf16 (round to nearest even), packed in 8 bytes
__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  auto rn = [](float f) -> uint32_t {
    const uint32_t u = __float_as_uint(f);
    if ((u & 0x7f800000u) == 0x7f800000u) return u >> 16;                       // inf / nan: truncate
    return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
  };
  return make_uint2(rn(a) | (rn(b) << 16), rn(c) | (rn(d) << 16));
}

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : slope * z; }

// logit activation e = f(z) and f'(z) (include/b200gat.h B200GAT_LOGIT_*).  GENERIC = false is the LeakyReLU of GAT.py:30
// with no trace of the other variants in the generated code (the CSC pass's code generation is fragile).
template <bool GENERIC>
__device__ __forceinline__ float logit_act(float z, float slope, int act) {
  if (!GENERIC) return leaky(z, slope);
  if (act == B200GAT_LOGIT_TANH) return tanhf(z);
  if (act == B200GAT_LOGIT_LOGSIGMOID) return fminf(z, 0.f) - log1pf(expf(-fabsf(z)));   // -softplus(-z), stable
  return leaky(z, slope);
}
template <bool GENERIC>
__device__ __forceinline__ float logit_act_grad(float z, float slope, int act) {
  if (!GENERIC) return z > 0.f ? 1.f : slope;
  if (act == B200GAT_LOGIT_TANH) { const float t = tanhf(z); return 1.f - t * t; }
  if (act == B200GAT_LOGIT_LOGSIGMOID) return 1.f / (1.f + expf(z));                     // sigmoid(-z)
  return z > 0.f ? 1.f : slope;
}

// ---- attention dropout (GAT.py:61): keep-multiplier of coefficient (edge e, head h), e = position in the ORIGINAL
// [edges ; loops] order.  Either read from a caller-supplied [E', H] tensor (parity tests) or generated in the kernel:
// Philox4x32-10, key = seed[0], counter = (e, h / 4, seed[1]), word h % 4; keep iff word >= p * 2^32.  The forward (CSR
// order) and the backward (CSC order) regenerate the identical multiplier from (seed, e, h): no [E', H] tensor exists.
struct DropoutSpec {
  const float* mask;       // tensor mode, or nullptr
  const uint64_t* seed;    // Philox mode (mask == nullptr): DEVICE pointer to two words, or nullptr (dropout off)
  uint32_t threshold;      // keep iff word >= threshold
  float scale;             // 1 / (1 - p); 0 when p >= 1
  __host__ __device__ bool active() const { return mask != nullptr || seed != nullptr; }
};
// host: validate the ABI's (mask, b200gat_dropout) pair
int make_dropout(const float* mask, const b200gat_dropout& d, DropoutSpec* out);

struct DropoutKey { uint32_t k0, k1, c2, c3; };
__device__ __forceinline__ DropoutKey dropout_key(const DropoutSpec& d) {
  DropoutKey k{0u, 0u, 0u, 0u};
  if (d.mask == nullptr && d.seed != nullptr) {
    const unsigned long long s0 = __ldg(reinterpret_cast<const unsigned long long*>(d.seed));
    const unsigned long long s1 = __ldg(reinterpret_cast<const unsigned long long*>(d.seed) + 1);
    k.k0 = static_cast<uint32_t>(s0); k.k1 = static_cast<uint32_t>(s0 >> 32);
    k.c2 = static_cast<uint32_t>(s1); k.c3 = static_cast<uint32_t>(s1 >> 32);
  }
  return k;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// call only when d.active()
__device__ __forceinline__ float dropout_mult(const DropoutSpec& d, const DropoutKey& key, int e, int h, int H) {
  if (d.mask) return __ldg(d.mask + int64_t(e) * H + h);
  const uint4 r = philox4x32_10(static_cast<uint32_t>(e), static_cast<uint32_t>(h) >> 2, key.c2, key.c3, key.k0, key.k1);
  const int s = h & 3;
  const uint32_t w = s == 0 ? r.x : (s == 1 ? r.y : (s == 2 ? r.z : r.w));
  return w >= d.threshold ? d.scale : 0.f;
}

// B200GAT_LOGIT_HEAD_SOFTMAX (run_act_func_experiment.py:111, nn.Softmax() on the [E', H] logits = softmax over the HEADS
// of one edge): e_h = exp(z_h - max_h' z_h') / sum_h' exp(z_h' - max), z_h' = a[h' * a_stride] + b[h'].  The two rows are the
// destination's s_dst values (stride 1, or stride 4 inside the backward's 16-byte row records) and the source's s_src.
__device__ __forceinline__ float head_softmax(const float* a, int a_stride, const float* b, int h, int H) {
  float mx = -INFINITY;
  for (int t = 0; t < H; ++t) mx = fmaxf(mx, __ldg(a + t * a_stride) + __ldg(b + t));
  float sum = 0.f, mine = 0.f;
  for (int t = 0; t < H; ++t) {
    const float v = expf(__ldg(a + t * a_stride) + __ldg(b + t) - mx);
    sum += v;
    if (t == h) mine = v;
  }
  return mine / sum;
}
// e = f(s_dst[i,h] + s_src[j,h]) for any logit activation; sd = s_dst[i,h] (already loaded by the caller)
template <bool GENERIC>
__device__ __forceinline__ float edge_logit(const float* s_dst_row, const float* s_src_row, float sd, int h, int H,
                                            float slope, int act) {
  if (GENERIC && act == B200GAT_LOGIT_HEAD_SOFTMAX) return head_softmax(s_dst_row, 1, s_src_row, h, H);
  return logit_act<GENERIC>(sd + __ldg(s_src_row + h), slope, act);
}

template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, G));
  return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
  return v;
}

}  // namespace b200gat
