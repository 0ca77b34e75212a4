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
  B200GAT_REQUIRE(L.logit_activation >= B200GAT_LOGIT_LEAKY_RELU && L.logit_activation <= B200GAT_LOGIT_TANH,
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
