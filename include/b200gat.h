/* b200gat.h — C ABI of the B200-native GAT layer hot path (libb200gat.so).
 *
 * Every entry point replaces a stretch of the reference's Python/PyG layer (there is no FFI in the reference;
 * the boundary a maintainer would bind is GraphAttentionLayer.forward and the autograd backward it implies):
 *
 *   b200gat_csr_build ..... GAT.py:38   add_self_loops + the (dst-)grouping PyG's softmax/scatter do implicitly
 *   b200gat_proj_fwd ...... GAT.py:42-52  per-head Linear x3 + stack/transpose (Wh, s_src, s_dst)
 *   b200gat_edge_fwd ...... GAT.py:53-67  propagate: gather, LeakyReLU logits, segment softmax, dropout mask,
 *                                         weighting, concat/mean, scatter-add, + bias
 *   b200gat_edge_bwd ...... autograd of GAT.py:53-67 (run_inductive.py:84), recomputing alpha
 *   b200gat_proj_bwd ...... autograd of GAT.py:42-52 (gX, gW)
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.  All data pointers are DEVICE pointers on the
 *     caller's current device; `stream` is a cudaStream_t passed as void*.
 *   - the caller owns every buffer, including workspaces; the library never allocates, frees or retains
 *     pointers past a call and never synchronises the device.
 *   - return 0 = ok; < 0 = argument error detected on the host before any launch (B200GAT_E_*);
 *     > 0 = cudaError_t of a failed launch.  b200gat_last_error() holds a message for the calling thread.
 *   - fp32 values, int32 graph indices (E + N < 2^31), int64 only for the incoming COO edge_index.
 *
 * Internal feature layout: a layer with H heads of C channels keeps per-node rows of width Dp = H * c_pad,
 * c_pad = round_up(C, 4), head-major; pad channels are zero (W / bw / a1 / a2 pad rows are zero).
 */
#ifndef B200GAT_H
#define B200GAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GAT_ABI_VERSION 15

enum {
  B200GAT_OK = 0,
  B200GAT_E_NULL = -1,       /* a required pointer is NULL */
  B200GAT_E_SHAPE = -2,      /* inconsistent / unsupported sizes */
  B200GAT_E_ALIGN = -3,      /* pointer or leading dimension not 16-byte aligned where required */
  B200GAT_E_WORKSPACE = -4,  /* workspace too small */
  B200GAT_E_UNSUPPORTED = -5
};

/* Activations between layers (GATNet.py:63-75 applies F.elu to every hidden layer's output) can be fused into the
 * layer boundary: the PRODUCING layer keeps its pre-activation output (out_activation in b200gat_edge_bwd: the upstream
 * gradient is d/d act(out) and is multiplied by act'(out) on the fly), the CONSUMING layer applies the activation
 * while it loads x (x_activation in b200gat_proj_fwd / b200gat_proj_bwd).  act(out) is never written to memory. */
enum { B200GAT_ACT_NONE = 0, B200GAT_ACT_ELU = 1 };

/* Destination-sorted CSR (+ source-sorted CSC) of [edge_index ; self loops].  All arrays int32 on device. */
typedef struct {
  int64_t num_nodes;       /* N */
  int64_t num_edges;       /* E' = E + N */
  const int32_t* rowptr;   /* [N+1]  CSR by destination */
  const int32_t* col;      /* [E']   source of each CSR entry */
  const int32_t* eid;      /* [E']   position of the entry in the ORIGINAL [edges ; loops] order */
  const int32_t* colptr;   /* [N+1]  CSC by source (stable w.r.t. CSR order) */
  const int32_t* crow;     /* [E']   destination of each CSC entry */
  const int32_t* ceid;     /* [E']   original position of each CSC entry */
  int64_t span;            /* max |source - destination| over the edges (status[1] of b200gat_csr_build), or < 0 if
                              unknown.  The edge kernels use it to pick their schedule: a small span (block-diagonal
                              graph batches) keeps the gathered rows L2-resident and favours occupancy; a large one
                              (one big graph) streams them from HBM and favours more gathers in flight per warp */
  /* Scheduling by degree.  The edge kernels walk a row with ONE lane group; a power-law hub (the 2.4 M-node graph has
   * rows of 40 k edges) would keep a single warp busy long after the rest of the grid has drained.  Rows whose degree
   * exceeds B200GAT_HUB_DEGREE are therefore listed here (b200gat_hub_rows) and processed by a second launch with one
   * whole CTA per (row, head); the row-per-group kernels read a row's end from rowend / colend, where those rows are
   * empty.  num_hub_* == 0 (then the four pointers may be NULL): no row is treated specially. */
  const int32_t* hub_rows; /* [num_hub_rows] destination rows with in-degree  > B200GAT_HUB_DEGREE (any order) */
  int64_t num_hub_rows;
  const int32_t* rowend;   /* [N] rowptr[i + 1], but rowptr[i] for the rows listed in hub_rows */
  const int32_t* hub_cols; /* [num_hub_cols] source rows      with out-degree > B200GAT_HUB_DEGREE (any order) */
  int64_t num_hub_cols;
  const int32_t* colend;   /* [N] colptr[j + 1], but colptr[j] for the rows listed in hub_cols */
  int64_t max_in_degree;   /* largest in- / out-degree (count[1] of b200gat_hub_rows); hub rows longer than */
  int64_t max_out_degree;  /* B200GAT_GIANT_DEGREE are cut into segments of that many edges, one CTA per segment */
} b200gat_graph;

#define B200GAT_MAX_PEERS 7   /* other GPUs of one NVSwitch node */
#define B200GAT_HUB_DEGREE 512
#define B200GAT_GIANT_DEGREE 4096

/* Layer geometry, GAT.py:8 (input_channels, output_channels, num_heads, concat). */
typedef struct {
  int64_t in_channels;     /* F_in */
  int64_t out_channels;    /* C    */
  int64_t heads;           /* H    */
  int64_t c_pad;           /* round_up(C, 4) */
  int32_t concat;          /* GAT.py:63-66 */
  float negative_slope;    /* GAT.py:30 (0.2) */
  int32_t logit_activation;   /* B200GAT_LOGIT_*: the function applied to a_i + a_j before the softmax (GAT.py:58) */
  int32_t reserved;
} b200gat_layer;

/* Logit activations.  GAT.py:30 fixes LeakyReLU(0.2); run_act_func_experiment.py:15,111 runs the same layer with
 * LogSigmoid, Tanh and nn.Softmax().  The last one is applied to the [E', H] logit tensor with no `dim`, i.e. ACROSS THE
 * HEADS of one edge (B200GAT_LOGIT_HEAD_SOFTMAX; uniform attention with one head): every (row, head) work item computes
 * all H logits of its edges in the forward; its backward couples the heads of an edge, so b200gat_edge_bwd runs it in two
 * passes through a per-edge scratch buffer (b200gat_edge_bwd_args.edge_scratch) — the only variant that has one. */
enum { B200GAT_LOGIT_LEAKY_RELU = 0, B200GAT_LOGIT_LOGSIGMOID = 1, B200GAT_LOGIT_TANH = 2, B200GAT_LOGIT_HEAD_SOFTMAX = 3 };

/* Attention dropout (GAT.py:61: F.dropout on the [E', H] coefficients, scaled 1/(1-p), NOT renormalised).  The keep-
 * multiplier of coefficient (edge e, head h) — e = position in the ORIGINAL [edges ; loops] order — has two sources:
 *   the `mask` tensor of the args structs, if given (parity tests: the reference's own masks), else
 *   p > 0: generated INSIDE the kernels — Philox4x32-10, key = seed[0], counter = (e, h / 4, seed[1]), word h % 4, keep
 *   iff word >= p * 2^32, multiplier 1/(1-p).  Forward (CSR order) and backward (CSC order) regenerate the identical value
 *   from (seed, e, h); no [E', H] tensor exists.  `seed` is a DEVICE pointer to two uint64 words read by the kernels, so a
 *   captured CUDA graph draws a fresh mask on every replay once the caller's RNG kernel (inside the graph) rewrites them.
 * b200gat_dropout_mask() writes the same multipliers as a tensor (tests: Philox mode == tensor mode, statistics). */
typedef struct { float p; float reserved; const uint64_t* seed; } b200gat_dropout;
int b200gat_dropout_mask(const b200gat_dropout* d, int64_t num_edges, int64_t heads, float* mask /* out [E', H] */, void* stream);

int b200gat_abi_version(void);
/* number of kernels this library has launched in the process so far (its own kernels; CUB sort passes excluded) */
uint64_t b200gat_launch_count(void);
/* copies the calling thread's last error message (NUL-terminated) into buf; returns its length */
int b200gat_last_error(char* buf, size_t buf_len);

/* ---- K0: graph ingestion (GAT.py:38) ------------------------------------------------------------------- */
size_t b200gat_csr_workspace_bytes(int64_t num_nodes, int64_t num_input_edges);
/* edge_index: int64 [2, E] contiguous (row 0 = source, row 1 = target).  status: device int32[2], set to
 * {number of out-of-range indices, max |source - destination| (b200gat_graph.span)}; the arrays are still well-formed (indices clamped) if non-zero. */
int b200gat_csr_build(const int64_t* edge_index, int64_t num_input_edges, int64_t num_nodes,
                      int32_t* rowptr, int32_t* col, int32_t* eid,
                      int32_t* colptr, int32_t* crow, int32_t* ceid,
                      int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Lists the rows of a CSR / CSC pointer array whose degree exceeds B200GAT_HUB_DEGREE: list[0 .. *count) (device, any
 * order, at most `cap` entries written; cap >= ptr[num_rows] / B200GAT_HUB_DEGREE + 1 always suffices), *count (device
 * int32[2]) = {their number, the largest degree}, ends[r] (device [num_rows]) = ptr[r + 1], or ptr[r] for a listed row.
 * The caller reads count back and passes it as num_hub_rows / max_in_degree (num_hub_cols / max_out_degree). */
int b200gat_hub_rows(const int32_t* ptr, int64_t num_rows, int32_t* list, int64_t cap, int32_t* count, int32_t* ends,
                     void* stream);

/* ---- K1: projection + attention logits (GAT.py:42-52) ----------------------------------------------------- */
typedef struct {
  b200gat_layer layer;
  int64_t num_nodes;
  const float* x;  int64_t ldx;   /* [N, F_in], row stride ldx */
  const float* w;                 /* [Dp, F_in] packed ws[h].weight, pad rows zero */
  const float* bw;                /* [Dp]       packed ws[h].bias */
  const float* a1; const float* a2;   /* [Dp]   attentions1/2[h].weight */
  const float* b1; const float* b2;   /* [H]    attentions1/2[h].bias */
  float* wh;                      /* out [N, Dp] */
  float* s_src; float* s_dst;     /* out [N, H]  (a1·Wh + b1 gathered at the source, a2·Wh + b2 at the target) */
  void* workspace; size_t workspace_bytes;
  void* x_split; size_t x_split_bytes;   /* optional out: the tensor-core operand splits of x AND of w (two fp16 planes +
                                            scale each), kept by the caller for b200gat_proj_bwd (w must not change in
                                            between); b200gat_proj_split_bytes() bytes, 256-byte aligned.  NULL: the splits
                                            live in workspace and are redone in the backward */
  int32_t x_activation;                  /* B200GAT_ACT_*: the projection consumes act(x) */
  const uint32_t* x_amax;                /* optional: device word holding the bit pattern of an upper bound of max|x|
                                            (b200gat_edge_fwd's out_amax of the producing layer); saves one pass over x */
  /* Projection fused with the all-gather of Wh (row-partitioned multi-GPU execution, GAT.py:42-52 on the own row block):
   * with num_peers > 0 the GEMM's store epilogue ALSO writes every Wh tile into the other GPUs' memory over NVLink
   * (peer-mapped pointers, e.g. torch symmetric memory), so the exchange overlaps the GEMM tile by tile instead of
   * following it as a collective.  wh_peers[k] = address, in THIS process, of the place of this rank's row block inside
   * peer k's full [P * block, Dp] buffer (same leading dimension as wh).  The caller synchronises the GPUs before anyone
   * reads the gathered buffer.  Tensor-core path only (b200gat_proj_split_bytes() != 0), else B200GAT_E_UNSUPPORTED. */
  float* wh_peers[B200GAT_MAX_PEERS];
  int32_t num_peers;
  /* Optional bf16 storage of the GATHERED rows (a separately toleranced mode, never the default): wh_bf16 [N, Dp] receives a
   * bf16 (round-to-nearest-even) copy of wh next to the fp32 one.  b200gat_edge_fwd gathers from it when given there, which
   * halves the bytes every edge pulls through L2 / HBM — and over NVLink when the rows are all-gathered (row-partitioned
   * mode).  The arithmetic (logits, softmax, accumulation) stays fp32. */
  void* wh_bf16;
} b200gat_proj_fwd_args;
size_t b200gat_proj_fwd_workspace_bytes(const b200gat_layer* layer, int64_t num_nodes);
/* bytes of an x_split buffer for this geometry; 0 when the shape runs on the CUDA-core path (pass x_split = NULL) */
size_t b200gat_proj_split_bytes(const b200gat_layer* layer, int64_t num_nodes);
int b200gat_proj_fwd(const b200gat_proj_fwd_args* a, void* stream);

/* ---- K2: fused edge forward (GAT.py:53-67) ---------------------------------------------------------------- */
typedef struct {
  b200gat_layer layer;
  b200gat_graph graph;
  const float* wh;                /* [N, Dp] */
  const float* s_src; const float* s_dst;   /* [N, H] */
  const float* bias;              /* [H*C] if concat else [C] */
  const float* mask;              /* optional [E', H] dropout keep-multiplier in ORIGINAL edge order (GAT.py:61) */
  float* out;  int64_t ldo;       /* [N, D_out] */
  float* rowmax; float* rowsum;   /* out [N, H]: softmax statistics kept for the recomputing backward */
  float* o_heads;                 /* out [N, Dp]; required iff !concat && H > 1 (per-head aggregate) */
  uint32_t* out_amax;             /* optional out: device word <- bit pattern of max|out| (for the next layer's x_amax) */
  b200gat_dropout dropout;        /* in-kernel attention dropout, used when mask == NULL and dropout.p > 0 */
  const void* wh_bf16;            /* optional [N, Dp] bf16 copy of wh (b200gat_proj_fwd_args.wh_bf16): gathered instead of wh.
                                     Offered without dropout and with LeakyReLU logits (else B200GAT_E_UNSUPPORTED) */
} b200gat_edge_fwd_args;
int b200gat_edge_fwd(const b200gat_edge_fwd_args* a, void* stream);

/* ---- K3: fused edge backward (recomputes alpha; no per-edge tensor) ------------------------------------------ */
typedef struct {
  b200gat_layer layer;
  b200gat_graph graph;
  const float* gout; int64_t ldgo;    /* [N, D_out] upstream gradient */
  const float* out;  int64_t ldo;     /* [N, D_out] forward output (used when concat || H == 1) */
  const float* o_heads;               /* [N, Dp]   (used when !concat && H > 1) */
  const float* bias;
  const float* wh; const float* s_src; const float* s_dst;
  const float* rowmax; const float* rowsum;
  const float* mask;                  /* optional, as in forward */
  const float* a1; const float* a2;   /* [Dp] */
  float* g_t;                         /* out [N, Dp]: total gradient w.r.t. Wh */
  float* g_bw; float* g_a1; float* g_a2;   /* out [Dp] */
  float* g_b1; float* g_b2;           /* out [H] */
  float* g_bias;                      /* out [D_out] */
  void* workspace; size_t workspace_bytes;
  int32_t out_activation;             /* B200GAT_ACT_*: gout is the gradient w.r.t. act(out); `out` is pre-activation.
                                         Needs a concat-like layer (concat || H == 1) */
  void* g_t_split; size_t g_t_split_bytes;   /* optional out: gT as the tensor-core operand split consumed by
                                         b200gat_proj_bwd (b200gat_edge_bwd_split_bytes() bytes, 256-byte aligned).
                                         When given, g_t is scratch (it holds the un-finished gWh on return) */
  b200gat_dropout dropout;            /* as in the forward (same p and seed words => the same mask) */
  float* edge_scratch; size_t edge_scratch_bytes;   /* B200GAT_LOGIT_HEAD_SOFTMAX only: E' * H * 4 bytes (d loss / d e per
                                         CSC entry and head, consumed by the second pass); NULL otherwise */
  int32_t gather_bf16;                /* the gatherable gradient rows G are written (by the prep pass, into the workspace) and
                                         gathered (by the CSC pass) as bf16: the backward's half of the bf16 mode above */
  const float* rowrec_in;             /* optional [N, H, 4]: the prep pass already ran in the CONSUMING layer's gX GEMM
                                         (b200gat_proj_bwd_args.fuse_prep): gout is then G itself (activation applied,
                                         directly gatherable), out_activation is ignored and g_bias is NOT written */
} b200gat_edge_bwd_args;
size_t b200gat_edge_bwd_workspace_bytes(const b200gat_layer* layer, int64_t num_nodes);
/* bytes of a g_t_split buffer; 0 when the projection backward of this geometry runs on the CUDA-core path */
size_t b200gat_edge_bwd_split_bytes(const b200gat_layer* layer, int64_t num_nodes);
int b200gat_edge_bwd(const b200gat_edge_bwd_args* a, void* stream);

/* ---- K3 staged: the three passes of b200gat_edge_bwd as separate calls, for destination-row partitioned multi-GPU
 * execution of ONE large graph (the caller all-gathers `rowrec` and the gatherable gradient rows between prep and
 * csc, and all-reduces g_s_dst between csc and finish).  Row counts are the caller's OWN block; `crow` holds GLOBAL
 * destination ids indexing rowrec / g / g_s_dst. ------------------------------------------------------------------- */
typedef struct {
  b200gat_layer layer;
  int64_t num_rows;                   /* own destination rows */
  const float* gout; int64_t ldgo;    /* [rows, D_out] */
  const float* out;  int64_t ldo;     /* [rows, D_out] (concat || H == 1) */
  const float* o_heads;               /* [rows, Dp]    (!concat && H > 1) */
  const float* bias;
  const float* s_dst; const float* rowmax; const float* rowsum;   /* [rows, H] */
  float* rowrec;                      /* out [rows, H, 4]: {s_dst, rowmax, 1/(rowsum+1e-16), Drow} */
  float* g_pad;                       /* out: padded / 1/H-scaled copy of G ([rows, Dp] if concat || H == 1 else
                                         [rows, c_pad]); NULL iff gout is directly gatherable (concat-like, C % 4 == 0) */
  float* g_bias;                      /* out [D_out] (this block's partial column sums) */
  int32_t out_activation;             /* as in b200gat_edge_bwd_args: gout is d/d act(out); needs g_pad (G = gout * act'(out)) */
  void* g_pad_bf16;                   /* alternative to g_pad: the same rows (same element layout) written as bf16, ALWAYS */
} b200gat_edge_bwd_prep_args;
int b200gat_edge_bwd_prep(const b200gat_edge_bwd_prep_args* a, void* stream);

typedef struct {
  b200gat_layer layer;
  int64_t num_rows;                   /* own SOURCE rows */
  const int32_t* colptr;              /* [rows+1] CSC of the edges whose source is in the own block */
  const int32_t* crow;                /* global destination ids */
  const int32_t* ceid;                /* original edge positions (mask lookup) or NULL */
  const float* wh; const float* s_src;       /* own rows: [rows, Dp], [rows, H] */
  const float* rowrec;                /* ALL nodes: [N, H, 4] */
  const float* mask;                  /* optional [E', H] in original edge order */
  const float* g; int64_t ldg; int64_t g_head_stride;   /* ALL nodes: G[i,h,c] = g[i*ldg + h*g_head_stride + c] */
  float* g_wh;                        /* out [rows, Dp] */
  float* g_s_src;                     /* out [rows, H] */
  float* g_s_dst;                     /* in/out ALL nodes [N, H]: zero-initialised by the caller, accumulated atomically */
  int64_t span;                       /* as b200gat_graph.span for the gathered rows `g` (< 0: unknown) */
  const int32_t* hub_cols; int64_t num_hub_cols;   /* own source rows with out-degree > B200GAT_HUB_DEGREE (or NULL / 0) */
  const int32_t* colend;              /* [rows] as b200gat_graph.colend (required when num_hub_cols > 0) */
  int64_t max_out_degree;             /* as b200gat_graph.max_out_degree for the own source rows */
  b200gat_dropout dropout;            /* as in the forward; ceid required (global original edge positions) */
  const void* g_bf16;                 /* alternative to g: the rows as bf16 (ldg / g_head_stride still count ELEMENTS) */
} b200gat_edge_bwd_csc_args;
int b200gat_edge_bwd_csc(const b200gat_edge_bwd_csc_args* a, void* stream);

typedef struct {
  b200gat_layer layer;
  int64_t num_rows;
  const float* wh;                    /* [rows, Dp] */
  const float* a1; const float* a2;   /* [Dp] */
  const float* g_s_src; const float* g_s_dst;   /* [rows, H] (g_s_dst already reduced over all ranks) */
  float* g_t;                         /* in: g_wh, out: gT  [rows, Dp] */
  float* g_bw; float* g_a1; float* g_a2;   /* out [Dp] (partial sums of this block) */
  float* g_b1; float* g_b2;           /* out [H] */
  void* g_t_split; size_t g_t_split_bytes;   /* optional out, as in b200gat_edge_bwd_args: gT as the tensor-core operand split
                                         consumed by b200gat_proj_bwd (g_t then keeps the un-finished gWh) */
} b200gat_edge_bwd_finish_args;
int b200gat_edge_bwd_finish(const b200gat_edge_bwd_finish_args* a, void* stream);

/* ---- K4: projection backward -------------------------------------------------------------------------------- */
typedef struct {
  b200gat_layer layer;
  int64_t num_nodes;
  const float* g_t;                   /* [N, Dp] */
  const float* x; int64_t ldx;        /* [N, F_in] */
  const float* w;                     /* [Dp, F_in] */
  float* g_x; int64_t ldgx;           /* out [N, F_in] or NULL (input does not require grad) */
  float* g_w;                         /* out [Dp, F_in] */
  void* workspace; size_t workspace_bytes;
  const void* x_split; size_t x_split_bytes;   /* optional in: the splits of x and w written by b200gat_proj_fwd */
  int32_t x_activation;               /* as in the forward (used when x_split is absent) */
  const void* g_t_split; size_t g_t_split_bytes;   /* optional in: gT as written by b200gat_edge_bwd (g_t may be NULL) */
  int32_t parts;                      /* 0 = both products; B200GAT_PROJ_BWD_GX / _GW = only that one.  The two GEMMs are
                                         independent: a caller may issue gW on a second stream so that it overlaps the
                                         previous layer's edge backward, which only waits for gX (each call needs its own
                                         workspace) */
  /* Fusion across the layer boundary in the backward: this layer's gX IS the upstream gradient of the PRODUCING layer k
   * (whose output is this layer's x), and the first pass of layer k's edge backward ("prep": G = gX * act'(out_k), the row
   * records {s_dst, rowmax, 1/rowsum, Drow = <G, out_k - bias_k>}, g_bias_k = column sums of G) only streams gX and out_k
   * once more.  With fuse_prep = layer k's prep arguments (layer, num_rows, out / ldo, bias, s_dst, rowmax, rowsum, rowrec,
   * g_bias, out_activation; gout / g_pad ignored) the gX GEMM does that pass in its epilogue: g_x then receives G, and
   * layer k's b200gat_edge_bwd is called with rowrec_in = that rowrec.  Offered when b200gat_proj_bwd_can_fuse_prep()
   * returns 1 (tensor-core path; producer concat-like with 8 / 16 / 32 / 64 / 128 / 256 channels per head). */
  const b200gat_edge_bwd_prep_args* fuse_prep;
} b200gat_proj_bwd_args;
int b200gat_proj_bwd_can_fuse_prep(const b200gat_layer* consumer, int64_t num_nodes, const b200gat_layer* producer);
enum { B200GAT_PROJ_BWD_GX = 1, B200GAT_PROJ_BWD_GW = 2 };
size_t b200gat_proj_bwd_workspace_bytes(const b200gat_layer* layer, int64_t num_nodes);
int b200gat_proj_bwd(const b200gat_proj_bwd_args* a, void* stream);

/* ---- graph-level readout head of GATNet (GATNet.py:72-75), fused ----------------------------------------------------
 *   pooled = scatter_mean(act(x), batch, dim=0)   hid = relu(lin1(pooled))   logp = log_softmax(lin2(hid), dim=1)
 * act = ELU when x_activation is set: x is then the PRE-activation output of the last GAT layer (layer-boundary fusion as
 * above) and the gradient handed back is the one w.r.t. act(x).  `batch` need not be sorted; graph ids outside
 * [0, num_graphs) are skipped and counted in *status (device int32, read lazily by the caller).  Empty graphs pool to 0. */
typedef struct {
  int64_t num_nodes, num_graphs;
  int64_t in_channels, hidden, classes;   /* lin1: [hidden, in_channels], lin2: [classes, hidden] */
} b200gat_readout_geom;

typedef struct {
  b200gat_readout_geom geom;
  const float* x; int64_t ldx;            /* [N, in_channels] */
  const int64_t* batch;                   /* [N] graph id of every node */
  const float* w1; const float* b1;       /* lin1.weight [hidden, in_channels], lin1.bias [hidden] */
  const float* w2; const float* b2;       /* lin2.weight [classes, hidden], lin2.bias [classes] */
  float* pooled;                          /* out [G, in_channels] per-graph means   (kept for the backward) */
  float* counts;                          /* out [G] nodes per graph                (kept for the backward) */
  float* hidden_out;                      /* out [G, hidden] relu(lin1(pooled))      (kept for the backward) */
  float* logp;                            /* out [G, classes] */
  int32_t* status;                        /* out: device word, number of nodes with a graph id outside [0, G) */
  int32_t x_activation;                   /* B200GAT_ACT_* */
} b200gat_readout_fwd_args;
int b200gat_readout_fwd(const b200gat_readout_fwd_args* a, void* stream);

typedef struct {
  b200gat_readout_geom geom;
  const int64_t* batch;
  const float* w1; const float* w2;
  const float* pooled; const float* counts; const float* hidden_out; const float* logp;   /* from the forward */
  const float* g_logp;                    /* [G, classes] upstream gradient */
  float* g_x; int64_t ldgx;               /* out [N, in_channels] gradient w.r.t. act(x), or NULL */
  float* g_w1; float* g_b1; float* g_w2; float* g_b2;   /* out, overwritten */
  void* workspace; size_t workspace_bytes;   /* G * (classes + hidden + in_channels) * 4 bytes */
} b200gat_readout_bwd_args;
int b200gat_readout_bwd(const b200gat_readout_bwd_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GAT_H */
