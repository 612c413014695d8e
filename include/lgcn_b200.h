/*
 * lgcn_b200.h — C ABI of the B200-native LightGCN hot path (liblgcn_b200.so).
 *
 * The reference (saamiya225/Graph-and-sequential-recommendation-systems, LightGCN_work/code) has no
 * C/FFI operator interface on this path: everything below replaces *library call sites* inside its
 * Python files.  Each entry point cites the reference lines it stands in for.  INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a BORROWED DEVICE pointer unless the name ends in `_host`;
 *   - nothing here allocates, frees or synchronises: scratch space is a caller-provided workspace
 *     whose size comes from the matching *_workspace_bytes() query (host-side arithmetic only);
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *   - return value: 0 = enqueued, non-zero = rejected (bad argument / launch error); the message is
 *     available from lgcn_last_error() on the calling thread;
 *   - dense tables are row-major float32 with row stride d; node ids are users 0..n_users-1 followed
 *     by items n_users..N-1 (code/model.py:209, code/dataloader.py:223-227);
 *   - adjacency is CSR with int32 indptr/indices and float32 values.
 */
#ifndef LGCN_B200_H
#define LGCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGCN_ABI_VERSION 1
#define LGCN_MAX_Z 8          /* max number of own-row addends in the SpMM epilogue (layers L <= 8) */
#define LGCN_MAX_TOPK 128     /* max k of the fused score/top-k kernel */
#define LGCN_MAX_PEERS 7      /* other GPUs of one NVSwitch box a K1 launch can store its rows into */

typedef void* lgcn_stream_t;  /* cudaStream_t */

int lgcn_abi_version(void);
const char* lgcn_last_error(void);
/* Device properties the host side sizes grids with: out[0]=SM count, out[1]=max dyn smem per block,
 * out[2]=compute capability major*10+minor.  Host-side query, no stream. */
int lgcn_device_info(int32_t* out_host);
/* Enable access from the current device to memory of `peer_device` (multi-GPU row partition: K1 stores its
 * finished rows straight into the peers' buffers over NVLink).  Idempotent; host-side, no stream. */
int lgcn_enable_peer_access(int32_t peer_device);
/* diagnostic for the peer-memory path: kernel store of `value` to n floats at ptr, synchronous; out_host int32[4] =
 * {pointer type, owning device, current device, canAccessPeer} */
int lgcn_debug_poke(float* ptr, float value, int32_t n, int32_t* out_host, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4  device CSR builder + degree normaliser
 * replaces  code/dataloader.py:133-142 (UserItemNet, users_D/items_D) and :223-234
 *           (dok/lil block assignment -> CSR, rowsum, power(-0.5), D.A.D)
 *
 * in : train_user/train_item  int64[E]   (code/dataloader.py:121-123 dtype)
 * out: indptr int32[N+1], indices int32[<=2E] (sorted within a row), vals float32[<=2E],
 *      deg float32[N] (weighted degree = row sum, duplicates counted, dataloader.py:133-136),
 *      dinv float32[N] (deg^-1/2, 0 for deg 0), nnz_out int64[1] (number of stored entries)
 * Duplicate (u,i) pairs are summed like scipy's csr_matrix((ones,(u,i))).
 * -------------------------------------------------------------------------------------------*/
size_t lgcn_csr_build_workspace_bytes(int64_t E, int32_t n_users, int32_t m_items);
int lgcn_csr_build(const int64_t* train_user, const int64_t* train_item, int64_t E,
                   int32_t n_users, int32_t m_items,
                   int32_t* indptr, int32_t* indices, float* vals, float* deg, float* dinv,
                   int64_t* nnz_out, int32_t* status_out /* int32[1]: !=0 -> id out of range */,
                   void* workspace, size_t workspace_bytes, lgcn_stream_t stream);

/* Row-block build for the multi-GPU row partition ("a CSR block per rank"): every rank assembles ONLY the
 * rows [row_begin,row_end) of the same normalised adjacency, from an edge stream it may see in chunks.
 *   1. lgcn_degree_accumulate over every chunk: deg_counts int32[N] += weighted degree of every node (zero it first);
 *      lgcn_degree_finalize -> deg/dinv float32[N] exactly as lgcn_csr_build computes them.
 *   2. lgcn_csr_rows_emit over every chunk: both directions of each edge whose row is owned are appended as packed keys
 *      inside `workspace` at *cursor_dev (uint64[1], zero it first); n_keys_cap = sum of deg_counts over the owned rows.
 *   3. lgcn_csr_rows_finish(n_keys = *cursor_dev read back by the host): sort, merge duplicates, write the block's
 *      indptr int32[rows+1] (LOCAL offsets), indices (GLOBAL column ids), vals; nnz_out int64[1].
 * The block equals rows [row_begin,row_end) of lgcn_csr_build's output bit for bit.
 * status_out: 1 = id out of range, 2 = more keys than n_keys_cap. */
int lgcn_degree_accumulate(const int64_t* train_user, const int64_t* train_item, int64_t E, int32_t n_users, int32_t m_items,
                           int32_t* deg_counts, int32_t* status_out, lgcn_stream_t stream);
int lgcn_degree_finalize(const int32_t* deg_counts, int32_t n_nodes, float* deg, float* dinv, lgcn_stream_t stream);
size_t lgcn_csr_rows_workspace_bytes(int64_t n_keys_cap, int32_t n_rows_local);
int lgcn_csr_rows_emit(const int64_t* train_user, const int64_t* train_item, int64_t E, int32_t n_users, int32_t m_items,
                       int32_t row_begin, int32_t row_end, int64_t n_keys_cap, uint64_t* cursor_dev, int32_t* status_out,
                       void* workspace, size_t workspace_bytes, lgcn_stream_t stream);
int lgcn_csr_rows_finish(int64_t n_keys, int64_t n_keys_cap, int32_t n_users, int32_t m_items, int32_t row_begin, int32_t row_end,
                         const float* dinv, int32_t* indptr, int32_t* indices, float* vals, int64_t* nnz_out,
                         void* workspace, size_t workspace_bytes, lgcn_stream_t stream);

/* Row-major sorted COO (torch coalesced layout, code/dataloader.py:183-190,244) -> int32 CSR.
 * rows/cols int64[nnz] sorted by (row,col); writes indptr int32[n_rows+1] and indices int32[nnz]. */
int lgcn_coo_to_csr(const int64_t* rows, const int64_t* cols, int64_t nnz, int32_t n_rows,
                    int32_t* indptr, int32_t* indices, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K1  CSR SpMM with fused epilogue
 * replaces  torch.sparse.mm(g, x) per layer + stack/mean   code/model.py:216-222
 *           and SparseAddmmBackward (loss.backward(), code/utils.py:61); A_hat is symmetric, so the
 *           same CSR serves the transpose product.
 *
 *   acc[i,:] = sum_{j in row i} vals[j] * X[indices[j],:]            i in [0,n_rows)
 *   g[i,:]   = alpha*acc[i,:] + beta * sum_{t<nz} Z_t[i,:]
 *   plain    : Y[i,:] = g[i,:]
 *   adam     : torch.optim.Adam update of P,M,V rows with gradient g (code/utils.py:51,62);
 *              Y may be NULL.
 * row_mask / col_mask (optional bitmaps over rows / columns, bit i of word i>>5): rows whose bit is 0 are
 * skipped (Y not written); columns whose bit is 0 are rows of X known to be all-zero and are never read.
 * Used by the training step to drop work nobody consumes; NULL = full product.
 * X is indexed by GLOBAL column id; Y, Z_t, P, M, V by LOCAL row i (callers pre-offset the
 * pointers for a row partition).  d in {16,32,64,128,256}.
 *
 * Scheduling plan (degree-binned load balancing).  lgcn_spmm_plan_* turn the rows into WORK ITEMS
 * {row, start, end, seg_ref}: a row with at most seg_len non-zeros is one item, a longer row is cut into
 * equal segments whose partial sums meet in plan->partials (the last-arriving segment adds them in part
 * order and runs the epilogue: deterministic, no float atomics).  Items are laid out by DESCENDING
 * length.  plan_host == NULL runs one item per row in natural order without segmentation.
 * -------------------------------------------------------------------------------------------*/
typedef struct {
    int32_t seg_len;          /* rows with more non-zeros than this are segmented                 */
    int32_t n_long;           /* number of segmented rows                                          */
    int32_t n_segs;           /* total number of segments                                          */
    int32_t n_items;          /* n_rows - n_long + n_segs                                          */
    int32_t d_max;            /* partials holds n_segs*d_max floats                                */
    int32_t pad;
    const int32_t* items;     /* int32[4*n_items]: row,start,end,seg_ref(-1 = whole row), 16-B aligned */
    const int32_t* seginfo;   /* int32[4*n_segs]: part,n_parts,slot_base,long_id                    */
    int32_t* counters;        /* int32[n_long], zero-initialised, self-resetting                   */
    float* partials;          /* float32[n_segs*d_max]                                             */
    float* acc;               /* column-slab blocking (lgcn_spmm_plan_*_slab): float32[n_rows*d_max] running sums carried from one
                                 slab's launch to the next; NULL for a whole-row plan                                      */
    const int32_t* hinted_indices; /* optional (whole-row plans, gathered table larger than L2): copy of `indices` with bit 31 set
                                 for HOT columns (lgcn_spmm_hint_indices); their rows are gathered with an L2 evict-last
                                 policy, all other gathers and the (col,val) stream with evict-first                         */
} lgcn_spmm_plan_t;

/* Fused SpMM + all-gather for the row partition: pointers to the PEER GPUs' copies of Y (and of P for
 * the Adam epilogue), mapped into this process (CUDA IPC) and pre-offset to this rank's row block like Y itself.  The
 * epilogue stores every finished row locally and into each peer over NVLink; the ranks then only need a barrier. */
typedef struct {
    int32_t n_peers;
    int32_t multicast;            /* != 0: n_peers == 1 and y[0] / p[0] are NVSwitch MULTICAST addresses (multimem) of the
                                     row block — one multimem.st per 16 bytes reaches every replica, this GPU's included */
    void* y[LGCN_MAX_PEERS];      /* NULL entries are skipped by the Adam variant when Y is not written */
    void* p[LGCN_MAX_PEERS];      /* Adam variant only */
} lgcn_spmm_peers_t;

/* Device-side rendezvous of the ranks of one NVSwitch box (replaces a 4-byte NCCL all-reduce per exchanged layer, so the
 * whole row-partitioned step is capturable in ONE CUDA graph and has no collective launch on its path).
 * flags_local uint32[world] lives on this GPU; peer_flags_host[p] is rank p's flags array mapped into this process
 * (symmetric memory / CUDA IPC; entry `rank` = flags_local).  The kernel (one CTA, one thread per rank) does
 * fence.sys, stores the new epoch into slot `rank` of every rank's array (st.release.sys) and spins (ld.acquire.sys)
 * until all `world` local slots have reached it: every store any rank issued BEFORE its barrier — K1's peer / multimem
 * row stores included — is visible to the kernels enqueued after it.  epoch_dev uint32[1] counts the barriers of this
 * rank (device-resident, so graph replays keep counting); all ranks must enqueue the same sequence of barriers.
 * A rank that waits longer than timeout_ms sets err_dev[0] = 1 + the rank it waited for and gives up (no hang). */
int lgcn_rank_barrier(uint32_t* flags_local, void* const* peer_flags_host, int32_t rank, int32_t world,
                      uint32_t* epoch_dev, int32_t* err_dev, int32_t timeout_ms, lgcn_stream_t stream);

/* Kernel copy of n_bytes (multiple of 16, both pointers 16-byte aligned) where one side may be MAPPED PINNED HOST memory
 * (with unified addressing every cudaHostAlloc'ed block is): at the two ends of the captured training step it replaces the
 * H2D memcpy of the batch and the D2H memcpy of the loss (code/Procedure.py:52-54, code/utils.py:64 on the reference side),
 * so one step through the public API is "fill the pinned staging block, replay one graph, wait".  dst_is_host != 0 adds a
 * system-scope fence after the stores. */
int lgcn_copy_words(void* dst, const void* src, int64_t n_bytes, int32_t dst_is_host, lgcn_stream_t stream);

/* hinted_out[j] = indices[j] | (col_weight[indices[j]] >= threshold ? 1 << 31 : 0); col_weight int32[n_cols] (e.g. degrees) */
int lgcn_spmm_hint_indices(const int32_t* indices, int64_t nnz, const int32_t* col_weight, int32_t threshold,
                           int32_t* hinted_out, lgcn_stream_t stream);

/* counts_out int32[4] = {n_long, n_segs, longest item, rows with an item} (device).  A row is cut into at most 2048 segments. */
int lgcn_spmm_plan_count(const int32_t* indptr, int32_t n_rows, int32_t seg_len,
                         int32_t* counts_out, lgcn_stream_t stream);
/* COLUMN-SLAB BLOCKING for graphs whose gathered table exceeds L2 (BASELINE config 5).  The *_slab variants plan only the
 * entries of each row whose column lies in [col_lo, col_hi) — contiguous, since rows are sorted by column.  One layer is
 * then one lgcn_spmm_f32 launch per slab IN ASCENDING COLUMN ORDER, all with the same plan.acc: a segment that is not its
 * row's first starts from acc[row], one that is not its row's last stores its running sum there, the last runs the epilogue
 * (empty rows get theirs from the slab planned with is_first_slab != 0).  X is read from HBM once per layer instead of once
 * per non-zero; the summation order of a row is unchanged.  n_items = counts[3] - counts[0] + counts[1]. */
int lgcn_spmm_plan_count_slab(const int32_t* indptr, const int32_t* indices, int32_t n_rows, int32_t seg_len,
                              int32_t col_lo, int32_t col_hi, int32_t is_first_slab, int32_t* counts_out, lgcn_stream_t stream);
int lgcn_spmm_plan_fill_slab(const int32_t* indptr, const int32_t* indices, int32_t n_rows, int32_t seg_len, int32_t max_len,
                             int32_t col_lo, int32_t col_hi, int32_t is_first_slab,
                             int32_t* items_out, int32_t* seginfo_out,
                             void* workspace, size_t workspace_bytes, lgcn_stream_t stream);
size_t lgcn_spmm_plan_workspace_bytes(int32_t max_len);
/* fills items_out int32[4*n_items] and seginfo_out int32[4*n_segs]; max_len = counts_out[2] */
int lgcn_spmm_plan_fill(const int32_t* indptr, int32_t n_rows, int32_t seg_len, int32_t max_len,
                        int32_t* items_out, int32_t* seginfo_out,
                        void* workspace, size_t workspace_bytes, lgcn_stream_t stream);

typedef struct {           /* device-resident Adam scalars, written by lgcn_adam_init / lgcn_adam_tick */
    float step_size;       /* (float)(lr / (1 - beta1^t)), computed in double like torch's Python scalars */
    float bc2_sqrt;        /* (float)sqrt(1 - beta2^t)                                                  */
    float beta1, beta2;    /* (float)beta                                                               */
    float w1, w2;          /* (float)(1 - beta) with the subtraction done in double                     */
    float eps;
    float pad0;
    int32_t step;          /* t */
    int32_t pad1;
    double lr_d, beta1_d, beta2_d;
} lgcn_adam_scalars_t;

int lgcn_spmm_f32(const int32_t* indptr, const int32_t* indices, const float* vals,
                  int32_t n_rows, int32_t d, const float* X, float* Y,
                  float alpha, float beta, const float* const* z_host, int32_t nz,
                  const lgcn_spmm_plan_t* plan_host, const uint32_t* row_mask, const uint32_t* col_mask,
                  const lgcn_spmm_peers_t* peers_host, lgcn_stream_t stream);

/* profiling hook: selects a tuning variant (unroll / CTA size / occupancy cap / L1 policy) of the d=64
 * plain kernel; 0 = shipped configuration.  Returns the previous value.  Results are identical. */
int lgcn_debug_spmm_variant(int variant);

/* measurement hook (bench.py `roofline_l2`, scripts/l2_gather_ceiling.py): the plainest kernel with K1's access pattern —
 * groups of lanes walk `run` consecutive entries of idx int32[n_idx] and sum the d-float rows of X they name (16-byte
 * gathers, several rows in flight); out float32[ceil(n_idx/run)*d].  variant 0..3 = {8,16} lanes x {4,8} rows in flight.
 * What it reaches on an L2-resident X is the L2->SM gather bandwidth K1 is bounded by on this box. */
int lgcn_debug_gather_rows(const float* X, const int32_t* idx, int64_t n_idx, int32_t d, int32_t run,
                           int32_t variant, float* out, lgcn_stream_t stream);

int lgcn_spmm_adam_f32(const int32_t* indptr, const int32_t* indices, const float* vals,
                       int32_t n_rows, int32_t d, const float* X, float* Y /* may be NULL */,
                       float alpha, float beta, const float* const* z_host, int32_t nz,
                       float* P, float* M, float* V, const lgcn_adam_scalars_t* scalars_dev,
                       const lgcn_spmm_plan_t* plan_host, const uint32_t* row_mask, const uint32_t* col_mask,
                       const lgcn_spmm_peers_t* peers_host,
                       int32_t clear_first_addend /* != 0: rows of Z_0 that are non-zero are zeroed after they are read — Z_0 = G and
                                                     this launch is its last reader in the step (replaces lgcn_bpr_clear_rows) */,
                       lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Adam (torch.optim.Adam defaults: betas (0.9,0.999), eps 1e-8, no weight decay, no amsgrad)
 * replaces  self.opt.step()  code/utils.py:51,62
 * lgcn_adam_tick: step += 1 and refresh the bias-correction scalars on the device (double
 * arithmetic, like torch's Python-side scalars).  lgcn_adam_f32: dense update of n floats.
 * -------------------------------------------------------------------------------------------*/
int lgcn_adam_init(lgcn_adam_scalars_t* scalars_dev, double lr, double beta1, double beta2, double eps,
                   int32_t step, lgcn_stream_t stream);
int lgcn_adam_tick(lgcn_adam_scalars_t* scalars_dev, lgcn_stream_t stream);
/* The head of a captured training step in one launch: lgcn_adam_tick, plus EITHER lgcn_batch_advance on advance_ctl_dev
 * (resident epoch; NULL otherwise) OR the pull of a host batch: stage_bytes (multiple of 16) from stage_src (mapped pinned
 * host memory) to stage_dst (device), as lgcn_copy_words does. */
int lgcn_step_begin(lgcn_adam_scalars_t* scalars_dev, int32_t* advance_ctl_dev, int32_t B_cap,
                    void* stage_dst, const void* stage_src, int64_t stage_bytes, lgcn_stream_t stream);
int lgcn_adam_f32(float* P, float* M, float* V, const float* G, int64_t n,
                  const lgcn_adam_scalars_t* scalars_dev, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2  fused BPR forward + closed-form gradient
 * replaces  getEmbedding gathers + bpr_loss arithmetic + their autograd backward
 *           code/model.py:125-134, :162-173 ; code/utils.py:55-61
 *
 * out  float32[N,d] propagated embeddings (users then items)
 * users/pos/neg int64[>= offset+B] (item ids are 0-based item indices, NOT offset by n_users)
 * batch_ctl_dev int32[4] (device) = {offset, B, total, B_global}: the batch is entries [offset, offset+B)
 *      of users/pos/neg, B <= B_cap.  Keeping it on the device lets ONE CUDA graph serve every batch
 *      of an epoch (lgcn_batch_advance moves the window: offset += B; B = min(B_cap, total-offset)).
 * inv_norm = 1/B_global (the means in code/model.py:170,173); pass <= 0 to take it from the batch
 *      descriptor instead: 1/ctl[3] when ctl[3] > 0 (global batch of a sharded step), else 1/B
 * loss_out float32[4]: {bpr = mean softplus(neg-pos), reg = 0.5*sum(|u|^2+|p|^2+|n|^2)/B,
 *                       total = bpr + decay*reg, running sum of total (host resets it)}
 * G float32[N,d]: += c_bpr*d(bpr)/d(out) + c_reg*d(reg)/d(out)   (scatter-add; G may be NULL for a
 *      forward-only call).  own_begin/own_end restrict the scatter to rows in [own_begin,own_end)
 *      (row partition); pass 0,N for all rows.
 * deterministic != 0: rows are reduced in a fixed order (no float atomics).
 * workspace must be zero-filled once after allocation (it holds a self-resetting arrival counter).
 * -------------------------------------------------------------------------------------------*/
size_t lgcn_bpr_workspace_bytes(int32_t B_cap, int32_t d);
int lgcn_bpr_fwd_bwd(const float* out, const int64_t* users, const int64_t* pos, const int64_t* neg,
                     int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items,
                     int32_t d, float inv_norm, float decay, float c_bpr, float c_reg,
                     float* loss_out, float* G, int32_t own_begin, int32_t own_end,
                     int32_t deterministic, void* workspace, size_t workspace_bytes,
                     uint32_t* clear_mask /* optional: the m0 bitmap of lgcn_batch_masks; the bits of this batch's rows are cleared
                                             (its last reader, the masked forward layer, ran before), so the next lgcn_batch_masks
                                             can skip its memset */,
                     float* loss_host_mapped /* optional: mapped pinned host float[4], receives a copy of loss_out (no D2H memcpy) */,
                     lgcn_stream_t stream);
/* Feature partition (dist_mode='featpart', SURVEY.md §8e for graphs that FIT one GPU): the d embedding columns are split
 * over the P ranks of one NVSwitch box, every rank holds the (N, d/P) column slice of every table and the whole CSR.  The
 * propagation then needs NO exchange at all (a CSR SpMM is independent per column) — the only cross-rank dependence of a
 * training step is the five dot products per triple, 40 KB per rank per step:
 *   lgcn_bpr_feat_partial  this rank's share of <u,p>, <u,n>, |u|^2, |p|^2, |n|^2 for every triple, stored into record
 *                          slot `part` of EVERY rank (records_peer[q] = rank q's record buffer mapped into this process)
 *   lgcn_rank_barrier
 *   lgcn_bpr_feat_finish   lgcn_bpr_fwd_bwd on the slice with the dot products replaced by the sum of the P records in rank
 *                          order (every rank computes the same loss bits); gradient rows of this rank's columns only
 * Record buffer: float32[2 * n_parts * B_cap * 8] per rank (16-byte aligned), double-buffered by the parity of
 * scalars_dev->step (the device-resident Adam step counter), so one barrier per step orders both the read-after-write and
 * the write-after-read hazard.  No counterpart in the reference (code/model.py:162-173 runs on one device). */
typedef struct {
    int32_t n_parts, part;
    const lgcn_adam_scalars_t* scalars_dev;
    float* records_local;
    float* records_peer[LGCN_MAX_PEERS + 1];     /* indexed by rank, own entry included; only the partial call stores to them */
} lgcn_bpr_feat_t;
int lgcn_bpr_feat_partial(const float* out_slice, const int64_t* users, const int64_t* pos, const int64_t* neg,
                          int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items, int32_t d_local,
                          const lgcn_bpr_feat_t* feat_host, void* workspace, size_t workspace_bytes, lgcn_stream_t stream);
int lgcn_bpr_feat_finish(const float* out_slice, const int64_t* users, const int64_t* pos, const int64_t* neg,
                         int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items,
                         int32_t d_local, float inv_norm, float decay, float c_bpr, float c_reg,
                         float* loss_out, float* G_slice, int32_t deterministic, const lgcn_bpr_feat_t* feat_host,
                         void* workspace, size_t workspace_bytes,
                         uint32_t* clear_mask, float* loss_host_mapped, lgcn_stream_t stream);
/* zero the rows of G a batch touched (cheaper than a full memset when B << N) */
int lgcn_bpr_clear_rows(float* G, const int64_t* users, const int64_t* pos, const int64_t* neg,
                        int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t d,
                        lgcn_stream_t stream);
int lgcn_batch_advance(int32_t* batch_ctl_dev, int32_t B_cap, lgcn_stream_t stream);
/* bitmaps over the N nodes for the current batch window: m0 = the 3B rows the loss reads, m1 = m0 plus their
 * neighbours in the CSR (m1 is cleared first, m0 if clear_first; m1 may be NULL).  uint32[(n_nodes+31)/32] each. */
int lgcn_batch_masks(const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t B_cap,
                     const int32_t* batch_ctl_dev, int32_t n_users, int32_t n_nodes,
                     const int32_t* indptr, const int32_t* indices, uint32_t* m0, uint32_t* m1,
                     int32_t clear_first /* 0: m0 is known to be all-zero (lgcn_bpr_fwd_bwd cleared it) */, lgcn_stream_t stream);

/* the same m0 restricted to the row block [row_begin,row_end) and indexed by LOCAL row: uint32[(row_end-row_begin+31)/32]
 * (row partition: every rank prunes the last forward layer to the batch rows it owns) */
int lgcn_batch_masks_rows(const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t B_cap,
                          const int32_t* batch_ctl_dev, int32_t n_users, int32_t row_begin, int32_t row_end,
                          uint32_t* m0_local, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Popularity gate
 * replaces  _fuse_item_embeddings  code/model.py:139-157 ; bpr_loss with use_pop_gate  code/model.py:162-183 and its backward
 *
 * params float32[lgcn_popgate_param_count(d, H1, H2)] = W1[H1] b1[H1] W2[d][H1] b2[d] G1[H2][2d] c1[H2] G2[H2] c2[1]
 *   (weights/biases of pop_mlp.0, pop_mlp.2, gate_mlp.0, gate_mlp.2 in nn.Linear layout); item_pop float32[m_items] is the
 *   standardised log1p(item degree) of code/model.py:74-77.  d in {32,64,128}, H1 <= 64, H2 <= 128.
 * lgcn_popgate_fuse: fused_out[m_items,d] = g*x + (1-g)*v for every item row x = out[n_users+i]; gate_out[m_items] optional.
 * lgcn_popgate_bpr_fwd_bwd: loss_out = {bpr - entropy_coeff*mean gate entropy over the 2B pos/neg gates, reg on (u, fused pos,
 *   fused neg), total = loss + decay*reg, running sum}; G[N,d] += d total/d out; params_grad += d total/d params
 *   (G == params_grad == NULL: forward only).  Float atomics: not bit-reproducible from run to run.
 *   workspace: zero-filled once (self-resetting arrival counter).
 * -------------------------------------------------------------------------------------------*/
int32_t lgcn_popgate_param_count(int32_t d, int32_t pop_hidden, int32_t gate_hidden);
int lgcn_popgate_fuse(const float* out, int32_t n_users, int32_t m_items, int32_t d, const float* item_pop,
                      const float* params, int32_t pop_hidden, int32_t gate_hidden, float temperature,
                      float* fused_out, float* gate_out, lgcn_stream_t stream);
size_t lgcn_popgate_bpr_workspace_bytes(int32_t B_cap);
int lgcn_popgate_bpr_fwd_bwd(const float* out, const int64_t* users, const int64_t* pos, const int64_t* neg,
                             int32_t B_cap, const int32_t* batch_ctl_dev, int32_t n_users, int32_t m_items, int32_t d,
                             const float* item_pop, const float* params, int32_t pop_hidden, int32_t gate_hidden,
                             float temperature, float entropy_coeff, float decay,
                             float* loss_out, float* G, float* params_grad,
                             void* workspace, size_t workspace_bytes, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K3  user x item scores fused with the train-item mask and per-row top-k
 * replaces  getUsersRating matmul (code/model.py:114-123), the -(1<<10) mask
 *           (code/Procedure.py:177-181) and torch.topk (code/Procedure.py:183)
 *
 * users_emb float32[n_users,d], items_emb float32[m_items,d]; users int64[Bt] (NULL = 0..Bt-1)
 * score(b,i) = fp32 FMA chain over k = 0..d-1 starting from 0 (exactly reproducible on a CPU)
 * masked cells (i in the CSR row of user b; stored column = mask_col_offset + i) score -1024
 * order: score descending, ties -> lowest item id first
 * idx_out int64[Bt,k], val_out float32[Bt,k]
 * -------------------------------------------------------------------------------------------*/
size_t lgcn_score_topk_workspace_bytes(int32_t Bt, int32_t m_items, int32_t k);
int lgcn_score_topk(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                    int32_t m_items, int32_t d, const int32_t* mask_indptr, const int32_t* mask_indices,
                    int32_t mask_col_offset, int32_t k, int64_t* idx_out, float* val_out,
                    void* workspace, size_t workspace_bytes, lgcn_stream_t stream);
/* Tensor-core path of the same operation (tcgen05 TF32 MMA + TMA + TMEM, d = 64, k <= 24): an approximate
 * two-pass filter (sampled tile maxima -> row threshold -> the ~50 items that reach it) picks candidates, the
 * candidates are rescored with the exact fp32 FMA chain and the top-k is certified against everything that was
 * filtered out.  16384 <= m_items < 2^27.  Rows that cannot be certified get flags_out[b] != 0
 * (1 certificate failed, 2 fewer than k candidates, 3 candidate list overflowed; count in n_flagged_out, device
 * int32[1]) and MUST be recomputed with lgcn_score_topk; all other rows are bit-identical to it.
 * workspace must be 1024-byte aligned.
 * THE MASK IS GIVEN IN POSITION SPACE: the kernel scores the items in an interleaved order (so that a block of 64
 * consecutive positions is a spread-out sample of the ids).  mask_indices[j] - mask_col_offset must be the POSITION
 * of the train item, ascending within each row; lgcn_score_topk_tc_item_positions maps item ids to positions
 * (position space = [0, lgcn_score_topk_tc_position_space(m_items))), and lgcn_csr_build over (user, position)
 * pairs yields exactly such a CSR (user rows, col offset n_users).  idx_out holds item ids as usual. */
int lgcn_score_topk_tc_supported(int32_t d, int32_t k);
size_t lgcn_score_topk_tc_workspace_bytes(int32_t Bt, int32_t m_items, int32_t k);
int32_t lgcn_score_topk_tc_position_space(int32_t m_items);
/* diagnostics: {offset of tau, offset of the event counts, lists per row, item tiles, tiles per split, offset of the
 * sample maxima, their row stride, events per list} of the workspace layout for this (Bt, m_items) */
int lgcn_score_topk_tc_debug_layout(int32_t Bt, int32_t m_items, int64_t* out8_host);
/* the same mapping on the host (no device needed): position of an item; item at a position, -1 for a hole */
int32_t lgcn_score_topk_tc_host_position(int32_t item, int32_t m_items);
int32_t lgcn_score_topk_tc_host_item(int32_t pos, int32_t m_items);
int lgcn_score_topk_tc_item_positions(const int64_t* items, int64_t n, int32_t m_items, int64_t* pos_out, lgcn_stream_t stream);
int lgcn_score_topk_tc(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                       int32_t m_items, int32_t d, const int32_t* mask_indptr, const int32_t* mask_indices,
                       int32_t mask_col_offset, int32_t k, int64_t* idx_out, float* val_out,
                       int32_t* flags_out, int32_t* n_flagged_out,
                       void* workspace, size_t workspace_bytes, lgcn_stream_t stream);
/* dense scores float32[Bt,m_items] for the unfused API (getUsersRating returns the matrix) */
int lgcn_score_dense(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                     int32_t m_items, int32_t d, float* scores, lgcn_stream_t stream);

/* The same matrix on tensor cores (d = 64): 3xTF32 — every operand is split exactly into hi + lo and the product is
 * lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator, so the scores carry fp32-class accuracy (|err| <~ 1e-6 |u||v|; not the
 * bit-exact FMA chain of lgcn_score_dense).  workspace must be 1024-byte aligned. */
int lgcn_score_dense_tc_supported(int32_t d);
size_t lgcn_score_dense_tc_workspace_bytes(int32_t Bt, int32_t m_items);
int lgcn_score_dense_tc(const float* users_emb, const float* items_emb, const int64_t* users, int32_t Bt,
                        int32_t m_items, int32_t d, float* scores, void* workspace, size_t workspace_bytes, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * On-device ranking metrics (precision/recall/NDCG sums over users)
 * replaces  test_one_batch/getLabel/RecallPrecision_ATk/NDCGatK_r
 *           code/Procedure.py:89-121 ; code/utils.py:173-200,212-217
 * topk_idx int64[Bt,k_max]; test CSR over the same Bt rows (sorted item ids); ks int32[nk];
 * sums_out float64[3*nk] += {precision, recall, ndcg} summed over rows, in a FIXED order (per-row values go through the
 * workspace, then one fixed-shape reduction per column): the result has the same bits on every call, which is what lets
 * the multi-GPU Test (ranked lists gathered, metrics computed on every rank) equal the single-GPU one bit for bit.
 * -------------------------------------------------------------------------------------------*/
size_t lgcn_rank_metrics_workspace_bytes(int32_t Bt, int32_t nk);
int lgcn_rank_metrics(const int64_t* topk_idx, int32_t Bt, int32_t k_max,
                      const int32_t* test_indptr, const int32_t* test_indices,
                      const int32_t* ks, int32_t nk, double* sums_out,
                      void* workspace, size_t workspace_bytes, lgcn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Host-side negative sampler with the semantics of code/sources/sampling.cpp:27-56 (per user
 * train_num/user_num triples, uniform positive, rejection-sampled negative, glibc rand()).
 * All pointers are HOST pointers.  allpos CSR: indptr int64[user_num+1], items int32.
 * out int32[user_num*(train_num/user_num)*(2+neg_num)].  Returns rows written or <0.
 * -------------------------------------------------------------------------------------------*/
void lgcn_sampler_seed(uint32_t seed);
/* The generator behind the sampler is glibc's rand() restated inside the library (same stream as srand/rand on glibc, on any
 * platform); its state (int32[33]) can be saved and restored, which is what lets an epoch's sample be drawn ahead of time and
 * rewound if the next call turns out to be something else. */
void lgcn_sampler_get_state(int32_t* state33_host);
int lgcn_sampler_set_state(const int32_t* state33_host);
int64_t lgcn_sample_negative(int32_t user_num, int32_t item_num, int64_t train_num,
                             const int64_t* allpos_indptr_host, const int32_t* allpos_items_host,
                             int32_t neg_num, int32_t* out_host);

/* sample_negative_ByUser (code/sources/sampling.cpp:58-86): one row {user, pos, neg...} per entry of users_host int32[n_listed],
 * and randint (sampling.cpp:22-25,100): both continue the same rand() stream as lgcn_sample_negative. */
int64_t lgcn_sample_negative_by_user(const int32_t* users_host, int64_t n_listed, int32_t user_num, int32_t item_num,
                                     const int64_t* allpos_indptr_host, const int32_t* allpos_items_host,
                                     int32_t neg_num, int32_t* out_host);
int32_t lgcn_randint(int32_t end);

/* ---------------------------------------------------------------------------------------------
 * Ingest of the reference's interaction files
 * replaces  the per-line Python parsing in Loader.__init__  code/dataloader.py:82-115
 * "uid item item ..." per line, whitespace separated; blank lines and lines without items are skipped; every
 * (uid, item) pair is appended in file order.  HOST pointers.  Returns the number of pairs in the file (< 0 on
 * error); at most `capacity` pairs are stored (call with capacity 0 / null outputs to size the arrays).
 * max_user/max_item: largest ids among the lines that have items (-1 for an empty file).
 * -------------------------------------------------------------------------------------------*/
int64_t lgcn_parse_interactions(const char* path, int64_t* users_out_host, int64_t* items_out_host, int64_t capacity,
                                int64_t* max_user_out_host, int64_t* max_item_out_host);

/* ---------------------------------------------------------------------------------------------
 * Device-side BPR sampler + shuffle
 * replaces  utils.UniformSample_original + utils.shuffle  code/utils.py:68-81,142-151 ; code/Procedure.py:50-55
 * indptr/indices: the adjacency CSR (user rows hold n_users + item).  Every user gets train_num/n_users triples
 * (uniform positive of its row, rejection-sampled negative outside it); output is already permuted, int64,
 * users_out/pos_out/neg_out[n] with n = (train_num/n_users)*n_users.  Counter-based RNG keyed by (seed, epoch).
 * status_out (device int32[1], zero it first; may be NULL): bit 0 = some user has no train item, bit 1 = some user has no
 * admissible negative — the triples of such users are fabricated and the epoch must not be trained on (the host sampler
 * returns an error for the same input, the reference's C++ divides by zero, code/sources/sampling.cpp:43).
 * -------------------------------------------------------------------------------------------*/
int lgcn_sample_bpr(const int32_t* indptr, const int32_t* indices, int32_t n_users, int32_t m_items,
                    int64_t train_num, uint64_t seed, uint64_t epoch,
                    int64_t* users_out, int64_t* pos_out, int64_t* neg_out, int32_t* status_out, lgcn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LGCN_B200_H */
