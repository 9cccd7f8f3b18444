/*
 * gsage_b200.h -- C ABI of the B200-native GraphSAGE minibatch hot path.
 *
 * The reference (Lolash/graphSAGE-pytorch) has no FFI layer: its boundary is the Python
 * class surface of src/models.py.  This header is what a maintainer would bind from those
 * classes (ctypes stub in INTEGRATION.md); every entry point names the reference lines it
 * replaces.  Conventions:
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - the caller owns every buffer (inputs, outputs, scratch); nothing is allocated,
 *     nothing synchronises with the host, every launch goes to `stream` (a cudaStream_t
 *     passed as void*), so a sequence of calls can be captured into a CUDA graph.
 *   - row counts that are only known on the device (|U| after unique) are passed as
 *     `const int32_t* num_rows_dev` (nullable => `max_rows` rows are live); kernels are
 *     launched for `max_rows` and rows >= *num_rows_dev exit.
 *   - node ids and row indices are int32 (N < 2^31), CSR offsets int64, -1 pads lists.
 *   - return value: 0 on success, a negative GS_ERR_* for bad arguments, or a positive
 *     cudaError_t from the launch.  gs_error_string() decodes both.
 */
#ifndef GSAGE_B200_H_
#define GSAGE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 16

#define GS_OK 0
#define GS_ERR_BAD_ARG (-1)
#define GS_ERR_UNSUPPORTED (-2)
#define GS_ERR_WORKSPACE (-3)
#define GS_ERR_ALIGNMENT (-4)

#define GS_AGG_MEAN 0 /* src/models.py:311-314 */
#define GS_AGG_MAX 1  /* src/models.py:316-326 */

#define GS_SELF_KEEP 0 /* leave sampled rows as drawn                                    */
#define GS_SELF_DROP 1 /* gcn=False: remove the row's own id (src/models.py:298)         */
#define GS_SELF_ONCE 2 /* gcn=True : own id present exactly once (src/models.py:285)     */

#define GS_PREC_FP32 0      /* SIMT FFMA, fp32 accumulate: the 1e-5 parity mode          */
#define GS_PREC_TF32 1      /* tcgen05 kind::tf32, one pass                              */
#define GS_PREC_TF32X3 2    /* tcgen05 kind::tf32, 3-term split (fp32-faithful)          */

#define GS_MAX_FANOUT 32

#define GS_UNIQUE_MARKED 1      /* gs_unique_remap_bitmap_ex: the ids are already marked (by gs_sample_neighbors_ex)   */
#define GS_UNIQUE_LEAVE_MARKS 2 /* ... and the clear pass is left to the next gs_sample_neighbors_ex (clear_bitmap)    */

typedef void* gs_stream_t; /* cudaStream_t */

int gs_version(void);
const char* gs_error_string(int code);
/* number of kernels this library has launched since load / since the last reset (bench.py's gpu_launches) */
int64_t gs_launch_count(void);
void gs_launch_count_reset(void);
/* Diagnostics: enqueue a one-thread kernel that writes the GPU's %globaltimer (ns) to *slot (device memory).
 * Markers placed behind the launches of a step give per-kernel completion times on every branch of the
 * step graph (native.timeline_begin / timeline_read); not counted by gs_launch_count. */
int gs_debug_stamp(uint64_t* slot, gs_stream_t stream);
/* Programmatic dependent launch (every kernel of the library starts with griddepcontrol.launch_dependents
 * + griddepcontrol.wait and is launched with the programmatic-stream-serialization attribute, so the next
 * kernel of a chain is scheduled while the current one drains).  mode: 1 = on, 0 = off, -1 = follow the
 * GS_PDL environment variable (default on).  The attribute is fixed when a launch is captured into a graph;
 * a caller running two chains side by side switches it off for one of them (early-resident dependents of
 * one chain would otherwise hold SM slots the other chain needs). */
void gs_set_pdl(int32_t mode);

/* ------------------------------------------------------------------------------------
 * K1  neighbour sampler.  Replaces GraphSage._get_unique_neighs_list's sampling half,
 * src/models.py:279-285: rows with degree < k keep every neighbour, others draw k distinct
 * uniform ones (Floyd's subset algorithm on Philox4x32-10 keyed by (seed, offset, row)).
 * Rows are written sorted ascending, -1 padded to `stride` (>= k+1 when self_mode is ONCE).
 * offset_dev (nullable): a device counter added to `offset`, so a captured CUDA graph draws
 * fresh samples on every replay (the host bumps / a kernel increments the counter).
 * ------------------------------------------------------------------------------------ */
int gs_sample_neighbors(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                        const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                        int32_t k, int32_t stride, int32_t self_mode,
                        uint64_t seed, uint64_t offset, const int64_t* offset_dev,
                        int32_t* out_nbr, int32_t* out_cnt, gs_stream_t stream);

/* The same with work of the neighbouring kernels of a preparation chain folded in (all four nullable):
 *   queue_desc / fetch_dst  gs_fetch_batch fused: the rows are the next batch of the queue ({address, rows, next,
 *                           ticket}: 4 int64, ticket zero), max_rows == b_sz, num_rows_dev and nodes NULL; row r's node is
 *                           also written to fetch_dst[r] (the batch's id list, e.g. the label index of the loss)
 *   mark_bitmap             the "mark" pass of the following gs_unique_remap_bitmap_ex (flag GS_UNIQUE_MARKED): every id
 *                           of a live row (its node and the drawn neighbours) sets its bit
 *   clear_bitmap            the "clear" pass of the PRECEDING gs_unique_remap_bitmap_ex (flag GS_UNIQUE_LEAVE_MARKS): the
 *                           rows of this call are exactly the ids it emitted; each zeroes the word of its own node.
 *                           Not together with mark_bitmap (they would race on a word).
 *   prefetch_table          (layer 1) the table the next kernel gathers from: row id of every drawn neighbour and of the
 *                           node itself is requested into L2 (cp.async.bulk.prefetch.L2, prefetch_row_bytes bytes at
 *                           prefetch_table + id * prefetch_ld_bytes; both multiples of 16), so the DRAM fetch of the
 *                           gathered rows starts here and overlaps the launch gap / ramp of gs_agg_fwd. */
int gs_sample_neighbors_ex(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                           const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                           int32_t k, int32_t stride, int32_t self_mode,
                           uint64_t seed, uint64_t offset, const int64_t* offset_dev,
                           int32_t* out_nbr, int32_t* out_cnt,
                           int64_t* queue_desc, int32_t* fetch_dst, uint32_t* mark_bitmap, uint32_t* clear_bitmap,
                           const void* prefetch_table, int64_t prefetch_ld_bytes, int32_t prefetch_row_bytes,
                           gs_stream_t stream);

/* Batch queue of the device-resident train loop (src/utils.py:141-145: every batch of an epoch
 * is a slice of one shuffled id array).  queue_desc is 4 DEVICE int64: {address of an int32
 * [rows x b_sz] array, rows, next, ticket (zero; used by gs_sample_neighbors_ex)}; copies row next % rows into dst and increments next -- a
 * kernel of the step's graph, so back-to-back replays need no host-side copy between them. */
int gs_fetch_batch(int64_t* queue_desc, int32_t b_sz, int32_t* dst, gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K2  unique + remap.  Replaces src/models.py:286-288 (set.union / dict(zip)) and the
 * index lookups of :306 and :274 (_nodes_map).  Input ids = nodes[r] and nbr[r][j] (>= 0).
 * Output: uniq ascending, *num_uniq_dev, nbr_idx[r][j] = position of nbr[r][j] in uniq
 * (-1 stays -1), self_idx[r] = position of nodes[r].  LSD radix sort (8-bit digits over
 * `id_bits` bits) + run-length heads + binary-search remap; one CTA when it fits in
 * shared memory, multi-CTA otherwise.  Either of nbr_idx / self_idx may be NULL.
 * ------------------------------------------------------------------------------------ */
size_t gs_unique_workspace_bytes(int32_t max_rows, int32_t stride);
int gs_unique_remap(const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                    const int32_t* nbr, int32_t stride, int32_t id_bits,
                    int32_t* uniq, int32_t* num_uniq_dev, int32_t* nbr_idx, int32_t* self_idx,
                    void* workspace, size_t workspace_bytes, gs_stream_t stream);

/* K2, bitmap path: same outputs as gs_unique_remap (ascending unique ids, bit for bit) in
 * O(M + N/32) memory-parallel work: mark ids in a num_nodes-bit map, popcount-scan it, emit
 * set bits at their rank, remap by rank lookup, clear the touched words.  The first
 * ceil(num_nodes/32) words (rounded up to 4096) of `workspace` are the bitmap: the caller
 * zeroes the workspace ONCE; every call leaves the bitmap zeroed again. */
size_t gs_unique_bitmap_workspace_bytes(int64_t num_nodes);
int gs_unique_remap_bitmap(const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                           const int32_t* nbr, int32_t stride, int64_t num_nodes,
                           int32_t* uniq, int32_t* num_uniq_dev, int32_t* nbr_idx, int32_t* self_idx,
                           void* workspace, size_t workspace_bytes, gs_stream_t stream);
/* flags: GS_UNIQUE_MARKED | GS_UNIQUE_LEAVE_MARKS -- the mark / clear passes run inside the neighbouring sampler
 * launches (gs_sample_neighbors_ex); with both set the call is two launches (scan, emit + remap). */
int gs_unique_remap_bitmap_ex(const int32_t* nodes, const int32_t* num_rows_dev, int32_t max_rows,
                              const int32_t* nbr, int32_t stride, int64_t num_nodes,
                              int32_t* uniq, int32_t* num_uniq_dev, int32_t* nbr_idx, int32_t* self_idx,
                              void* workspace, size_t workspace_bytes, int32_t flags, gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K3  gather-segment-reduce.  Replaces GraphSage.aggregate, src/models.py:300-326 (the
 * embed_matrix gather, the dense mask, its normalisation and mask.mm / the MAX loop).
 * out[r,:] = mean or max over j < cnt[r] of table[nbr[r*stride+j], :].  cnt==0 gives NaN
 * for MEAN (0/0 as in the reference) and NaN for MAX (the reference raises).
 * `table` rows are `ld` floats apart; ld % 4 == 0 and 16-byte aligned bases required.
 * argmax (MAX only, nullable) receives the winning table row per element, first on ties.
 * ------------------------------------------------------------------------------------ */
/* Occupancy of the K3 forward grid for subsequent launches (process-wide, like gs_set_pdl): at most
 * `ctas_per_sm` persistent CTAs of 8 warps per SM; 0 = full occupancy (default).  Lower it for launches
 * that run beside a critical chain of larger CTAs (trainer.PipelinedTrainer), restore it afterwards. */
void gs_set_agg_ctas(int32_t ctas_per_sm);
/* Process-wide switch like gs_set_pdl: launches issued while it is on are background work of a two-branch
 * step (trainer.PipelinedTrainer's preparation branch).  Every kernel of the library prefers the maximum
 * shared-memory carveout so that CTAs of both branches can share an SM; the K3 forward kernel, whose HBM
 * throughput needs the L1 that carveout removes, does so only while this switch is on. */
void gs_set_background(int32_t on);
/* Process-wide switch like gs_set_pdl: launches issued while it is on may read their INDEX inputs (neighbour / self
 * index lists, labels and label_index, num_rows_dev) before they wait for the previous kernels of the stream, so that
 * those loads run under the tail of the predecessor (programmatic dependent launch lets a whole chain of launches be
 * resident and waiting).  Only correct when those inputs were produced before the chain began: by another stream joined
 * with an event (trainer.PipelinedTrainer's preparation branch), by a host-synchronised copy, or by a launch without
 * the attribute.  Off by default: every kernel waits first.  Honoured by gs_sage_top_sup and by the TMA-fed
 * gs_sage_gemm_fwd_ex. */
void gs_set_early_reads(int32_t on);
int gs_agg_fwd(const float* table, int64_t ld, int32_t dim,
               const int32_t* nbr, int32_t stride, const int32_t* cnt,
               const int32_t* num_rows_dev, int32_t max_rows, int32_t mode,
               float* out, int64_t ld_out, int32_t* argmax, int64_t ld_arg, gs_stream_t stream);

/* Autograd of K3 plus the self-row gather of src/models.py:265: scatter-adds
 *   grad_table[nbr[r][j], :] += grad_agg[r, :] / cnt[r]          (MEAN)
 *   grad_table[argmax[r][c], c] += grad_agg[r, c]                (MAX)
 *   grad_table[self_idx[r], :] += grad_self[r, :]                (when grad_self != NULL)
 * grad_table must be zeroed by the caller.  mask_table (nullable, [table rows x ld_mask]) is the
 * ReLU output the table rows were produced with (src/models.py:219): contributions to elements
 * whose output was <= 0 are dropped, i.e. grad_table receives d(pre-activation) directly. */
int gs_agg_bwd(const float* grad_agg, int64_t ld_ga, const float* grad_self, int64_t ld_gs, int32_t dim,
               const int32_t* nbr, int32_t stride, const int32_t* cnt, const int32_t* self_idx,
               const int32_t* argmax, int64_t ld_arg, const int32_t* num_rows_dev, int32_t max_rows, int32_t mode,
               float* grad_table, int64_t ld_gt, const float* mask_table, int64_t ld_mask, gs_stream_t stream);

/* K3 forward over a ROW-PARTITIONED bf16 feature table (BASELINE.json configs[4]: features
 * split by contiguous node-id blocks across the GPUs of one box).  shard_bases_host is a HOST
 * array of `num_shards` (<= 8) device pointers: shard s holds rows [s*rows_per_shard,
 * (s+1)*rows_per_shard) as bf16, `ld` elements apart (ld % 8 == 0); a base may be local HBM or
 * a peer mapping obtained with gs_peer_open -- gathered rows are then read straight over
 * NVLink, no collective.  nbr holds GLOBAL node ids; mode GS_AGG_MEAN (src/models.py:311-314, fp32
 * accumulate) or GS_AGG_MAX (:316-326; no argmax, the raw features take no gradient).  out_self (nullable) receives the fp32 copy of row self_nodes[r]: the
 * self_feats gather of src/models.py:265 for layer 1, which K4 then reads with self_idx = NULL. */
int gs_agg_fwd_bf16_sharded(const void* const* shard_bases_host, int32_t num_shards, int64_t rows_per_shard,
                            int64_t ld, int32_t dim, const int32_t* nbr, int32_t stride, const int32_t* cnt,
                            const int32_t* self_nodes, const int32_t* num_rows_dev, int32_t max_rows,
                            float* out_agg, int64_t ld_agg, float* out_self, int64_t ld_self, int32_t mode,
                            gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K4  SageLayer.  Replaces src/models.py:215-219: out = relu(W . [self | agg]^T)^T with the
 * concat never materialised: X[r,:] = [ self_table[self_idx[r], :dim] | agg[r, :dim] ]
 * (gcn: X = agg only, W is [out_dim x dim]).  self_idx NULL => identity.
 * ------------------------------------------------------------------------------------ */
int gs_sage_gemm_fwd(const float* self_table, int64_t ld_self, const int32_t* self_idx,
                     const float* agg, int64_t ld_agg, int32_t dim,
                     const float* weight, int64_t ldw, int32_t out_dim, int32_t gcn,
                     const int32_t* num_rows_dev, int32_t max_rows,
                     float* out, int64_t ld_out, int32_t relu, int32_t precision, gs_stream_t stream);

/* K3 writing the SageLayer's whole input row (the operands torch.cat joins at src/models.py:217), for layers whose
 * table is the raw feature table:  out_x[r, 0:dim] = table[self_nodes[r], :]  (src/models.py:265; skipped when self_nodes
 * is NULL),  out_x[r, agg_off : agg_off+dim] = mean / max over the row's list (:300-326); out_x_lo (nullable, same layout)
 * receives x - trunc_tf32(x) of every element written: the low half of K4's 3-term tf32 split.  The layer's GEMMs then
 * read dense operands (gs_sage_gemm_fwd_ex with self_idx NULL takes its all-TMA kernel). */
int gs_agg_fwd_x(const float* table, int64_t ld, int32_t dim, const int32_t* nbr, int32_t stride, const int32_t* cnt,
                 const int32_t* self_nodes, const int32_t* num_rows_dev, int32_t max_rows, int32_t mode,
                 float* out_x, int64_t ld_x, int32_t agg_off, float* out_x_lo, gs_stream_t stream);

/* The same with one more output: zero_out (nullable, [max_rows x ld_zero], same shape as out) is zero-filled for the
 * live rows and columns.  It is the buffer the backward scatter of the layer above accumulates d(out) into
 * (gs_agg_bwd / gs_sage_top_sup): the fill rides on the forward epilogue instead of being a launch of its own. */
/* x_lo / weight_lo (nullable): the low halves x - trunc_tf32(x) of the input rows (laid out like self_table .. agg,
 * which must then be one dense [max_rows x 2*dim] matrix: self_idx NULL, agg == self_table + dim, ld_self == ld_agg) and
 * of the weight.  With both given (or precision GS_PREC_TF32) the call runs the all-TMA kernel: no gather, no in-kernel
 * split.
 * l2_normalize != 0 (forward only; relu != 0, out_dim == 128, tensor-core precisions): the optional epilogue the original
 * GraphSAGE applies and this reference does not (src/models.py:219 ends at the ReLU): every output row is divided by
 * max(||row||_2, 1e-12).  Off everywhere the reference's numbers are reproduced. */
int gs_sage_gemm_fwd_ex(const float* self_table, int64_t ld_self, const int32_t* self_idx,
                        const float* agg, int64_t ld_agg, int32_t dim,
                        const float* weight, int64_t ldw, int32_t out_dim, int32_t gcn,
                        const int32_t* num_rows_dev, int32_t max_rows,
                        float* out, int64_t ld_out, int32_t relu, int32_t precision,
                        float* zero_out, int64_t ld_zero, const float* x_lo, const float* weight_lo,
                        int32_t l2_normalize, gs_stream_t stream);

/* dW[h,k] += sum_r dZ[r,h] X[r,k],  dZ = grad_out * (out > 0) when relu.  grad_w must be
 * zeroed by the caller (partials are accumulated with atomics). */
int gs_sage_gemm_bwd_w(const float* self_table, int64_t ld_self, const int32_t* self_idx,
                       const float* agg, int64_t ld_agg, int32_t dim,
                       const float* grad_out, int64_t ld_go, const float* out, int64_t ld_out,
                       int32_t out_dim, int32_t gcn, int32_t relu,
                       const int32_t* num_rows_dev, int32_t max_rows,
                       float* grad_w, int64_t ldw, int32_t precision, gs_stream_t stream);

/* Up to three such problems -- the weight gradients of the layers of one step and of its classifier (gcn = 1, agg = the
 * classifier's input rows, grad_out = d(logits)), independent leaves of the step's dependency graph -- in one call.  Every
 * argument is a HOST array of n (device pointers / sizes per problem).  grad_out_cols_host (nullable; entries 0 =
 * out_dim): the zero-padded width of a problem's grad_out rows that may be read, so that a width that is not a multiple
 * of 4 (47 classes) still takes the 16-byte copy path.  With a tensor-core precision all run as ONE grid whose CTAs
 * are split in proportion to the work (rows x K x out_dim), so none queues behind another; otherwise one after the other. */
int gs_sage_gemm_bwd_w_group(int32_t n, const float* const* self_table_host, const int64_t* ld_self_host,
                             const int32_t* const* self_idx_host, const float* const* agg_host,
                             const int64_t* ld_agg_host, const int32_t* dim_host,
                             const float* const* grad_out_host, const int64_t* ld_go_host,
                             const float* const* out_host, const int64_t* ld_out_host,
                             const int32_t* out_dim_host, const int32_t* grad_out_cols_host,
                             const int32_t* gcn_host, const int32_t* relu_host,
                             const int32_t* const* num_rows_dev_host, const int32_t* max_rows_host,
                             float* const* grad_w_host, const int64_t* ldw_host, int32_t precision, gs_stream_t stream);

/* The two-problem form with common gcn / relu. */
int gs_sage_gemm_bwd_w_pair(const float* const* self_table_host, const int64_t* ld_self_host,
                            const int32_t* const* self_idx_host, const float* const* agg_host,
                            const int64_t* ld_agg_host, const int32_t* dim_host,
                            const float* const* grad_out_host, const int64_t* ld_go_host,
                            const float* const* out_host, const int64_t* ld_out_host,
                            const int32_t* out_dim_host, int32_t gcn, int32_t relu,
                            const int32_t* const* num_rows_dev_host, const int32_t* max_rows_host,
                            float* const* grad_w_host, const int64_t* ldw_host, int32_t precision, gs_stream_t stream);

/* dX[r,k] = sum_h dZ[r,h] W[h,k]  -> grad_self[r,:dim] (non-gcn) and grad_agg[r,:dim]. */
int gs_sage_gemm_bwd_x(const float* grad_out, int64_t ld_go, const float* out, int64_t ld_out,
                       const float* weight, int64_t ldw, int32_t dim, int32_t out_dim, int32_t gcn, int32_t relu,
                       const int32_t* num_rows_dev, int32_t max_rows,
                       float* grad_self, int64_t ld_gs, float* grad_agg, int64_t ld_ga, int32_t precision,
                       gs_stream_t stream);

/* ReLU backward in place: grad[r,c] = 0 where out[r,c] <= 0.  Used before the tensor-core
 * backward kernels (called with relu = 0), which stream dZ with cp.async. */
int gs_relu_bwd_inplace(float* grad, int64_t ld_g, const float* out, int64_t ld_out, int32_t dim,
                        const int32_t* num_rows_dev, int32_t max_rows, gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Classification, src/models.py:25-27: logp = log_softmax(emb . W^T + b).
 * ------------------------------------------------------------------------------------ */
int gs_cls_fwd(const float* emb, int64_t ld_emb, int32_t rows, int32_t dim,
               const float* weight, const float* bias, int32_t num_classes,
               float* logp, int32_t precision, gs_stream_t stream);
/* backward through log_softmax + Linear; grad_w / grad_b accumulate (zero them first);
 * grad_emb (nullable) is overwritten.  scratch: rows*num_classes floats (grad of the logits). */
int gs_cls_bwd(const float* grad_logp, const float* logp, const float* emb, int64_t ld_emb,
               int32_t rows, int32_t dim, const float* weight, int32_t num_classes,
               float* grad_emb, int64_t ld_ge, float* grad_w, float* grad_b, float* scratch, int32_t precision,
               gs_stream_t stream);
/* Supervised loss of src/utils.py:153,162-163 fused with its gradient:
 * y_r = labels[label_index ? label_index[r] : r]   (label_index = the batch's node ids, :153)
 * loss[0] = -mean_r logp[r, y_r];  grad_logp[r,c] = -(c==y_r) / rows  (nullable). */
int gs_nll_fwd_bwd(const float* logp, const int64_t* labels, const int32_t* label_index, int32_t rows,
                   int32_t num_classes, float* loss, float* grad_logp, gs_stream_t stream);

/* The supervised tail in one call (classifier forward, NLL mean of src/utils.py:153,162-163,
 * and their backward).  Heads with <= 64 classes and dim <= 256 (every configuration of the
 * reference) run as ONE kernel; larger ones as logits GEMM -> fused bias + log_softmax + NLL +
 * d logits + grad_b -> grad_w GEMM -> grad_emb GEMM.  loss[0] is overwritten; grad_w / grad_b
 * accumulate (zero them first); grad_emb (nullable) is overwritten; scratch: rows*num_classes
 * floats.  mask_relu_input != 0: emb is the ReLU output of the last SageLayer
 * (src/models.py:219) and grad_emb is returned already multiplied by (emb > 0).
 * zero_loss == 0: loss[0] was zeroed by the caller (off the critical path of a captured step)
 * and is accumulated into; != 0: it is zeroed here first.
 * num_rows_dev (nullable, device int32): only the first min(*num_rows_dev, rows) rows are a batch -- the mean
 * of the NLL is taken over them and grad_emb rows beyond them are left untouched (one-launch heads only). */
int gs_cls_nll_fwd_bwd(const float* emb, int64_t ld_emb, int32_t rows, int32_t dim,
                       const float* weight, const float* bias, int32_t num_classes,
                       const int64_t* labels, const int32_t* label_index,
                       float* logp, float* loss, float* grad_emb, int64_t ld_ge,
                       float* grad_w, float* grad_b, float* scratch, int32_t mask_relu_input, int32_t zero_loss,
                       const int32_t* num_rows_dev, int32_t precision, gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * The TOP layer of a supervised step in ONE launch (csrc/sage_top.cu): for the batch rows r < rows
 *   X[r]  = [ table[self_idx[r]] | mean_j table[nbr_idx[r][j]] ]   (gcn: the mean only)   src/models.py:260-266,300-314
 *   h[r]  = relu(W . X[r])                                                                src/models.py:215-219
 *   logp  = log_softmax(Wc . h + bc);  loss[0] = -mean_r logp[r, y_r]                     src/models.py:25-27, src/utils.py:162-163
 * and the whole backward of it: grad_cls_w / grad_cls_b accumulate (zero them first); out_dz = d(pre-activation) of
 * the layer; grad_table[t,:] += contributions of dX = dZ . W through the self gather and the mean, multiplied by
 * (table[t,:] > 0) -- `table` is the ReLU output of the layer below, so grad_table (zeroed by the caller, e.g. by
 * gs_sage_gemm_fwd_ex) receives that layer's d(pre-activation).  out_agg / out_dz are the B / A operands of this
 * layer's gs_sage_gemm_bwd_w (relu = 0).  out_h, logp, grad_table, grad_cls_* are nullable.
 * out_dlog (nullable, [rows x ld_dlog], ld_dlog >= 64): d(logits) of the batch rows, columns num_classes..63 zero.  With
 * it and grad_cls_w NULL the classifier's weight gradient dlog^T . h is left to gs_sage_gemm_bwd_w_group (problem with
 * gcn = 1, agg = out_h, grad_out = out_dlog, grad_out_cols = 64 or num_classes rounded up to 4): a handful of row chunks
 * there instead of one atomic add per CTA and element here.
 * Supported: dim == out_dim == 128, MEAN, num_classes <= 64, stride <= 16, precision TF32X3 / TF32 (mma.sync
 * m16n8k8 in the same 3-term split as K4); anything else returns GS_ERR_UNSUPPORTED and the caller runs the layer
 * as gs_agg_fwd -> gs_sage_gemm_fwd -> gs_cls_nll_fwd_bwd -> gs_sage_gemm_bwd_x -> gs_agg_bwd.
 * cls_reps (1..16) > 1: CTA b adds its classifier-gradient contribution into replica b % cls_reps -- replica 0 is
 * grad_cls_w / grad_cls_b, replica r > 0 is cls_w_replicas + (r-1)*num_classes*128 and cls_b_replicas + (r-1)*64 (both
 * zeroed by the caller; gs_dp_allreduce_clip_sgd folds them in and clears them, see seg_extra_host): 64 CTAs adding into
 * one block serialise in the L2 atomic units otherwise.
 * With gs_set_early_reads(1) the index lists, labels and row count are read before the wait for the previous kernels
 * of the stream (see there for when that is allowed).
 * workspace: gs_sage_top_workspace_bytes() of device memory, zeroed ONCE by the caller (loss partials + a ticket
 * that every launch leaves at zero again); loss[0] is overwritten, nothing needs zeroing per step.
 * ------------------------------------------------------------------------------------ */
size_t gs_sage_top_workspace_bytes(void);
int gs_sage_top_sup(const float* table, int64_t ld_table, const int32_t* nbr_idx, int32_t stride,
                    const int32_t* cnt, const int32_t* self_idx, const int32_t* num_rows_dev, int32_t max_rows,
                    const float* weight, int64_t ldw, int32_t dim, int32_t out_dim, int32_t gcn,
                    const float* cls_w, const float* cls_b, int32_t num_classes, const int64_t* labels,
                    const int32_t* label_index, float* out_h, int64_t ld_h, float* out_agg, int64_t ld_agg,
                    float* out_dz, int64_t ld_dz, float* out_dlog, int64_t ld_dlog, float* logp, float* loss,
                    float* grad_cls_w, float* grad_cls_b, float* grad_table, int64_t ld_gt, void* workspace,
                    size_t workspace_bytes, int32_t precision, float* cls_w_replicas, float* cls_b_replicas,
                    int32_t cls_reps, gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Update step of src/utils.py:185-187: per-model clip_grad_norm_(max_norm) then SGD.
 * params / grads / numels are DEVICE arrays describing the `num_tensors` tensors of ONE
 * model (max_numel = the largest numel, sizes the grid); grads are divided by `grad_div`
 * first (data-parallel mean of summed gradients); max_norm <= 0 disables clipping;
 * zero_grads != 0 also clears the gradients (src/utils.py:189-191).  norm_scratch: 1 float.
 * ------------------------------------------------------------------------------------ */
int gs_clip_sgd(float* const* params, float* const* grads, const int64_t* numels, int32_t num_tensors,
                int64_t max_numel, float max_norm, float lr, float grad_div, int32_t zero_grads,
                float* norm_scratch, gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Data-parallel exchange + update in ONE kernel (src/utils.py:184-191 on the
 * mean gradient of all ranks): all-reduce of the flat gradient buffer over NVLink peer memory
 * (push into every peer's receive slot, per-slice flags, rank-ordered sum => bit-identical
 * replicas), per-model clip_grad_norm_(max_norm), SGD(lr), gradients zeroed.
 *   flat_grad      this rank's flat fp32 gradient (n_total % 4 == 0), summed in place
 *   peer_regions_host  HOST array of `world` device pointers: rank r's exchange region of
 *                  gs_dp_region_bytes(n_total, world) bytes (zeroed; own entry = local
 *                  pointer, others = gs_peer_open mappings).  NULL when world == 1.
 *   seg_*_host     HOST arrays describing the `num_segs` (<= 16) parameter tensors: device
 *                  pointer, offset of its gradient inside flat_grad (multiple of 4), numel,
 *                  clip group (0..3; the reference clips each model separately, :185-186)
 *   state          gs_dp_state_bytes() of zeroed device memory, private to this (flat_grad,
 *                  n_total); carries the epoch between calls
 *   timeout_ns     a peer that does not arrive within this time sets status != 0 (see
 *                  gs_dp_status) instead of hanging the GPU; 0 => 2 s
 *   step_counter   nullable device int64, incremented by one (the sampler's Philox offset_dev,
 *                  so a captured step draws fresh neighbours on every replay)
 * ------------------------------------------------------------------------------------ */
size_t gs_dp_state_bytes(void);
size_t gs_dp_region_bytes(int64_t n_total, int32_t world);
size_t gs_dp_region_recv_offset(void);
int gs_dp_allreduce_clip_sgd(float* flat_grad, int64_t n_total, void* const* peer_regions_host, int32_t rank,
                             int32_t world, float* const* seg_params_host, const int64_t* seg_offsets_host,
                             const int64_t* seg_numels_host, const int32_t* seg_groups_host, int32_t num_segs,
                             float max_norm, float lr, void* state, uint64_t timeout_ns, int64_t* step_counter,
                             float* const* seg_params_lo_host, float* const* seg_extra_host,
                             const int32_t* seg_extra_n_host, const int64_t* seg_extra_stride_host, gs_stream_t stream);
/* seg_extra_host (nullable HOST array of num_segs nullable device pointers): tensor k has seg_extra_n_host[k] (1..16) more
 * partial gradients, seg_extra_stride_host[k] floats apart (multiple of 4, >= numel rounded up to 4), which a producer
 * spread its atomic adds over (gs_sage_top_sup's cls_reps); they are added to the flat gradient before the exchange and
 * cleared. */
/* seg_params_lo_host (nullable HOST array of num_segs nullable device pointers): buffers shaped like the parameters that
 * receive p - trunc_tf32(p) of every updated element (gs_sage_gemm_fwd_ex's weight_lo); gs_split_lo initialises one. */
int gs_split_lo(const float* src, float* dst, int64_t n, gs_stream_t stream);
/* synchronises `stream`, then reports the epoch counter, the status word (0 ok, 1 peer
 * wait timed out, 2 grid barrier timed out) and the last step's 4 per-group gradient norms */
int gs_dp_status(const void* state, uint32_t* epoch_host, uint32_t* status_host, float* norms_host, gs_stream_t stream);

/* Peer memory for the two uses above: plain cudaMalloc'ed, zero-filled regions exported with
 * CUDA IPC (one process per GPU; handles are 64 bytes and travel over any host channel). */
int gs_peer_alloc(size_t bytes, void** out_ptr_host);
int gs_peer_free(void* ptr);
int gs_peer_export(void* ptr, unsigned char* handle64_host);
int gs_peer_open(const unsigned char* handle64_host, void** out_ptr_host);
int gs_peer_close(void* ptr);

/* ------------------------------------------------------------------------------------
 * K5  UnsupervisedLoss sampling, src/models.py:153-186.
 * gs_random_walk_pos: n_walks walks of walk_len steps per seed (6 x 1 in the reference,
 *   :50-51); a step yields a pair when next != seed and is_train[next] (:180).
 *   pos[(s*n_walks+w)*walk_len+t] = next or -1.  Zero-degree seeds produce no pairs (:171-172).
 * gs_negative_sample: `num_neg` distinct train nodes outside the `hops`-hop ball of each
 *   seed (:155-164).  The ball is marked in a per-seed bitmap of `num_nodes` bits held in
 *   `workspace`; when fewer than num_neg far train nodes exist all of them are returned.
 * offset_dev (nullable, device int64), as in gs_sample_neighbors: the Philox offset used is
 *   offset + (*offset_dev << 8), so a captured loop draws fresh pairs every replay.
 * ------------------------------------------------------------------------------------ */
int gs_random_walk_pos(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                       const int32_t* seeds, int32_t num_seeds, int32_t n_walks, int32_t walk_len,
                       const uint8_t* is_train, uint64_t seed, uint64_t offset, const int64_t* offset_dev,
                       int32_t* pos, gs_stream_t stream);
size_t gs_negative_workspace_bytes(int64_t num_nodes, int32_t num_seeds);
int gs_negative_sample(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                       const int32_t* seeds, int32_t num_seeds, int32_t hops, int32_t num_neg,
                       const int32_t* train_nodes, int32_t num_train, uint64_t seed, uint64_t offset,
                       const int64_t* offset_dev, int32_t* neg, int32_t* neg_cnt, void* workspace, size_t workspace_bytes,
                       gs_stream_t stream);

/* is_train (nullable, num_nodes bytes, 1 for the ids in train_nodes, which must be distinct): with it the number of far
 * train nodes is known from the marking pass (|train| - |train nodes in the ball|), and when at least 4 * num_neg of them
 * exist the negatives are drawn by rejection sampling (uniform train node, rejected when inside the ball or already
 * taken) instead of two walks over the whole train list per seed; same distribution, order of acceptance. */
int gs_negative_sample_ex(const int64_t* rowptr, const int32_t* col, int64_t num_nodes,
                          const int32_t* seeds, int32_t num_seeds, int32_t hops, int32_t num_neg,
                          const int32_t* train_nodes, int32_t num_train, const uint8_t* is_train,
                          uint64_t seed, uint64_t offset, const int64_t* offset_dev, int32_t* neg, int32_t* neg_cnt,
                          void* workspace, size_t workspace_bytes, gs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K6  pair losses, src/models.py:65-132.  Pairs are grouped per seed: seed s owns
 * pos_idx[pos_ptr[s]..pos_ptr[s+1]) and neg_idx[neg_ptr[s]..), all row indices into `emb`
 * (seed_idx[s] is the seed's own row).  Seeds with no positive or no negative pair are
 * skipped (:75-76, :110-111).  mode 0 = "normal" (get_loss_sage, Q), 1 = margin.
 * fwd writes loss[0] and per-pair coefficients into `coef` (len = total pairs) for bwd.
 * ------------------------------------------------------------------------------------ */
int gs_pair_loss_fwd(const float* emb, int64_t ld, int32_t dim,
                     const int32_t* seed_idx, int32_t num_seeds,
                     const int32_t* pos_ptr, const int32_t* pos_idx,
                     const int32_t* neg_ptr, const int32_t* neg_idx,
                     int32_t mode, float q, float margin,
                     float* loss, float* coef_pos, float* coef_neg, float* loss_sum_scratch, int32_t* num_active,
                     gs_stream_t stream);
int gs_pair_loss_bwd(const float* emb, int64_t ld, int32_t dim,
                     const int32_t* seed_idx, int32_t num_seeds,
                     const int32_t* pos_ptr, const int32_t* pos_idx,
                     const int32_t* neg_ptr, const int32_t* neg_idx,
                     const float* coef_pos, const float* coef_neg, const int32_t* num_active,
                     const float* grad_loss, float* grad_emb, int64_t ld_ge, gs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GSAGE_B200_H_ */
