/*
 * dreamgnn.h -- C ABI of libdreamgnn.so: hand-written sm_100a kernels for DREAM-GNN's
 * message-passing hot path (GCMC relation layers, FGCN GraphConv, MLP-decoder edge gather,
 * graph / normaliser / kNN construction and the per-iteration augmentation rebuild).
 *
 * The reference (Ryan-Yanlong/DREAM-GNN) is pure Python and has no FFI of its own; each entry
 * point below replaces the third-party call (DGL / torch.sparse / numpy / scipy) the reference
 * makes at the cited file:line. Citations are into /root/reference/.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless a parameter says "host". The caller (PyTorch in the
 *     shipped host code) owns all memory: inputs, outputs and workspaces. Kernels never allocate.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and never
 *     synchronises the device. The library keeps no mutable global state besides the launch counter
 *     and a thread-local error string, so calls are re-entrant across host threads / streams.
 *   - Return value: 0 on success; DG_ERR_* (<0) for argument errors; a positive cudaError_t when a
 *     launch failed. dg_last_error() gives a thread-local message. No C++ exception crosses the ABI.
 *   - Indices are int32 (the reference narrows with graph.int(), train.py:199-200). Feature matrices
 *     are row-major with an explicit leading dimension in ELEMENTS; rows must be 16-byte aligned
 *     (ld % 4 == 0 for fp32) and the width d a multiple of 4 -- the host pads 341 -> 344.
 *   - `*_workspace_bytes` are pure host functions.
 */
#ifndef DREAMGNN_H_
#define DREAMGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_ABI_VERSION 5

#define DG_OK 0
#define DG_ERR_INVALID_ARGUMENT (-1)
#define DG_ERR_WORKSPACE_TOO_SMALL (-2)
#define DG_ERR_UNSUPPORTED (-3)

typedef void* dg_stream_t; /* cudaStream_t */

/* exported with default visibility; everything else in the library is hidden */
#if defined(__GNUC__)
#define DG_API __attribute__((visibility("default")))
#else
#define DG_API
#endif

/* ---- library bookkeeping ------------------------------------------------------------------ */
DG_API int dg_abi_version(void);
DG_API const char* dg_last_error(void);
/* Number of kernels this library has launched since load / last reset (bench.py "gpu_launches"). */
DG_API unsigned long long dg_launch_count(void);
DG_API void dg_reset_launch_count(void);

/* ---- index primitives ---------------------------------------------------------------------- */
/* Exclusive prefix sum of n int32 values; out[n] (one past the end) receives the total, so `out`
 * must hold n+1 elements. in == out is allowed only if out is sized n+1 and in aliases out[0..n). */
DG_API size_t dg_scan_workspace_bytes(int64_t n);
DG_API int dg_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n,
                          void* workspace, size_t workspace_bytes, dg_stream_t stream);

/* Stable LSD radix sort of n (uint64 key, int32 value) pairs on the low `key_bits` bits.
 * Results land in keys_out / vals_out; keys_in / vals_in are clobbered. vals may be NULL. */
DG_API size_t dg_sort_workspace_bytes(int64_t n);
DG_API int dg_sort_pairs_u64(uint64_t* keys_in, int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out,
                      int64_t n, int key_bits, void* workspace, size_t workspace_bytes,
                      dg_stream_t stream);

/* ---- graph construction (replaces DGL's lazy COO->CSC/CSR inside update_all, layers.py:229-232,
 *      and dgl.heterograph / in_degrees / out_degrees at data_loader.py:448-488) ---------------- */

/* COO -> canonical CSR: rows ascending, columns ascending inside a row, ties by edge id.
 *   row/col [n_edges]; indptr [n_rows+1]; indices [n_edges]; eid [n_edges] = input position of the
 *   edge stored in each slot (the permutation edge values / transposes are carried with).
 * Deterministic: the result does not depend on the order of the input list. */
DG_API size_t dg_csr_build_workspace_bytes(int64_t n_edges, int64_t n_rows);
DG_API int dg_csr_build(const int32_t* row, const int32_t* col, int64_t n_edges, int64_t n_rows,
                 int64_t n_cols, int32_t* indptr, int32_t* indices, int32_t* eid,
                 void* workspace, size_t workspace_bytes, dg_stream_t stream);

/* norm[i] = 1/sqrt(indptr[i+1]-indptr[i]) as fp32 with degree 0 -> 0; bit-exact with
 * data_loader.py:454-457 (_calc_norm) when indptr is the CSR of all relations into node type i. */
DG_API int dg_degree_norm(const int32_t* indptr, int64_t n_rows, float* norm, dg_stream_t stream);

/* ---- per-iteration edge dropout rebuild (augmentation.py:13-124) --------------------------- */
/* flags[offset + perm[i]] = 1 for i < num_keep: "keep edge e iff rank(e) < num_keep" with
 * num_keep = max(1, int(E*(1-rate))) evaluated by the caller (augmentation.py:48-52).
 * flags must be zeroed by the caller beforehand. perm is torch.randperm's int64 output. */
DG_API int dg_keep_flags_from_perm(const int64_t* perm, int64_t num_keep, int64_t offset,
                            uint8_t* flags, dg_stream_t stream);

/* Uniform random num_keep-subset of n edges as a flag array, without a sort: flags[i] = 1 for the num_keep smallest of the
 * keys (rnd[i] << bits(n-1)) | i, found by an 8-pass radix SELECT (histogram + pick per digit). rnd = n random int64 drawn
 * by the caller (63 random bits each, e.g. torch's Tensor.random_()). Same distribution as "the first num_keep entries of a
 * random permutation" (augmentation.py:48-52, :113-118) at 9 streaming reads of rnd instead of a 7-pass radix sort of
 * (key, index) pairs. Used inside captured CUDA graphs; the eager path keeps th.randperm + dg_keep_flags_from_perm, whose
 * kept sets are the reference's for a given generator state. Exactly num_keep flags are set (keys are distinct). */
DG_API size_t dg_random_subset_workspace_bytes(void);
DG_API int dg_random_subset_flags(const int64_t* rnd, int64_t n, int64_t num_keep, uint8_t* flags, void* workspace,
                                  size_t workspace_bytes, dg_stream_t stream);

/* Compact a canonical CSR to the edges whose flag (indexed by eid) is set, preserving order, so
 * the dropped graph's CSR is again canonical without a sort. vals / out_vals may be NULL.
 * out_indices / out_eid / out_vals must hold the number of kept edges (known to the caller). */
DG_API size_t dg_csr_compact_workspace_bytes(int64_t n_rows);
DG_API int dg_csr_compact(const int32_t* indptr, const int32_t* indices, const int32_t* eid,
                   const float* vals, int64_t n_rows, const uint8_t* keep_by_eid,
                   int32_t* out_indptr, int32_t* out_indices, int32_t* out_eid, float* out_vals,
                   void* workspace, size_t workspace_bytes, dg_stream_t stream);

/* CSR -> COO row ids: row[s] = r for indptr[r] <= s < indptr[r+1]. */
DG_API int dg_csr_expand_rows(const int32_t* indptr, int64_t n_rows, int32_t* row, dg_stream_t stream);

/* ---- SpMM: out[i,:] = epilogue( dst_scale[i] * sum_{s in row i} vals[s] * src_scale[j_s] * x[j_s,:] )
 *      with j_s = indices[s].
 *   Replaces DGL update_all(copy_u,sum) with the cj / ci normalisers fused (layers.py:224-234) when
 *   vals == NULL, and torch.spmm(adj, support) (+ bias, layers.py:312-314) when vals != NULL.
 *   The backward is the same call on the transposed CSR with the two scales swapped -- atomic-free
 *   and bit-reproducible. vals, src_scale, dst_scale, bias may each be NULL.
 *   flags: DG_SPMM_ACCUMULATE adds into `out`; DG_SPMM_RELU applies max(.,0) last; DG_SPMM_PREFETCH is a
 *   hint that the gathered operand does not stay resident in L2 (x larger than ~half of it): the kernel then
 *   issues L2 prefetches (prefetch.global.L2) for the rows it will gather one group ahead, so the bytes in flight
 *   towards HBM are not bounded by the registers holding demand loads. Results are identical with and without. */
#define DG_SPMM_ACCUMULATE 1
#define DG_SPMM_RELU 2
#define DG_SPMM_PREFETCH 4
/* DG_SPMM_ROWSPLIT: few, long rows (the real datasets: ~700 rows of ~600 edges) -- one CTA per row whose 8 warps take
 * contiguous eighths of the row and are summed in warp order (deterministic; the rounding order differs from the
 * warp-per-row kernel's). The host sets it when warp-per-row would leave most SMs idle. */
#define DG_SPMM_ROWSPLIT 8
DG_API int dg_spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* vals,
                    const float* src_scale, const float* dst_scale, const float* bias,
                    const float* x, int64_t ldx, float* out, int64_t ldo,
                    int64_t n_rows, int64_t d, int flags, dg_stream_t stream);
/* bf16 feature storage, fp32 accumulate, fp32 output (the 2e-2 path). x is __nv_bfloat16 row-major,
 * ldx % 8 == 0 and d % 8 == 0. */
DG_API int dg_spmm_csr_bf16(const int32_t* indptr, const int32_t* indices, const float* vals,
                     const float* src_scale, const float* dst_scale, const float* bias,
                     const void* x_bf16, int64_t ldx, float* out, int64_t ldo,
                     int64_t n_rows, int64_t d, int flags, dg_stream_t stream);

/* ---- MLP decoder over scored pairs (layers.py:341-379) --------------------------------------
 * lin1 is split exactly: lin1(cat(a,b)) = a W1a^T + b W1b^T + b1, so the caller passes the two
 * node-level projections pd = h_drug W1a^T + b1 [n_drug,128] and ps = h_dis W1b^T [n_dis,128] and
 * the kernel gathers them per pair (SDDMM-style) instead of materialising the [E,256] concat:
 *   z1 = drop(relu(pd[src]+ps[dst]));  z2 = drop(relu(W2 z1 + b2));  out = w3.z2 + b3
 * Hidden widths 128 / 64 are fixed by the reference (layers.py:349-351). Dropout uses a counter
 * based generator keyed by (seed, pair, unit); p == 0 disables it (eval mode). When `seed_dev` is non-NULL
 * the seed is read from that device word instead of `seed`, so a captured CUDA graph draws a fresh mask
 * on every replay.
 * `perm` (nullable, [n_pairs]): src/dst are given in a caller-chosen processing order and perm[i] is the label
 * position of the i-th processed pair -- out[perm[i]] receives its logit and the backward reads dout[perm[i]];
 * z2_save and dz1 stay in processing order. The host sorts the pairs by drug once per decoder graph, so
 * consecutive pairs share their pd row and only the ps rows are a random gather (the label-order gather of both
 * operands misses L2 at the 20M-pair shape: 77 MB of node rows against ~63 MB of L2 per die). NULL = identity. */
#define DG_DEC_H1 128
#define DG_DEC_H2 64
DG_API int dg_decoder_fwd_f32(const int32_t* src, const int32_t* dst, const int32_t* perm, int64_t n_pairs,
                       const float* pd, const float* ps, const float* w2, const float* b2,
                       const float* w3, const float* b3, float dropout_p, uint64_t seed,
                       const uint64_t* seed_dev, float* out, float* z2_save /* nullable: [n_pairs,64] kept for backward */,
                       dg_stream_t stream);
/* Backward: regenerates z1 (same dropout mask), reads the z2 the forward saved, writes
 * dz1 [n_pairs,128] = d loss / d (pd[src]+ps[dst]) and the parameter gradients dw2 [64,128], db2 [64],
 * dw3 [64], db3 [1] (per-CTA partials summed in CTA order). d pd / d ps are then two deterministic
 * segment sums of dz1 over the decoder graph's CSR / CSC (dg_spmm_csr_f32 with indices = edge ids)
 * -- no atomics. Optionally the sum by SOURCE node is fused into the kernel's epilogue: pair_slot [n_pairs]
 * (16-byte aligned) numbers, in processing order, the runs of equal source inside aligned 16-pair groups (a new
 * slot starts at every multiple of 16 and wherever the source changes); the kernel writes slot_rows
 * [n_slots, 128] = sum of dz1 over each run, and the caller adds the slots of each node (a segment sum over
 * n_slots ~ n_pairs / 16 rows instead of n_pairs). Both NULL = off. Tensor-core kernel only. */
DG_API size_t dg_decoder_bwd_workspace_bytes(int64_t n_pairs);
DG_API int dg_decoder_bwd_f32(const int32_t* src, const int32_t* dst, const int32_t* perm, int64_t n_pairs,
                       const float* pd, const float* ps, const float* w2, const float* w3,
                       float dropout_p, uint64_t seed, const uint64_t* seed_dev, const float* z2,
                       const float* dout,
                       float* dz1, float* dw2, float* db2, float* dw3, float* db3,
                       const int32_t* pair_slot, float* slot_rows,
                       void* workspace, size_t workspace_bytes, dg_stream_t stream);

/* ---- dense projections on the tcgen05 tensor cores -------------------------------------------------
 * C[b] = diag(row_scale[b]) * A[b] * B[b]^T for b < batch; A [M,K] row-major (lda), B [N,K] row-major
 * (ldb), C [M,N] row-major (ldc); batch strides in elements, stride 0 = operand shared by all batches.
 * precision 0: error-compensated 3xTF32 (fp32-level accuracy, the 1e-5 parity path); 1: single TF32.
 * Replaces the cuBLAS GEMMs behind th.matmul(att, basis) @ feat (layers.py:120-121, 220-221, 392),
 * th.mm(input, weight) (layers.py:311) and their backward. row_scale may be NULL. */
DG_API size_t dg_gemm_nt_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch, int a_batched, int b_batched);
DG_API int dg_gemm_nt_f32(const float* A, int64_t lda, int64_t stride_a, const float* B, int64_t ldb,
                          int64_t stride_b, float* C, int64_t ldc, int64_t stride_c, int64_t M, int64_t N,
                          int64_t K, int64_t batch, const float* row_scale, int precision, void* workspace,
                          size_t workspace_bytes, dg_stream_t stream);
/* General form: trans_a != 0 means A is given as [K,M] row-major (lda >= M), trans_b != 0 means B is given as
 * [K,N] row-major (ldb >= N); the tensor cores read such operands MN-major, so no transposed copy is made.
 * C = op(A) * op(B)^T with op(A) [M,K], op(B) [N,K]; dg_gemm_nt_f32 is the trans_a = trans_b = 0 case. The weight
 * gradients dW_r = X^T dY_r (backward of layers.py:220-221, 311) use trans_a = trans_b = 1 on X and dY as stored.
 * Workspace bound: dg_gemm_nt_workspace_bytes (valid for every orientation). */
DG_API int dg_gemm_f32(const float* A, int64_t lda, int64_t stride_a, int trans_a, const float* B, int64_t ldb,
                       int64_t stride_b, int trans_b, float* C, int64_t ldc, int64_t stride_c, int64_t M, int64_t N,
                       int64_t K, int64_t batch, const float* row_scale, int precision, void* workspace,
                       size_t workspace_bytes, dg_stream_t stream);

/* fp32 GEMM for the small dense layers of the real-dataset shapes (csrc/small_gemm.cu): C[b] = op(A[b]) . op(B[b])^T (+ bias
 * [N]), same operand conventions as dg_gemm_f32 (op(A) [M, K], op(B) [N, K]; trans_x: stored [K, M] / [K, N]; stride 0 = an
 * operand shared by the batch), but plain fp32 FMAs in ascending k, no alignment requirement, ONE launch. reduce_batch != 0
 * sums over the batch into a single C [M, N] (the batch is walked inside the k-loop). Very long K with few output tiles is
 * split over CTAs and the partial tiles are added inside the kernel by the last CTA of each tile in split order
 * (deterministic). Replaces th.addmm / th.matmul (cuBLAS SIMT sgemm + split-K reduction + bias epilogue launches) behind
 * nn.Linear at layers.py:139-142, 281-282, 366-369 and the 128 -> 128 projections of layers.py:220-221.
 * tickets: dg_small_gemm_tickets(M, N, K, batch) device int32 (0 = none needed: NULL is fine), ZERO on entry and zero again
 * on completion, not shared with a call that may run concurrently. */
DG_API size_t dg_small_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch);
DG_API int64_t dg_small_gemm_tickets(int64_t M, int64_t N, int64_t K, int64_t batch);
DG_API int dg_small_gemm_f32(const float* A, int64_t lda, int64_t stride_a, int trans_a, const float* B, int64_t ldb,
                      int64_t stride_b, int trans_b, const float* bias, float* C, int64_t ldc, int64_t stride_c, int64_t M,
                      int64_t N, int64_t K, int64_t batch, int reduce_batch, void* workspace, size_t workspace_bytes,
                      int32_t* tickets, dg_stream_t stream);

/* ---- row-streaming helpers around the aggregation kernels ------------------------------------ */
/* Deterministic column sums of a row-major fp32 matrix (fixed slab partition, fixed summation order, no atomics):
 *   out[j] = sum_i y[i][j],  y = x            (gate == NULL)
 *                            y = x * (gate>0) (gate != NULL: the ReLU mask of the aggregation epilogue)
 * and, when y != NULL, y is written ([n_rows, d], leading dimension ldy; must not alias x). One pass over the matrix.
 * Replaces autograd's `grad.sum(0)` for the GraphConvolution / nn.Linear bias gradients (backward of layers.py:311-314,
 * layers.py:139-142) and the `grad * (out > 0)` of F.relu's backward (layers.py:247), and gives the column means of
 * utils.py:89-90. d % 4 == 0, rows 16-byte aligned.
 * tickets (may be NULL): ceil(d / 128) device int32 that are ZERO on entry and zero again on completion, not shared with
 * any call that may run concurrently. With them a matrix of <= 16 384 rows takes ONE launch (the CTA that finishes last
 * adds the slab partials, in slab order: the result does not depend on which CTA that is) instead of two. */
DG_API size_t dg_colsum_workspace_bytes(int64_t n_rows, int64_t d);
DG_API int dg_colsum_f32(const float* x, int64_t ldx, const float* gate, int64_t ldg, float* y, int64_t ldy,
                         int64_t n_rows, int64_t d, float* out, void* workspace, size_t workspace_bytes,
                         int32_t* tickets, dg_stream_t stream);
/* Rows of common_loss (utils.py:87-95): z[i,:] = c_i / max(||c_i||_2, eps) with c_i = x[i,:] - colsum/n_rows, widened to
 * float64 (the Gram-matrix form of the loss accumulates in float64); inv_norm[i] = 1 / max(||c_i||, eps).
 * z is [n_rows, d] with leading dimension ldz (two embeddings are written side by side into one [n, 2d] operand). */
DG_API int dg_center_normalize_f64(const float* x, int64_t ldx, const float* colsum, int64_t n_rows, int64_t d,
                                   double eps, double* z, int64_t ldz, double* inv_norm, dg_stream_t stream);
/* Backward of the row normalisation: dc[i,:] = (gz[i,:] - z[i,:] * <z[i,:], gz[i,:]>) * inv_norm[i] * s (fp32 out) with
 * s = scale * (*gout) (gout: device float, the upstream scalar gradient; NULL = 1) -- the product is linear in gz, so the
 * loss's constant factor and the upstream gradient need no pass of their own. The caller subtracts the column mean of dc
 * afterwards (backward of the centring). */
DG_API int dg_center_normalize_bwd_f64(const double* gz, int64_t ldgz, const double* z, int64_t ldz,
                                       const double* inv_norm, int64_t n_rows, int64_t d, float* dc, int64_t lddc,
                                       const float* gout, double scale, dg_stream_t stream);

/* ---- kNN similarity graphs (data_loader.py:278-344, utils.py:11-27) ------------------------- */
/* Per row of a float64 similarity block [n_rows, n_cols] (leading dimension ld), the k largest
 * entries under the tie rule (value descending, column ascending); out_idx [n_rows,k] ascending.
 * Replaces np.argpartition(-S, k)[:, :k] (data_loader.py:293). k <= 64. */
DG_API int dg_topk_rows_f64(const double* sim, int64_t n_rows, int64_t n_cols, int64_t ld, int k,
                     int32_t* out_idx, dg_stream_t stream);
/* Neighbour lists -> row-normalised symmetric kNN adjacency (A+A^T, +I, D^-1 A in float64, cast to
 * fp32; data_loader.py:294-308 + utils.py:11-17) as a canonical CSR/COO.
 *   nbr [n,k]; outputs sized for the worst case nnz_max = 2*n*k + n: row/col/val [nnz_max],
 *   indptr [n+1]; *nnz_out (device int32) receives the number of distinct entries. */
DG_API size_t dg_knn_graph_workspace_bytes(int64_t n, int k);
DG_API int dg_knn_graph_from_neighbors(const int32_t* nbr, int64_t n, int k, int32_t* indptr,
                                int32_t* row, int32_t* col, float* val, int32_t* nnz_out,
                                void* workspace, size_t workspace_bytes, dg_stream_t stream);

/* ---- fused row kernels between the aggregations (csrc/fused.cu) ---------------------------- */
/* y = dropout(act(x)) (dy == NULL) or its backward dx = dy * act'(x) * keep (dy != NULL), act: 0 identity, 1 LeakyReLU(slope),
 * 2 ReLU. Replaces `agg_act` -> `self.dropout` before ufc / ifc (layers.py:134-138) and F.dropout in GCN / FGCN
 * (layers.py:247, 281-282): one launch each way, the keep mask is a counter hash of (seed, row, column group) that the
 * backward regenerates. seed_dev (device uint64, may be NULL) overrides seed (CUDA-graph replays). d % 4 == 0. */
DG_API int dg_act_dropout_f32(const float* x, int64_t ldx, const float* dy, int64_t lddy, float* out, int64_t ldo,
                       int64_t n_rows, int64_t d, int act, float slope, float p, uint64_t seed,
                       const uint64_t* seed_dev, dg_stream_t stream);
/* Attention.forward (layers.py:324-338) over K = 2 views given as two [n_rows, d] matrices (the reference stacks them):
 * w_k = w2 . tanh(W1 z_k + b1), beta = dropout(softmax_k(w)), out = sum_k beta_k z_k; beta_out [n_rows, 2] optional.
 * W1 [hidden, d] row-major (nn.Linear weight), hidden <= 16, d % 4 == 0, d <= 512. */
DG_API int dg_attention_fwd_f32(const float* za, int64_t lda, const float* zb, int64_t ldb, int64_t n_rows, int64_t d,
                         const float* w1, const float* b1, const float* w2, int hidden, float p, uint64_t seed,
                         const uint64_t* seed_dev, float* out, int64_t ldo, float* beta_out, dg_stream_t stream);
/* Backward: recomputes the forward from za / zb, writes dza / dzb (either may be NULL) and the parameter gradients packed
 * as out_params = dW1 [16, d] (rows >= hidden zero) | db1 [16] | dw2 [16]; dbeta [n_rows, 2] (gradient of beta_out) may be
 * NULL. Deterministic: per-warp register accumulators summed over warps and CTAs in fixed order. */
DG_API size_t dg_attention_bwd_workspace_bytes(int64_t n_rows, int64_t d);
DG_API int dg_attention_bwd_f32(const float* za, int64_t lda, const float* zb, int64_t ldb, int64_t n_rows, int64_t d,
                         const float* w1, const float* b1, const float* w2, int hidden, float p, uint64_t seed,
                         const uint64_t* seed_dev, const float* dout, int64_t lddo, const float* dbeta, float* dza,
                         int64_t ldda, float* dzb, int64_t lddb, float* out_params, void* workspace,
                         size_t workspace_bytes, dg_stream_t stream);

/* ---- loss and optimiser tail of the iteration (csrc/loss.cu, csrc/optim.cu) ------------------ */
/* Mean binary cross entropy with logits over n scored pairs, targets optionally smoothed to t (1 - s) + s / 2
 * (nn.BCEWithLogitsLoss at train.py:291, LabelSmoothingBCELoss train.py:15-23): *loss (device float) =
 * mean_i [max(x, 0) - x t + log1p(exp(-|x|))], summed in float64 in a fixed order. Two launches. */
DG_API size_t dg_bce_logits_workspace_bytes(int64_t n);
DG_API int dg_bce_logits_fwd_f32(const float* logits, const float* target, int64_t n, float smoothing, float* loss,
                          void* workspace, size_t workspace_bytes, dg_stream_t stream);
/* Its gradient: dlogits[i] = *gout * (sigmoid(x_i) - t_i) / n; gout is a device float (the upstream gradient). */
DG_API int dg_bce_logits_bwd_f32(const float* logits, const float* target, int64_t n, float smoothing, const float* gout,
                          float* dlogits, dg_stream_t stream);
/* Tail of common_loss (utils.py:87-95) in Gram form: for G = Z^T Z [2d, 2d] float64 (leading dimension ldg) of the two
 * centred, normalised embeddings side by side, gs [2d, 2d] (dense) = G * S with S = +1 on the two diagonal d x d blocks
 * and -1 off them, and *loss (device float) = sum(G * gs) / n_rows^2 = (|G11|^2 + |G22|^2 - 2 |G12|^2) / n^2. One launch. */
DG_API int dg_gram_common_loss_f64(const double* G, int64_t ldg, int64_t d, double n_rows, double* gs, float* loss,
                            dg_stream_t stream);

/* Basis decomposition of the GCMC relation weights, W_r = sum_b att[r][b] * basis[b] (layers.py:120-121; the reference:
 * matmul(att, basis.view(B, -1))): w [n_rel, rows, d_pad] with the columns [d, d_pad) zero (the aggregation kernels'
 * width padding written in the same pass); att [n_rel, n_basis], basis [n_basis, rows, d] dense. n_rel, n_basis <= 4. */
DG_API int dg_basis_combine_fwd_f32(const float* att, const float* basis, int n_rel, int n_basis, int64_t rows, int64_t d,
                             int64_t d_pad, float* w, dg_stream_t stream);
/* Its backward from dw [n_rel, rows, d_pad]: dbasis[b] = sum_r att[r][b] * dw[r] (dense [n_basis, rows, d]) and
 * datt[r][b] = <dw[r], basis[b]> summed in float64 in a fixed order. Two launches. */
DG_API size_t dg_basis_combine_bwd_workspace_bytes(int64_t rows, int64_t d);
DG_API int dg_basis_combine_bwd_f32(const float* att, const float* basis, const float* dw, int n_rel, int n_basis, int64_t rows,
                             int64_t d, int64_t d_pad, float* dbasis, float* datt, void* workspace, size_t workspace_bytes,
                             dg_stream_t stream);

/* nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step() (train.py:297-300) over a list of fp32
 * tensors in two launches per DG_ADAM_MAX_TENSORS_PER_LAUNCH tensors: the global gradient norm from per-chunk float64
 * partial sums added in a fixed order, then per element
 *   g <- g * min(1, max_norm / (norm + 1e-6))        (written back to grad; max_norm <= 0: no clipping)
 *   g' = g + weight_decay * p;  m <- m + (g' - m)(1 - beta1);  v <- beta2 v + (1 - beta2) g'^2
 *   p <- p - lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
 * `tensors` is a HOST array (its entries are copied into the kernel arguments); every pointer inside is a device pointer
 * to `numel` contiguous floats. `step` (device float) holds t - 1 on entry and is incremented once per call. The learning
 * rate is read from lr_dev (device float) when non-NULL -- a scheduler can change it under a captured CUDA graph -- else
 * `lr`. Hyper-parameters are doubles because torch derives 1 - beta and the bias corrections from Python floats before
 * rounding to fp32 once. norm_out (device float, may be NULL) receives the gradient norm before clipping. */
typedef struct {
  void* param;
  void* grad;
  void* exp_avg;
  void* exp_avg_sq;
  int64_t numel;
} dg_adam_tensor_t;
#define DG_ADAM_MAX_TENSORS_PER_LAUNCH 48
DG_API size_t dg_adam_workspace_bytes(const dg_adam_tensor_t* tensors, int n_tensors);
DG_API int dg_adam_clip_step_f32(const dg_adam_tensor_t* tensors, int n_tensors, float* step, const float* lr_dev, double lr,
                          double beta1, double beta2, double eps, double weight_decay, double max_norm, float* norm_out,
                          void* workspace, size_t workspace_bytes, dg_stream_t stream);

/* Many device-to-device copies in one launch (per DG_COPY_MAX_PER_LAUNCH items): `items` is a HOST array of (dst, src, bytes)
 * triples, copied into the kernel arguments; one CTA per 16 KB chunk of one item, 16-byte accesses when both pointers are
 * aligned. Ranges must not overlap. Used to refresh the staged augmentation of a captured iteration (~60 tensors). */
typedef struct {
  void* dst;
  const void* src;
  int64_t bytes;
} dg_copy_t;
#define DG_COPY_MAX_PER_LAUNCH 96
DG_API int dg_multi_copy(const dg_copy_t* items, int n, dg_stream_t stream);

/* ---- measurement support ------------------------------------------------------------------- */
/* Read-bandwidth microbenchmark with the SpMM's access shape (scripts/l2_peak.py -> profiles/l2_peak.json): every warp
 * of a 148 * ctas_per_sm CTA grid reads `rows_per_warp` rows of `row_floats` fp32 from buf [n_rows, row_floats] with
 * 128-bit L1-bypassing loads, `rows_in_flight` (2 / 4 / 8 / 16) rows in flight; rows are consecutive across warps
 * (random = 0) or pseudo-random (random = 1). Bytes read = 148 * ctas_per_sm * 8 warps * rows_per_warp * row_floats * 4.
 * Not on the product path. */
DG_API int dg_bench_read_rows(const float* buf, int64_t n_rows, int64_t row_floats, int64_t rows_per_warp, int random,
                       int ctas_per_sm, int rows_in_flight, float* sink, dg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DREAMGNN_H_ */
