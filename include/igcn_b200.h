/*
 * igcn_b200.h -- C ABI of the B200 (sm_100a) hot-path library for INMO / IGCN collaborative
 * filtering.  One shared object (libigcn_b200.so), plain pointers and sizes, no torch types.
 *
 * The reference (WuYunfan/igcn_cf) is pure Python and has no FFI of its own; every entry point
 * below replaces a LIBRARY CALL SITE of the reference's hot path (cited per function, paths
 * relative to the reference root).  The Python host code in igcn_cf_b200/ binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - the caller owns all buffers; nothing is retained across calls;
 *  - return 0 on success, non-zero on failure (cudaError_t value, or a negative code for
 *    argument errors); igcn_last_error() returns a thread-local message;
 *  - node rows: users 0..U-1, items U..U+I-1 (reference utils.py:41-49);
 *  - embeddings are row-major fp32 [rows, D], D % 4 == 0, 16-byte aligned.
 */
#ifndef IGCN_B200_H
#define IGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGCN_ABI_VERSION 2
#define IGCN_MAX_ADD 8
#define IGCN_MAX_PEERS 8
#define IGCN_MEDIUM_NNZ 64

int igcn_abi_version(void);
const char *igcn_last_error(void);

/* CSR adjacency of one contiguous block of node rows, plus the long-row split plan that keeps
 * power-law rows from serialising on one half-warp.  Rows with more than `long_threshold`
 * non-zeros are cut into chunks (chunk_* arrays, one entry per chunk); each chunk's partial sum
 * goes to partial[chunk][D]; chunks are grouped by 32: the last chunk of a group to finish (counters[],
 * self-resetting) adds the group's partials IN CHUNK ORDER, and for rows of more than 32 chunks the last
 * group to finish adds the group sums IN GROUP ORDER, so the result does not depend on scheduling, on
 * the GPU count or on the table width.
 * Built by igcn_cf_b200/graph.py from the same data LightGCN.generate_graph builds
 * (model.py:85-94): val[e] = fl32(d_r^-1/2 * d_c^-1/2), columns sorted inside each row. */
typedef struct igcn_csr {
    int64_t n_rows;             /* rows in this block                                  */
    int64_t n_cols;             /* width of the gathered table (global node count)     */
    int64_t nnz;
    const int64_t *rowptr;      /* [n_rows + 1]                                        */
    const int32_t *col;         /* [nnz] global column ids                             */
    const float *val;           /* [nnz] or NULL (pattern-only: every value is 1)      */
    int32_t long_threshold;     /* rows with nnz > threshold are chunked               */
    int32_t n_chunks;
    const int32_t *chunk_row;   /* [n_chunks] local row id                             */
    const int64_t *chunk_begin; /* [n_chunks] first nnz of the chunk                   */
    const int32_t *chunk_len;   /* [n_chunks]                                          */
    const int32_t *chunk_first; /* [n_chunks] index of the row's first chunk           */
    const int32_t *chunk_count; /* [n_chunks] number of chunks of the row              */
    float *partial;             /* [n_chunks, D] scratch                               */
    int32_t *counters;          /* [2 * n_chunks] scratch, must be zero before first use */
    const int32_t *row_order;   /* [n_rows] visiting order (degree-descending) or NULL */
    int32_t n_long_rows;        /* leading entries of row_order with nnz > long_threshold          */
    int32_t n_medium_rows;      /* following entries with IGCN_MEDIUM_NNZ < nnz <= long_threshold:  */
                                /* one warp per row; both counts are ignored when row_order is NULL */
} igcn_csr;

/* Edge-dropout description for the INMO layer (reference NGCF.dropout_sp_mat, model.py:263-275,
 * called from IGCN.get_rep, model.py:435).  mode 0: keep everything (eval mode / p == 0);
 * mode 1: keep edge (row r, column-node c) iff hash32(seed, r, c) >= p * 2^32 (production:
 * counter-based, so the transposed backward pass regenerates the same mask); mode 2: explicit
 * keep bits (parity tests replay the reference's torch.rand draw): bit e of edge_keep for the
 * e-th non-zero in CSR order, bit r of self_keep for row r's global-template entry.
 * tperm[e] = CSR position of the reverse edge; only needed for mode 2 in the backward pass. */
typedef struct igcn_dropout {
    int32_t mode;
    float p;
    uint64_t seed;
    const uint64_t *seed_dev;   /* optional device counter mixed into seed (CUDA-graph replay)  */
    const uint32_t *edge_keep;
    const uint32_t *self_keep;
    const int64_t *tperm;
} igcn_dropout;

typedef struct igcn_step_state {
    uint64_t step;
    float adam_step_size;       /* lr / (1 - beta1^step)        */
    float adam_inv_sqrt_bc2;    /* 1 / sqrt(1 - beta2^step)     */
} igcn_step_state;

/* Y[r] = alpha * rowscale[r] * ( sum_e val[e] * X[col[e]] + sum_j add[j][r] )
 * Replaces dgl.ops.gspmm(g,'mul','sum',X,A.values()) at model.py:102 and model.py:442 (forward)
 * and its autograd backward (A is symmetric, so dX = A * dY uses the same CSR); with add[] =
 * the earlier layers and alpha = 1/(L+1) the last call also produces the layer mean that
 * torch.stack(...).mean(0) computes at model.py:104-105 / 444-445.
 * X: [n_cols, D]; Y and add[j]: [n_rows, D] (already offset to this row block); rowscale may be
 * NULL; add_host is a HOST array of n_add (<= IGCN_MAX_ADD) device pointers.
 *
 * Multi-GPU (rows sharded over one NVSwitch box): pass n_peers > 0 and peer_y_host[p] = the address of
 * the SAME output buffer on rank p (self included; peer-mapped symmetric memory).  Every finished row
 * is then stored into all n_peers copies from the kernel epilogue -- the all-gather that would follow
 * is fused into the SpMM as NVLink peer stores; the caller only needs a cross-rank barrier before
 * the next layer.  igcn_inmo_fwd / igcn_inmo_bwd take the same two arguments. */
int igcn_spmm(const igcn_csr *g, const float *X, float *Y, int32_t D,
              const float *const *add_host, int32_t n_add,
              const float *rowscale, float alpha,
              float *const *peer_y_host, int32_t n_peers, void *stream);

/* The two places where a training step does not need a full layer (same arguments as igcn_spmm plus
 * the selection; results for the selected rows are bit-identical to igcn_spmm's):
 *   igcn_spmm_rows  computes only the rows listed in row_list[0 .. *n_list) (GLOBAL row ids, ascending,
 *                   device-side count, at most max_list): the LAST forward layer of a training step,
 *                   whose output bpr_forward reads at the <= 3B rows of the batch only
 *                   (model.py:114-115 / 295-296).  Rows outside this block are skipped.
 *   igcn_spmm_cols  skips every non-zero whose column is not set in col_bits (bitmap over the n_cols
 *                   columns): the FIRST backward layer, whose input d_rep/(L+1) is non-zero on the
 *                   touched rows only -- the skipped terms are exact zeros. */
int igcn_spmm_rows(const igcn_csr *g, const float *X, float *Y, int32_t D,
                   const float *const *add_host, int32_t n_add,
                   const float *rowscale, float alpha,
                   const int64_t *row_list, const int32_t *n_list, int64_t max_list, int64_t row0,
                   float *const *peer_y_host, int32_t n_peers, void *stream);
int igcn_spmm_cols(const igcn_csr *g, const float *X, float *Y, int32_t D,
                   const float *const *add_host, int32_t n_add,
                   const float *rowscale, float alpha, const uint32_t *col_bits,
                   float *const *peer_y_host, int32_t n_peers, void *stream);

/* INMO template aggregation fused with edge dropout (IGCN.inductive_rep_layer, model.py:423-432,
 * after dropout_sp_mat, model.py:435):
 *   X0[r] = rowscale[r]/(1-p) * ( sum_{c in adj(r), tmpl[c] >= 0, keep(r,c)} E[tmpl[c]]
 *                                 + keep_self(r) * E[r < n_users ? glob_user : glob_item] )
 * where rowscale[r] = row_sum[r]^((alpha-1)/2 - 1/2) (model.py:374-377).  tmpl == NULL means the
 * identity map (feature_ratio == 1: template id == node id).  row0 = global id of local row 0. */
int igcn_inmo_fwd(const igcn_csr *g, const int32_t *tmpl, const float *rowscale,
                  const igcn_dropout *drop, const float *E, float *X0, int32_t D,
                  int64_t row0, int64_t n_users, int64_t glob_user, int64_t glob_item,
                  float *const *peer_x0_host, int32_t n_peers, void *stream);

/* Transposed INMO layer for the embedding gradient (autograd backward of model.py:430):
 *   dE[tmpl[c]] = sum_{r in adj(c), keep(r,c)} G[r]      for local rows c with tmpl[c] >= 0
 * G must already hold rowscale[r]/(1-p) * dX0[r] (igcn_spmm's rowscale/alpha epilogue does that).
 * The two global-template rows are produced by igcn_colsum_masked. */
int igcn_inmo_bwd(const igcn_csr *g, const int32_t *tmpl, const igcn_dropout *drop,
                  const float *G, float *dE, int32_t D, int64_t row0,
                  float *const *peer_de_host, int32_t n_peers, void *stream);

/* out[d] = sum over rows r in [row_begin,row_end) with keep_self(r) of G[r][d]; fixed two-stage
 * order (deterministic).  scratch: [ceil(n/256)+1, D] floats. */
int igcn_colsum_masked(const float *G, int64_t row_begin, int64_t row_end, int32_t D,
                       const igcn_dropout *drop, float *scratch, float *out, void *stream);

/* Triple sampler (BasicDataset.__getitem__, dataset.py:119-131, neg_ratio 1): user uniform over
 * users with a non-empty train list, positive uniform over the user's list, negative uniform over
 * items rejecting the user's list.  The user-by-item train CSR is passed as rows [0,n_users) of
 * `rowptr/col` with item = col - col_offset (the adjacency's user rows, col_offset = n_users).
 * out: int64 [B, 3].  Counter-based: (seed, step [+ *step_dev], triple index) fully determine the
 * draw; step_dev (may be NULL) is a device counter so a captured CUDA graph advances by itself. */
int igcn_sample_triples(const int64_t *rowptr, const int32_t *col, int64_t col_offset,
                        int64_t n_users, int64_t n_items, int64_t B, uint64_t seed,
                        uint64_t step, const uint64_t *step_dev, int64_t *out, void *stream);

/* Forward of the BPR step on a batch of triples (trainer.py:238-241 / 300-302 and, with w, the
 * auxiliary loss at trainer.py:304-311; L2 term of model.py:110-113 / 297-298):
 *   pos_i = <T[u_i] * w, T[off+p_i]>, neg_i = <T[u_i] * w, T[off+n_i]>
 *   sp[i] = softplus(neg_i - pos_i), sig[i] = sigmoid(neg_i - pos_i)
 *   l2[i] = |L[u_i]|^2 + |L[off+p_i]|^2 + |L[off+n_i]|^2      (L = l2_table, may be NULL)
 * triples: int64 [B,3]; w: [D] or NULL. */
int igcn_bpr_fwd(const float *table, const float *l2_table, const float *w,
                 const int64_t *triples, int64_t B, int64_t item_offset, int32_t D,
                 float *sp, float *sig, float *l2, void *stream);

/* The same forward when the COLUMNS of the tables are sharded over n_peers = 2, 4 or 8 ranks (multi-GPU training
 * without any exchange of layer embeddings: the propagation is linear and acts on every column independently, so a
 * rank that owns D / n_peers columns of the parameters propagates just those; the only quantities that couple the
 * columns are the dot products and squared norms of this step).  table / l2_table / w hold this rank's D columns.
 *   igcn_bpr_partial   this rank's partial sums of every triple, stored into record i of EVERY rank's exchange buffer
 *                      parts_peer_host[p] (float [2 parities][n_peers][cap][8]; slots slot0.. = pos, neg and, with
 *                      l2_table, the three squared norms; the parity is igcn_step_state.step & 1)
 *   igcn_bpr_combine   after a device barrier: adds the n_peers partials in the order of the single-GPU kernel's
 *                      lane tree (bit-identical to igcn_bpr_fwd on the full-width tables) and finishes sp / sig / l2. */
int igcn_bpr_partial(const float *table, const float *l2_table, const float *w, const int64_t *triples,
                     int64_t B, int64_t item_offset, int32_t D, float *const *parts_peer_host,
                     int32_t n_peers, int32_t rank, int32_t slot0, int64_t cap,
                     const igcn_step_state *state_dev, void *stream);
int igcn_bpr_combine(const float *parts, int64_t B, int64_t cap, int32_t n_peers, int32_t slot0,
                     int32_t has_l2, const igcn_step_state *state_dev, float *sp, float *sig, float *l2,
                     void *stream);

/* loss[0] = mean(sp) + l2_reg * mean(l2) + aux_reg * mean(aux_sp)   (trainer.py:241-243, 313-314)
 * acc[0] += loss * B ; acc[1] += B   (the AverageMeter of utils.py:126-135, kept on device so
 * the per-step loss.item() sync of trainer.py:247/318 disappears).  l2 / aux_sp may be NULL. */
int igcn_loss_finalize(const float *sp, const float *l2, const float *aux_sp, int64_t B,
                       int64_t B_aux, float l2_reg, float aux_reg, float *loss, double *acc,
                       void *stream);

/* Deterministic scatter plan for the gradient of a triple batch: slot s = kind * B + i
 * (kind 0 user, 1 positive, 2 negative) touches row id(s).  Sorts (id, slot) and emits
 * order[3B] (slots, ascending id then slot), seg_start[n_seg+1], seg_row[n_seg] (ascending), n_seg[0].
 * n_rows = exclusive upper bound of the row ids.  touched_bits (may be NULL): bitmap of n_rows bits,
 * cleared and then set for every touched row -- igcn_spmm_cols uses it in the backward pass.
 * One CTA bitonic sort (registers + shuffles + shared memory); 3B <= 16384. */
int igcn_bpr_plan(const int64_t *triples, int64_t B, int64_t item_offset, int64_t n_rows,
                  int32_t *order, int32_t *seg_start, int64_t *seg_row, int32_t *n_seg,
                  uint32_t *touched_bits, void *stream);

/* Gradient rows of the BPR step, one half-warp per touched row, contributions added in slot
 * order (no atomics; replaces the index_put_(accumulate=True) autograd backward of
 * model.py:114-115 / 295-296):
 *   d/dT[u_i]   += c_i * w*(T[n_i] - T[p_i]) + lam * L2row
 *   d/dT[p_i]   += -c_i * w*T[u_i] + lam * L2row      d/dT[n_i] += c_i * w*T[u_i] + lam * L2row
 * with c_i = scale * sig[i] / B and lam = scale * 2 * l2_coef / B applied to the row itself (only when
 * l2_on_table != 0).  accumulate == 0: G[row] = sum (G is expected pre-zeroed elsewhere);
 * accumulate != 0: G[row] += sum.  dw (may be NULL, needs w): dw[d] += scale/B * sum_i sig_i *
 * T[u_i][d] * (T[n_i][d] - T[p_i][d]) through dw_scratch [ceil(B/64), D]. */
int igcn_bpr_bwd(const float *table, const float *w, const int64_t *triples, int64_t B,
                 int64_t item_offset, int32_t D, const float *sig, float scale, float l2_coef,
                 int32_t l2_on_table, const int32_t *order, const int32_t *seg_start,
                 const int64_t *seg_row, const int32_t *n_seg, float *G, int32_t accumulate,
                 float *dw, float *dw_scratch, void *stream);

/* dw += scale / B * sum_i sig_i * T[u_i] .* (T[off+n_i] - T[off+p_i]): the gradient of the auxiliary loss with respect
 * to the weight vector w (trainer.py:304-311) on its own -- the same two kernels igcn_bpr_bwd runs when it is given dw,
 * so that the host can put them on a side stream (they depend on sig only, not on the backward propagation).
 * Fixed summation order: 64 triples per block in order, block partials in order. */
int igcn_bpr_dw(const float *table, const int64_t *triples, int64_t B, int64_t item_offset, int32_t D,
                const float *sig, float scale, float *dw, float *dw_scratch, void *stream);

/* dE[row] += coef * multiplicity(row) * E[row] for every touched row: gradient of
 * l2_reg * mean(|E[u]|^2 + |E[p]|^2 + |E[n]|^2) over RAW embedding rows (LightGCN.bpr_forward,
 * model.py:110-113), coef = 2 * l2_reg / B. */
int igcn_l2_rows_bwd(const float *E, float *dE, int32_t D, float coef, const int32_t *seg_start,
                     const int64_t *seg_row, const int32_t *n_seg, int64_t max_seg, void *stream);

/* torch.optim.Adam (trainer.py:43-45, 246) with default betas/eps semantics, one fused pass:
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) */
int igcn_adam(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1,
              float beta2, float eps, int64_t t, const igcn_step_state *state_dev, void *stream);

/* Device-resident step counter so that a whole training step can be replayed as one CUDA graph:
 * state->step += 1 and the Adam bias-correction factors for t = step are refreshed.  When
 * igcn_adam gets state_dev != NULL it reads the factors from there instead of using `t`. */
int igcn_step_tick(igcn_step_state *state_dev, float lr, float beta1, float beta2, void *stream);

/* Full-ranking scoring fused with the seen-item mask and per-user top-k; the score matrix never
 * reaches HBM.  Replaces torch.mm (model.py:122), the -inf index_put (trainer.py:149-161) and
 * torch.topk (trainer.py:163).
 *   score(u, j) = sum_{d ascending} rep[user_ids[b]][d] * rep[item_row0 + j][d]   (fp32 FMA chain)
 * masked when j is in mask_items[mask_ptr[u] .. mask_ptr[u+1]) (sorted ascending per user, may be
 * NULL), or j outside [item_lo, item_hi), or bit j of banned_bits set (may be NULL).
 * Output per user: k item ids (int32, -1 when fewer than k candidates) and scores, sorted by
 * (score descending, item ascending).  This is the exact CUDA-core kernel; it is also the
 * fallback the tensor-core path uses for users whose candidate bound does not verify: out_rows
 * (may be NULL) gives the output row of entry b, n_eval_dev (may be NULL) a device-side count.
 * n_item_splits > 1 (with split_keys [split_cap, n_item_splits, k] uint64 scratch): while the entry count is
 * <= split_cap the catalogue is cut into n_item_splits ranges scanned by separate CTAs and merged by a second
 * kernel -- a fallback list of a few users then costs a fraction of a millisecond instead of one CTA's serial
 * pass over all items; larger counts run unsplit.  Pass 1 / NULL / 0 for the plain behaviour. */
int igcn_score_topk_exact(const float *rep, const int64_t *user_ids, int64_t n_eval,
                          int64_t item_row0, int64_t n_items, int32_t D,
                          const int64_t *mask_ptr, const int32_t *mask_items,
                          int64_t item_lo, int64_t item_hi, const uint32_t *banned_bits,
                          int32_t k, int32_t *out_items, float *out_scores,
                          const int32_t *out_rows, const int32_t *n_eval_dev,
                          int32_t n_item_splits, uint64_t *split_keys, int64_t split_cap, void *stream);

/* ---- tensor-core scoring (tcgen05 / TMEM / bulk TMA), D <= 64, k <= 24 ------------------------
 * Same contract as igcn_score_topk_exact, split in three launches the host chains on one stream:
 *   igcn_tc_pack        fp32 rep rows -> fp16 operand images in the UMMA core-matrix layout, with one
 *                       extra K block carrying the rounding-error bound (c*|u| for users, |i| for items)
 *   igcn_tc_candidates  tcgen05.mma M128 x N256 tiles, accumulators in TMEM; the epilogue filters
 *                       s_hat > running threshold and !masked into <= 96 candidates per (user, split)
 *                       (a 3-input-max reduction and one warp vote per 32-column chunk; only chunks with a
 *                       score above some row's threshold are compared column by column)
 *   igcn_tc_finalize    exact fp32 re-scoring (same FMA order as the exact kernel), top-k, and the
 *                       proof check; users that fail it are appended to (fb_users, fb_rows, fb_count)
 *                       for igcn_score_topk_exact.
 * Item rows are packed relative to the mean item row (center_sum[D] = column sums of the item rows, written
 * by igcn_tc_pack through center_scratch [ceil(n_items/256)+1, D]; igcn_tc_finalize reads it back): a per-user
 * constant does not change the ranking, and the bound then scales with |i - mean| instead of |i|.
 * a_img / b_img must be zero-initialised by the caller (padding rows); sizes from igcn_tc_workspace.
 * mask_tile_ptr [ceil(n_eval/128), ceil(n_items/256)+1] + mask_entries ((row<<8)|col, uint16) is the
 * seen-item CSR bucketed by (user tile, item tile); dump (tests only) receives every s_hat.
 * n_head: the first n_head user tiles scan all item tiles in one CTA and use only list slot 0; the remaining
 * tiles are split n_splits ways -- the host sizes the split tail so that it fills the last wave of SMs
 * (0 = every user tile is split).
 * item_perm (int32 [n_items], NULL = identity): the SCAN ORDER of the items -- position p of the item image is
 * item item_perm[p].  Everything between pack and finalize is in position space (mask buckets, banned bitmap,
 * item_lo / item_hi, candidate lists); igcn_tc_finalize maps positions back to item ids.  The order does not
 * change the result (the final lists are proven exact or recomputed), it changes the cost: with the items most
 * likely to rank high first (the host passes train popularity), the running thresholds tighten within the first
 * tiles and the epilogue's compare-free path handles almost every later chunk.
 * stats (uint64 [5], NULL = off; selects an instrumented copy of the kernel): 32-column chunks seen, chunks that
 * left the compare-free path, 8-column groups compared, candidates appended, compactions. */
int igcn_tc_workspace(int64_t n_eval, int64_t n_items, int32_t D, int32_t n_splits,
                      int64_t *a_img_bytes, int64_t *b_img_bytes, int64_t *cand_slots);
int igcn_tc_pack(const float *rep, int64_t n_rep_elems, const int64_t *user_ids, int64_t n_eval,
                 int64_t item_row0, int64_t n_items, int32_t D, const int32_t *item_perm,
                 uint32_t *maxabs_bits, uint8_t *a_img, uint8_t *b_img, float *center_sum,
                 float *center_scratch, void *stream);
int igcn_tc_candidates(const uint8_t *a_img, const uint8_t *b_img, int64_t n_eval, int64_t n_items,
                       int32_t D, int32_t n_splits, int32_t n_head, int64_t item_lo,
                       int64_t item_hi,
                       const uint32_t *banned_bits, const int32_t *mask_tile_ptr,
                       const uint16_t *mask_entries, int32_t *cand_items, int32_t *cand_cnt,
                       float *cand_thr, float *dump, unsigned long long *stats, void *stream);
int igcn_tc_finalize(const float *rep, const int64_t *user_ids, int64_t n_eval, int64_t item_row0,
                     int32_t D, int32_t n_splits, const int32_t *cand_items,
                     const int32_t *cand_cnt, const float *cand_thr, const uint32_t *maxabs_bits,
                     const float *center_sum, int64_t n_items, const int32_t *item_perm,
                     int32_t k, int32_t *out_items, float *out_scores, int32_t *fb_count,
                     int64_t *fb_users, int32_t *fb_rows, void *stream);

/* ---- peer memory for the row-sharded multi-GPU path (one process per GPU, one NVSwitch box) ------
 * The reference has no distributed code at all; these entry points exist so that the
 * all-gather after each propagation layer can be fused into the SpMM as NVLink peer stores.
 *   igcn_peer_alloc    cudaMalloc + zero-fill `bytes`, return the pointer and its 64-byte IPC handle
 *   igcn_peer_open     map another rank's allocation from its handle (peer access enabled lazily)
 *   igcn_peer_close    unmap;    igcn_peer_free   release an igcn_peer_alloc allocation
 *   igcn_peer_barrier  device-side barrier on `stream` across n_peers ranks: flags_host[p] is rank p's
 *                      flag array (uint32 [IGCN_MAX_PEERS], peer-mapped, zero-initialised), epoch_dev a
 *                      local counter, status_dev a local word set to 1 if a peer did not arrive within
 *                      the timeout (120 s, IGCN_PEER_TIMEOUT_S overrides) -- the kernel then TRAPS: the context
 *                      dies and every later CUDA call of the process fails, instead of a hung GPU or layers that
 *                      silently read rows which never arrived.  CUDA-graph capturable.
 * Buffers from igcn_peer_alloc are owned by the library until igcn_peer_free; everything else in this
 * header operates on caller-owned memory. */
#define IGCN_PEER_HANDLE_BYTES 64
int igcn_peer_alloc(int64_t bytes, void **ptr_out, uint8_t *handle_out);
int igcn_peer_open(const uint8_t *handle, void **ptr_out);
int igcn_peer_close(void *ptr);
int igcn_peer_free(void *ptr);
int igcn_peer_barrier(uint32_t *const *flags_host, int32_t n_peers, int32_t rank,
                      uint32_t *epoch_dev, uint32_t *status_dev, void *stream);
/* Bulk alternative to the in-kernel peer stores for LARGE row blocks (hundreds of MB per layer, where
 * scattered 128-byte NVLink writes from the SpMM epilogue run far below link speed): after the layer kernel
 * has written this rank's rows into its own copy, stream elements [elem_offset, +n_elems) of that buffer
 * (fp32, multiples of 4) to the same place in every other rank's copy.  peer_host as in igcn_spmm. */
int igcn_peer_push(float *const *peer_host, int32_t n_peers, int32_t rank, int64_t elem_offset,
                   int64_t n_elems, void *stream);
/* Column-sharded training: write this rank's [rows, ds] slice into columns [col0, col0 + ds) of EVERY rank's
 * [rows, d] copy (its own included) -- the all-gather of the parameters before an evaluation or a checkpoint. */
int igcn_peer_push_cols(float *const *peer_host, int32_t n_peers, const float *src, int64_t rows,
                        int32_t ds, int32_t d, int32_t col0, void *stream);

/* out[b][j] = <rep[user_ids[b]], rep[item_row0 + j]> (fp32 FMA chain, ascending d): the dense score block of
 * LightGCN.predict (model.py:118-123, torch.mm at :122) for callers that want raw scores; n_eval <= 65535 per
 * call.  The evaluation path does not use it (scores never reach HBM there). */
int igcn_predict_scores(const float *rep, const int64_t *user_ids, int64_t n_eval, int64_t item_row0,
                        int64_t n_items, int32_t D, float *out, void *stream);

/* hit[u][j] = 1 if rec[u][j] is in eval_items[eval_ptr[u] .. eval_ptr[u+1]) (sorted), else 0:
 * the membership double loop of BasicTrainer.calculate_metrics (trainer.py:111-115). */
int igcn_hits(const int32_t *rec, int64_t n_users, int32_t k, const int64_t *eval_ptr,
              const int32_t *eval_items, float *hit, void *stream);

/* The per-user part of BasicTrainer.calculate_metrics (trainer.py:109-131) on the device: for every cut-off
 * topks_host[t] <= k, hit_num[t][u] = number of u's first topks[t] recommendations found in its eval list and
 * dcg[t][u] = sum_j hit_j / log2_table[j] -- log2_table is the reference's fp32 np.log2(arange(2, k + 2)) computed
 * by the host, the division is IEEE and the sum follows numpy's pairwise float32 row reduction, so both arrays are
 * bit-identical to what the reference's numpy expressions produce from the [U, k] hit matrix.  Only these
 * 2 x n_topks x U floats have to leave the device; the host finishes with the reference's own vector expressions. */
int igcn_user_metrics(const int32_t *rec, int64_t n_users, int32_t k, const int64_t *eval_ptr,
                      const int32_t *eval_items, const float *log2_table, const int32_t *topks_host,
                      int32_t n_topks, float *hit_num, float *dcg, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* IGCN_B200_H */
