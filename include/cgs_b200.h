/*
 * cgs_b200.h -- C ABI of the B200-native collapsed-Gibbs LDA hot path.
 *
 * pycisTopic has no C FFI on this path: run_cgs_model drives the third-party `lda`
 * package through its Python object protocol (reference src/pycisTopic/lda_models.py:248-287)
 * and, beneath that, the Cython entry points `lda._lda._sample_topics(WS, DS, ZS, nzw, ndz,
 * nz, alpha, eta, rands)` and `lda._lda._loglikelihood(nzw, ndz, nz, nd, alpha, eta)`
 * (SURVEY.md section 8b.2).  This header is the device-resident equivalent a maintainer
 * would bind instead (ctypes stub in INTEGRATION.md).  Plain pointers and sizes only.
 *
 * Conventions: every function returns 0 on success, non-zero on failure; the message is
 * available from cgs_last_error() (thread-local).  The caller owns every host buffer; a
 * handle owns its device state.  One corpus is bound to one CUDA device; any number of
 * models (one per n_topics value) may share a corpus.  Functions are asynchronous on the
 * model's stream unless they return data to the host.
 *
 * Layouts handed back to the host are the REFERENCE's: nzw (K x W), ndz (D x K), nz (K),
 * int32 row-major; topic_word (K x W) and doc_topic (D x K) float64 (lda_models.py:292-293,
 * :345-357).  On the device n_wk is kept region-major (W x Kp, Kp = K rounded up to 4) so
 * that the sampler reads one contiguous row per token.
 */
#ifndef CGS_B200_H
#define CGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgs_corpus cgs_corpus;
typedef struct cgs_model cgs_model;

/* ---- library ------------------------------------------------------------------------ */
const char *cgs_last_error(void);
int cgs_version(void);
int cgs_device_count(int *count);
int cgs_max_topics(void);

/* ---- corpus: token-list construction on device (replaces lda_models.py:151 +
 *      lda.utils.matrix_to_lists) -------------------------------------- */

/* Build from the reference's CistopicObject.binary_matrix layout: CSR with n_regions rows
 * and n_cells columns (host pointers; indptr int64[n_regions+1], indices int32[nnz] = cell
 * ids).  The transpose to cell-major token order (cell ascending, region ascending) runs on
 * the device.  Only cells in [cell_begin, cell_end) are kept (AD-LDA shard; pass 0, n_cells
 * for everything).  Duplicate (region, cell) entries collapse to one token (binary data). */
int cgs_corpus_from_regions_by_cells(cgs_corpus **out, int device,
                                     const int64_t *indptr, const int32_t *indices,
                                     int32_t n_regions, int32_t n_cells, int64_t nnz,
                                     int32_t cell_begin, int32_t cell_end);

/* Build from the layout run_cgs_model receives (lda_models.py:179): CSR with n_cells rows and
 * n_regions columns; indices need not be sorted (they are canonicalised on the device). */
int cgs_corpus_from_cells_by_regions(cgs_corpus **out, int device,
                                     const int64_t *indptr, const int32_t *indices,
                                     int32_t n_cells, int32_t n_regions, int64_t nnz,
                                     int32_t cell_begin, int32_t cell_end);

/* Same as cgs_corpus_from_cells_by_regions but the two arrays are DEVICE pointers on
 * `device` (already canonical: sorted, unique); used by the benchmark for resident inputs. */
int cgs_corpus_from_device_tokens(cgs_corpus **out, int device,
                                  const int64_t *d_cell_ptr, const int32_t *d_tok_region,
                                  int32_t n_cells, int32_t n_regions, int64_t n_tokens,
                                  int64_t global_token_offset);

int cgs_corpus_destroy(cgs_corpus *c);
int cgs_corpus_dims(const cgs_corpus *c, int32_t *n_cells, int32_t *n_regions, int64_t *n_tokens);
/* Token lists in the `lda` package's form: WS[i] = region id, DS[i] = cell id (local to the
 * shard), int32[n_tokens] each; either pointer may be NULL. */
int cgs_corpus_export_tokens(const cgs_corpus *c, int32_t *WS, int32_t *DS);
/* Global index of this shard's first token (z initialisation is i mod K over GLOBAL i). */
int cgs_corpus_set_token_offset(cgs_corpus *c, int64_t global_token_offset);
/* Size of the WHOLE matrix when this corpus holds one shard of its cells (AD-LDA): models created afterwards
 * size their concurrency caps (cells open at once, tokens in flight) against the matrix the shard is sampled
 * against, not against the shard -- otherwise a 1/8 shard would fill less than half of its GPU. */
int cgs_corpus_set_global_size(cgs_corpus *c, int64_t n_cells_global, int64_t n_tokens_global);

/* ---- model: one lda.LDA instance (lda_models.py:248-287) ----------------------------- */

/* alpha, eta are the values the sampler uses (already divided by K where the caller asked
 * for *_by_topic, lda_models.py:247-282).  Fails for alpha <= 0 or eta <= 0 (lda raises
 * ValueError).  `stream` is a cudaStream_t to launch on, or NULL for a private stream. */
int cgs_model_create(cgs_model **out, cgs_corpus *corpus, int32_t n_topics,
                     double alpha, double eta, uint64_t seed, void *stream);
int cgs_model_destroy(cgs_model *m);
int cgs_model_set_stream(cgs_model *m, void *stream);
int cgs_model_sync(cgs_model *m);

/* lda.LDA._initialize: z_i = i mod K, rebuild n_wk / n_dk / n_k (bit-exact). */
int cgs_model_init(cgs_model *m);
/* Replace the assignments by a host vector (int32[n_tokens], values in [0, K)) and rebuild
 * the three count tables from it (bit-exact bookkeeping tests). */
int cgs_model_set_z(cgs_model *m, const int32_t *z);
int cgs_model_get_z(cgs_model *m, int32_t *z);

/* n_sweeps collapsed-Gibbs sweeps over every token of the corpus (replaces
 * lda._lda._sample_topics).  Asynchronous. */
int cgs_model_sweep(cgs_model *m, int32_t n_sweeps);
/* One PART of a sweep: the cells at queue positions part, part + n_parts, ... (the queue is
 * longest-first, so the parts are token-balanced).  Calling it for part = 0 .. n_parts-1 is one
 * sweep; AD-LDA uses it to exchange n_wk deltas several times per sweep. */
int cgs_model_sweep_part(cgs_model *m, int32_t part, int32_t n_parts);
/* Diagnostics: runs ONE sweep with a device counter of the tokens actually sampled (must equal the
 * token count: every token is resampled exactly once per sweep, as in the sequential loop). */
int cgs_model_debug_sweep_visits(cgs_model *model, int64_t *tokens_sampled);
/* Tuning knobs of the sweep kernel: lanes that cooperate on one token (power of two, must
 * cover n_topics with at most 32 topics per lane) and warps that share one cell (1, 2, 4, 8, 12 or 16).
 * The defaults are chosen from n_topics and the mean cell size at cgs_model_create. */
int cgs_model_set_sweep_shape(cgs_model *m, int32_t lanes_per_token, int32_t warps_per_cell);
/* Upper bound on the CTAs of the sweep kernel (0 = fill the GPU).  With 1 CTA, 32 lanes per
 * token and 1 warp per cell the cells are sampled one after the other, one token at a time (the
 * concurrency studies of scripts/trace_check.py use it as their sequential end). */
/* Number of distinct uniform draws per sweep (power of two >= 128; 0 = one independent draw per token).
 * Default 131072: lda reuses rands[i % 131072] inside a sweep and that reuse measurably
 * speeds up the transient of the chain, so the device sampler keeps the same reuse density. */
int cgs_model_set_rng_period(cgs_model *m, int64_t period);
int cgs_model_set_max_blocks(cgs_model *m, int32_t max_blocks);
/* The topic totals n_k that a warp works with are re-synchronised often enough that at most
 * 1/divisor of a sweep is processed GPU-wide in between (default 16). */
int cgs_model_set_nk_staleness(cgs_model *m, double divisor);
int cgs_model_get_sweep_shape(const cgs_model *m, int32_t *lanes_per_token, int32_t *chunks_per_lane,
                              int32_t *warps_per_cell);
/* Deterministic mode (the reference is bit-reproducible for a fixed random_state, lda_models.py:104; the default
 * device schedule is not, because cells read the LIVE n_wk while other cells update it).  parts > 0: every sweep runs
 * as `parts` sub-sweeps; each reads n_wk / n_k from a copy taken when the sub-sweep starts, the warps of a cell work in
 * rounds on private copies of the cell's row that are merged with integer atomics between two barriers, and all count
 * updates are integer reductions (commutative) -- so the result is a pure function of (matrix, n_topics, alpha, eta,
 * seed, parts, sweep shape), whatever the scheduling.  The price is staleness of up to 1/parts of a sweep (DESIGN.md
 * section 5).  parts = 0 restores the default schedule. */
int cgs_model_set_deterministic(cgs_model *m, int32_t parts);
/* Arithmetic of the conditional p(k), its running sum and the inversion: 32 = float (default), 64 = double, the
 * reference's precision (lda._lda._sample_topics).  Same schedule, random bits and integer bookkeeping
 * either way: the A/B behind the statement that fp32 does not move the log-likelihood trace. */
int cgs_model_set_conditional_precision(cgs_model *m, int32_t bits);
/* Kernel launches issued by this model so far (for bench.py's gpu_launches). */
int cgs_model_launch_count(const cgs_model *m, int64_t *count);
/* Number of sweeps done so far. */
int cgs_model_sweeps_done(const cgs_model *m, int64_t *count);

/* lda._lda._loglikelihood with the model's alpha, eta.  For an AD-LDA shard
 * `partial` selects what is summed: 0 = everything, 1 = only the cell-side terms (n_dk, nd),
 * 2 = only the region-side terms (n_wk, n_k). */
int cgs_model_loglik(cgs_model *m, int partial, double *ll);

/* ---- consistency checks on the device state (bench.py runs them after its timed regions) ---- */
/* d_hist (DEVICE, int32[n_regions * Kp], overwritten) receives the (region, topic) histogram of this corpus' z;
 * *ndk_mismatches = entries of n_dk that differ from the per-cell topic histogram of z (+ topics out of range). */
int cgs_model_region_histogram(cgs_model *m, int32_t *d_hist, int64_t *ndk_mismatches);
/* out3[0] = n_dk mismatches (as above; only when d_hist == NULL), out3[1] = entries where n_wk differs from d_hist
 * (NULL: the local histogram of z is built first; an AD-LDA shard passes the histogram summed over all ranks),
 * out3[2] = topics whose n_k differs from the column sum of n_wk.  All zero = the three tables are exactly the
 * bookkeeping of z (the invariant lda._lda._sample_topics keeps). */
int cgs_model_verify_counts(cgs_model *m, const int32_t *d_hist, int64_t *out3);
/* out4 = { position-dependent 64-bit hash of n_wk, sum(n_wk), sum(n_k), hash of n_k }: replicas of an AD-LDA run
 * must agree on all four and the sums must equal the global token count. */
int cgs_model_state_hash(cgs_model *m, uint64_t *out4);

/* Reference-layout exports (host pointers, any may be NULL). */
int cgs_model_export_counts(cgs_model *m, int32_t *nzw, int32_t *ndz, int32_t *nz);
int cgs_model_export_distributions(cgs_model *m, double *topic_word, double *doc_topic);
/* DataFrame-ready layouts: topic_region (W x K) and cell_topic (K x D) float64
 * (lda_models.py:348-357 build these by transposing; the device writes them directly). */
int cgs_model_export_frames(cgs_model *m, double *topic_region, double *cell_topic);

/* ---- model-quality metrics (replaces the tmtoolkit calls of lda_models.py:291-304, :332-344 and
 *      utils.loglikelihood, utils.py:129-160, called at lda_models.py:305) ----
 * The reference computes them on the host from topic_word_ (K x W), doc_topic_ (D x K) and the token
 * matrix.  Here the O(K^2 W), O(K D) and O(N) parts run on the device from the resident counts and
 * only K x K / K x top_n sized results come back; the host finishes with O(K^3) arithmetic. */
/* Tokens per cell (doc_lengths = X.sum(axis=1), lda_models.py:295): int64[n_cells]. */
int cgs_corpus_export_cell_sizes(const cgs_corpus *c, int64_t *sizes);
/* gram = topic_word @ topic_word.T, float64[K x K], from an EXACT int64 Gram of n_wk.
 * metric_arun_2010 needs its eigenvalues (singular values of topic_word squared), metric_cao_juan_2009
 * its cosines. */
int cgs_model_topic_gram(cgs_model *m, double *gram);
/* mixture[k] = sum_d nd_d * doc_topic[d, k] (float64[K]) and *length_sq = sum_d nd_d^2: Arun's
 * cm2 = mixture / sqrt(length_sq); marginal_topic_distrib = mixture / sum(mixture). */
int cgs_model_cell_mixture(cgs_model *m, double *mixture, double *length_sq);
/* The top_n (<= 32) regions of every topic, most probable first: int32[K x top_n] (and their counts,
 * may be NULL) -- tmtoolkit's argsort(topic_word, axis=1)[:, :-(top_n+1):-1] with ties (equal counts)
 * ordered by region index descending, i.e. a stable argsort.  Fails like tmtoolkit when top_n exceeds
 * the number of regions. */
int cgs_model_top_regions(cgs_model *m, int32_t top_n, int32_t *top_regions, int32_t *top_counts);
/* codf[t][a][b] = number of cells that contain both top_regions[t][a] and top_regions[t][b]
 * (a == b: the document frequency), int32[K x top_n x top_n], from the device token list
 * (metric_coherence_mimno_2011's df / codf, lda_models.py:297-304). */
int cgs_model_codocument_frequencies(cgs_model *m, int32_t top_n, const int32_t *top_regions, int32_t *codf);
/* The value utils.loglikelihood(nzw, ndz, alpha, eta) returns (bug-compatible: its loops overwrite
 * instead of accumulating, utils.py:145, :155), called with the RAW alpha and eta (lda_models.py:305). */
int cgs_model_loglik_compat(cgs_model *m, double alpha_raw, double eta_raw, double *ll);

/* ---- AD-LDA exchange step ------------------------------------- */
/* Device buffers a host framework may wrap (torch via __cuda_array_interface__) to run the
 * collective: 0 = n_wk (int32, W*Kp), 1 = n_wk snapshot taken at the last exchange,
 * 2 = delta buffer (int32, W*Kp), 3 = n_k (int32, Kp), 4 = n_dk (int32, D*Kp),
 * 5 = z (uint16, n_tokens), 6 = token regions (int32, n_tokens), 7 = cell_ptr (int64, D+1). */
int cgs_model_device_buffer(cgs_model *m, int which, void **ptr, size_t *bytes);
int cgs_model_padded_topics(const cgs_model *m, int32_t *kp);
/* delta = n_wk - snapshot (the local changes since the last exchange). */
int cgs_model_delta_compute(cgs_model *m);
/* n_wk = snapshot + delta (delta now holds the sum over ranks), snapshot = n_wk,
 * n_k = column sums of n_wk. */
int cgs_model_delta_apply(cgs_model *m);
/* snapshot = n_wk (call once after init / set_z). */
int cgs_model_snapshot(cgs_model *m);
/* Peer-memory exchange: every rank passes the device pointers of ALL ranks' delta buffers
 * (peer-mapped, n_ranks entries, own rank included) and ONE kernel reduces them and applies
 * the result to the local n_wk / n_k.  The caller provides the cross-rank barrier before
 * (all deltas computed) and after (all reads done). */
int cgs_model_delta_apply_peers(cgs_model *m, void *const *peer_delta, int32_t n_ranks);

/* The whole synchronous exchange step behind the ABI: delta = n_wk - snapshot, ncclAllReduce(delta, int32, sum) on
 * the model's stream with the caller's communicator (an ncclComm_t: torch's ProcessGroupNCCL._comm_ptr(), or one
 * made by cgs_nccl_comm_create), then cgs_model_delta_apply.  libnccl.so.2 is resolved at first use (dlopen): the copy
 * the process has already loaded, else the file $CGS_NCCL_LIBRARY names, else the loader's default. */
int cgs_model_allreduce_delta(cgs_model *m, void *nccl_comm);
/* For hosts without their own NCCL plumbing: rank 0 calls cgs_nccl_unique_id and ships the 128 bytes to the other
 * ranks by any means; every rank then creates its communicator. */
int cgs_nccl_unique_id(void *id_out_128_bytes);
int cgs_nccl_comm_create(void **comm, int device, int32_t n_ranks, int32_t rank, const void *id_128_bytes);
int cgs_nccl_comm_destroy(void *comm);

/* Asynchronous exchange over peer memory, ONE fused kernel per round that runs BESIDE the sweep kernel
 * (csrc/exchange_kernels.cuh): capture the local changes, reduce-scatter + all-gather them through loads / stores on
 * the peers' arenas over NVLink, apply the other ranks' changes to the live n_wk with reductions.  Counts stay
 * exact; what a rank sees of the others is at most one round old.
 *   cgs_model_exchange_arena: the rank's arena (delta | total | flags; one cudaMalloc block, exportable with
 *     cgs_ipc_export).  Call after cgs_model_init / set_z and before the first sweep.
 *   cgs_model_exchange_open: peer_arena[q] = rank q's arena mapped into this process (cgs_ipc_open), own entry
 *     ignored.  The first round turns the local counts into global ones on every rank.
 *   cgs_model_exchange_round: enqueue one round on the model's exchange stream (highest priority); it starts when
 *     the work enqueued on the model's stream so far is done and overlaps what is enqueued there next.  Every rank
 *     must enqueue the same number of rounds.  n_ctas <= 0: default (64).
 *   cgs_model_exchange_wait: the model's stream waits for the last enqueued round (before loglik / exports). */
int cgs_model_exchange_arena(cgs_model *m, void **ptr, size_t *bytes);
int cgs_model_exchange_open(cgs_model *m, void *const *peer_arena, int32_t n_ranks, int32_t rank);
int cgs_model_exchange_round(cgs_model *m, int32_t n_ctas);
int cgs_model_exchange_wait(cgs_model *m);
/* The synchronous variant, for when nothing overlaps the exchange (between two sweeps): ONE kernel on the model's own
 * stream filling the GPU -- delta = n_wk - G (G = the global table of the last exchange, kept in the arena's receive
 * buffer), slice owners add all ranks' deltas to G and store the new G into every rank's receive buffer over NVLink,
 * n_wk = G, n_k = its column sums.  5 table passes of local traffic instead of the 8 of delta_compute + all-reduce +
 * delta_apply; same statistics as those.  Do not mix with cgs_model_exchange_round on one model. */
int cgs_model_exchange_sync(cgs_model *m, int32_t n_ctas);
/* The exchange OVERLAPPED with sampling, in one launch (csrc/sampler_kernels.cuh, sweep_kernel<.., XCH = true>).
 * Tokens are sorted by region inside a cell, so a sweep can be cut at a region index into two half-sweeps; while one
 * half is sampled, the rows of n_wk of the OTHER half -- changed by the previous launch, untouched by this one -- go
 * through the synchronous exchange on x_ctas blocks of the same grid (default: one per two SMs).  Every row is exchanged
 * once per sweep, right after it was sampled: same statistics as cgs_model_exchange_sync after every sweep, but the
 * NVLink round trips, the cross-rank waits and the table passes hide behind the sampling.
 *   cgs_corpus_set_region_split: the region index that cuts the sweep (every rank the same; pick the one that halves
 *     the tokens).  cgs_model_sweep_overlapped: n_sweeps sweeps, each two launches; the rows of the last half-sweep
 *     stay unpublished until the next call or cgs_model_exchange_flush (call it before reading counts /
 *     log-likelihoods).  Needs cgs_model_exchange_open and one cgs_model_exchange_sync after init.
 *   The exchange blocks of all ranks wait for each other, so they must be resident at once: x_ctas is clamped to what the
 *   sweep kernel can keep resident beside one sampling block, and nothing else may occupy the GPU while the launch
 *   starts (other work on the device delays the exchange; after 4 s a wait gives up and cgs_model_exchange_status
 *   reports the timeout). */
int cgs_corpus_set_region_split(cgs_corpus *c, int32_t region_split);
/* The general form: n_blocks region blocks, block b = regions [bounds[b], bounds[b + 1]) (bounds[0] = 0, bounds[n_blocks] =
 * n_regions; every rank the same).  A sweep is then n_blocks launches, each over one block of every cell, and while block b
 * is sampled the rows of block b - 1 are exchanged.  Two blocks are what the overlap needs and the measured best (L2-sized
 * slices of a 512 MB table did not pay: the sweep is issue / LSU bound, profiles/r2_region_blocks.txt).
 * cgs_model_sweep_blocked is the same block-wise sweep without any exchange. */
int cgs_corpus_set_region_blocks(cgs_corpus *c, int32_t n_blocks, const int32_t *bounds);
int cgs_model_sweep_blocked(cgs_model *m, int32_t n_sweeps);
int cgs_model_sweep_overlapped(cgs_model *m, int32_t n_sweeps, int32_t x_ctas);
int cgs_model_exchange_flush(cgs_model *m);
int cgs_exchange_sync_local(cgs_model *const *models, int32_t n_models, int32_t n_ctas);
/* Diagnostics: microseconds CTA 0 of the LAST synchronous exchange spent in { delta pass, barrier, reduce + broadcast of
 * its slice, barrier, apply pass } (device %globaltimer). */
int cgs_model_exchange_phase_times(cgs_model *m, double *us5);
/* Several ranks on ONE device (tests; more shards than GPUs): kernels that wait for each other must not be separate
 * launches on one GPU, so the rounds of all n_models (<= 8) local ranks go out as ONE cooperative launch (grid =
 * n_ctas x n_models, co-residency guaranteed); same semantics as calling cgs_model_exchange_round on each. */
int cgs_exchange_round_local(cgs_model *const *models, int32_t n_models, int32_t n_ctas);
/* Waits for the exchange stream; *timeouts = barrier waits that gave up after 4 s because a peer never arrived
 * (0 = every round completed; anything else means the tables must not be used). */
int cgs_model_exchange_status(cgs_model *m, int32_t *timeouts);

/* CUDA IPC plumbing so that ranks (one process per GPU) can map each other's delta buffers:
 * export a 64-byte handle of a cgs-owned device buffer, open a peer's handle, close it. */
int cgs_ipc_export(void *device_ptr, void *handle_out_64_bytes);
int cgs_ipc_open(int device, const void *handle_64_bytes, void **device_ptr);
int cgs_ipc_close(int device, void *device_ptr);
/* Plain cudaMalloc / cudaFree on `device`: IPC-exportable scratch for a host that wants to probe whether its ranks can map
 * each other's memory before it commits to the peer-memory exchanges. */
int cgs_device_malloc(int device, size_t bytes, void **device_ptr);
int cgs_device_free(int device, void *device_ptr);

/* ---- imputation (replaces the fp32 chunk product of diff_features.py:443-464) --------- */
/* out[r, c] = trunc_int32( (sum_k topic_region[r, k] * cell_topic[k, c]) * scale ) when
 * as_int32 != 0, else the fp32 product (times scale if scale != 1).  All pointers are HOST
 * pointers; topic_region is n_regions x n_topics, cell_topic is n_topics x n_cells, both
 * fp32 row-major.  row_nonzero[r] (uint8, may be NULL) flags rows with any non-zero entry. */
int cgs_impute_chunk(int device, const float *topic_region, const float *cell_topic,
                     int32_t n_regions, int32_t n_topics, int32_t n_cells,
                     float scale, int as_int32, void *out, uint8_t *row_nonzero);
/* Device-resident variant used by the benchmark (all pointers on `device`). */
int cgs_impute_chunk_device(int device, const float *d_topic_region, const float *d_cell_topic,
                            int32_t n_regions, int32_t n_topics, int32_t n_cells,
                            float scale, int as_int32, void *d_out, uint8_t *d_row_nonzero,
                            void *stream);

/* impute_accessibility multiplies EVERY chunk of topic_region with the same cell_topic
 * (diff_features.py:443-456): a plan holds the cell operand on the device in the layout the
 * tensor-core kernel reads (cells x topics, TF32 hi/lo split), prepared once.
 * cell_topic is n_topics x n_cells fp32 row-major, host or device memory. */
typedef struct cgs_impute_plan cgs_impute_plan;
int cgs_impute_plan_create(cgs_impute_plan **out, int device, const float *cell_topic,
                           int cell_topic_on_device, int32_t n_topics, int32_t n_cells, void *stream);
int cgs_impute_plan_destroy(cgs_impute_plan *plan);
/* One chunk of regions against the plan; host pointers (copies included) ... */
int cgs_impute_plan_chunk(cgs_impute_plan *plan, const float *topic_region, int32_t n_regions,
                          float scale, int as_int32, void *out, uint8_t *row_nonzero);
/* ... or device pointers, asynchronous on `stream`. */
int cgs_impute_plan_chunk_device(cgs_impute_plan *plan, const float *d_topic_region, int32_t n_regions,
                                 float scale, int as_int32, void *d_out, uint8_t *d_row_nonzero,
                                 void *stream);
int cgs_impute_plan_launch_count(const cgs_impute_plan *plan, int64_t *launches);

/* ---- normalize_scores (diff_features.py:511-574; inner function :540-548) ---------------------- */
/* out = log1p( x / (colsum(x) / scale_factor) ), x = n_rows (regions) x n_cols (cells) row-major.
 * dtype: 0 = int32 (out is float64, column sums exact in int64, as NumPy), 1 = float32 (out float32),
 * 2 = float64 (out float64).  Host pointers; the whole matrix and its result must fit on the device. */
int cgs_normalize_scores(int device, const void *x, int dtype, int64_t n_rows, int64_t n_cols,
                         double scale_factor, void *out);
/* Device-resident variant; d_colsum_scratch holds 8 * n_cols bytes and returns the per-cell totals
 * (int64 for dtype 0, float64 otherwise). */
int cgs_normalize_scores_device(int device, const void *d_x, int dtype, int64_t n_rows, int64_t n_cols,
                                double scale_factor, void *d_out, void *d_colsum_scratch, void *stream);

/* ---- find_highly_variable_features (diff_features.py:577-714): the reduction over the matrix (:636-639) -------- */
/* mean[r] = np.mean(x[r], dtype=float32), var[r] = np.var(x[r], dtype=float32) (population variance, two-pass, with
 * NumPy's casts; sums accumulated in fp64, so the float32 results are correctly rounded where NumPy's pairwise
 * float32 sums are a few ulps off).  x is n_rows (regions) x n_cols (cells) row-major with row stride ld; float32
 * [n_rows] outputs.  dtype: 0 = int32, 1 = float32, 2 = float64.  Host pointers; rows are processed in rounds. */
int cgs_row_mean_var(int device, int dtype, const void *x, int64_t ld, int64_t n_rows, int64_t n_cols, float *mean,
                     float *var);
int cgs_row_mean_var_device(int device, int dtype, const void *d_x, int64_t ld, int64_t n_rows, int64_t n_cols,
                            float *d_mean, float *d_var, void *stream);

/* ---- make_rankings (diff_features.py:274-337; the per-column closure :294-306) ---------------------------- */
/* out[r][c] = ranking of region r within cell (column) c by DESCENDING score, 0 = highest; int32, row-major with
 * row stride ld_out.  x is n_rows (regions) x n_cols (cells) row-major with row stride ld (elements).
 * Ties are broken at random, as the reference does (rng.permutation, then argsort): the column is read in the order
 * of a keyed bijection pi of [0, n_rows) -- 4-round Feistel, cycle walking, keyed by (seed, first_col + c), see
 * csrc/rankings_kernels.cuh and oracle/rankings_oracle.py -- and sorted stably, so a column without ties gets exactly
 * the reference's ranking and tied scores get a pseudo-random order that depends only on (seed, cell index).
 * NaN scores rank last.  dtype: 0 = int32, 1 = float32, 2 = float64.  Host pointers; columns are processed in
 * device-sized rounds (first_col = cell index of column 0, so a caller may split the cells itself). */
int cgs_rank_columns(int device, int dtype, const void *x, int64_t ld, int64_t n_rows, int64_t n_cols, uint64_t seed,
                     int64_t first_col, int32_t *out, int64_t ld_out);
/* Device-resident variant, asynchronous on `stream` (staging comes from the device's stream-ordered pool). */
int cgs_rank_columns_device(int device, int dtype, const void *d_x, int64_t ld, int64_t n_rows, int64_t n_cols,
                            uint64_t seed, int64_t first_col, int32_t *d_out, int64_t ld_out, void *stream);

/* ---- differential features (diff_features.py:839-1120: markers -> get_wilcox_test_pvalues + get_log2_fc) ---- */
/* For every row r < n_rows: scipy.stats.ranksums(fg[r], bg[r]) -- rank sum with average ranks for ties, normal
 * approximation without tie or continuity correction -- as statistic z and two-sided pvalue, and
 * log2fc = log2((mean(fg[r]) + 1e-12) / (mean(bg[r]) + 1e-12)); float64[n_rows] each.
 * A group is described by a row-major matrix (row stride ld_* elements) and either a list of its columns
 * (int32[n_*]) or NULL for columns 0 .. n_* - 1, so both call shapes of the reference are covered: two
 * pre-subset matrices (get_wilcox_test_pvalues(fg_mat, bg_mat)) or ONE imputed matrix with two index lists
 * (markers()); pass the same pointer and stride twice for the latter and it is uploaded once.
 * dtype: 0 = int32, 1 = float32, 2 = float64.  Host pointers; rows are processed in device-sized rounds. */
int cgs_ranksum_rows(int device, int dtype, int64_t n_rows,
                     const void *fg, int64_t ld_fg, const int32_t *fg_cols, int32_t n_fg,
                     const void *bg, int64_t ld_bg, const int32_t *bg_cols, int32_t n_bg,
                     double *statistic, double *pvalue, double *log2fc);
/* Device-resident variant (all pointers on `device`), asynchronous on `stream`: the consumer of an imputed
 * chunk that never leaves the GPU (cgs_impute_plan_chunk_device -> this). */
int cgs_ranksum_rows_device(int device, int dtype, int64_t n_rows,
                            const void *d_fg, int64_t ld_fg, const int32_t *d_fg_cols, int32_t n_fg,
                            const void *d_bg, int64_t ld_bg, const int32_t *d_bg_cols, int32_t n_bg,
                            double *d_statistic, double *d_pvalue, double *d_log2fc, void *stream);

/* markers() of n_contrasts contrasts on ONE host matrix (find_diff_features, diff_features.py:716-836, calls markers()
 * once per contrast on the same imputed matrix): x (n_rows x ld, row-major, host) is uploaded once, in rounds, and
 * every contrast tests a round while it is on the device.  Contrast k compares the columns
 * fg_cols[fg_offsets[k] .. fg_offsets[k+1]) with bg_cols[bg_offsets[k] .. bg_offsets[k+1]); statistic / pvalue /
 * log2fc are float64 [n_contrasts x n_rows] (cgs_ranksum_rows semantics). */
int cgs_ranksum_contrasts(int device, int dtype, int64_t n_rows, const void *x, int64_t ld, int32_t n_contrasts,
                          const int32_t *fg_cols, const int64_t *fg_offsets, const int32_t *bg_cols,
                          const int64_t *bg_offsets, double *statistic, double *pvalue, double *log2fc);

/* ---- fused consumers of the imputed matrix -----------------------------
 * impute_accessibility returns a dense regions x cells matrix (diff_features.py:434-441, :503) that normalize_scores
 * (:511-574) and find_diff_features (:716-836) then read back.  These entry points take the two FACTORS instead
 * (topic_region n_regions x n_topics, cell_topic n_topics x n_cells, fp32 row-major host pointers, exactly the
 * operands of cgs_impute_chunk) and recompute the product chunk by chunk on the device, so the matrix never exists. */
/* normalize_scores(impute_accessibility(...)): out receives log1p(x / (colsum(x) / norm_scale)) for the regions that
 * have a non-zero imputed value, in order (float64 when as_int32 != 0, else float32; room for n_regions rows);
 * *n_kept = number of rows written; row_nonzero (uint8[n_regions]) and column_totals (int64[n_cells] when
 * as_int32 != 0, else float64[n_cells]) may be NULL.  Two passes over the chunks: totals, then normalisation. */
int cgs_impute_normalized(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                          int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                          int32_t chunk_rows, void *out, uint8_t *row_nonzero, void *column_totals, int64_t *n_kept);
/* markers() for n_contrasts contrasts in one pass: contrast k compares the cells fg_cols[fg_offsets[k] ..
 * fg_offsets[k+1]) with bg_cols[bg_offsets[k] .. bg_offsets[k+1]); statistic / pvalue / log2fc are float64
 * [n_contrasts x n_regions] (cgs_ranksum_rows semantics); every chunk is imputed once and read by all contrasts. */
int cgs_impute_ranksum(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                       int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, int32_t chunk_rows,
                       int32_t n_contrasts, const int32_t *fg_cols, const int64_t *fg_offsets,
                       const int32_t *bg_cols, const int64_t *bg_offsets,
                       double *statistic, double *pvalue, double *log2fc, uint8_t *row_nonzero);
/* make_rankings(impute_accessibility(...)) (diff_features.py:274-337 on the result of :340-508): per block of cells the
 * whole regions x block product is imputed on the device and ranked there (cgs_rank_columns_device semantics, tie
 * order keyed by (seed, cell index)); out = int32 rankings, n_regions x n_cells row-major host memory.  The
 * reference ranks only the regions impute_accessibility kept: pass topic_region without its all-zero rows.
 * cell_block <= 0 picks a block of about 2^29 values.  first_cell = global index of cell 0 of cell_topic (keys the tie
 * order); out has row stride ld_out. */
int cgs_impute_rankings(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                        int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, uint64_t seed,
                        int32_t cell_block, int64_t first_cell, int32_t *out, int64_t ld_out);

/* ---- sharding the consumers over GPUs (SURVEY.md section 8e iii: regions are independent, cells couple only through
 * the per-cell totals of normalize_scores) -----------------------------------------------------------------------
 * Every rank / device gets a contiguous range of regions (rows of topic_region) and the whole cell_topic.
 * cgs_impute_ranksum, cgs_impute_chunk* and cgs_impute_row_stats with norm_scale == 0 need nothing else: call them on
 * the shard.  The normalised consumers need the per-cell totals of the WHOLE matrix: every rank computes the totals
 * of its shard (cgs_impute_column_totals: int64[n_cells] when as_int32 != 0, else float64[n_cells]), the ranks add
 * them up -- the one exchange step, an all-reduce of n_cells values (torch.distributed / NCCL between processes, a
 * host sum between the threads of one process) -- and pass the sum to the _shard variants below.  make_rankings
 * shards by CELLS instead: cgs_impute_rankings on cell_topic[:, c0:c1] with first_cell = c0 and out + c0, ld_out =
 * total number of cells (the tie order is keyed by the global cell index). */
/* The imputed matrix itself (diff_features.py:443-486) for a range of regions, written straight to its final place:
 * rows with keep[r] != 0 (all rows when keep == NULL) are copied, in order and without gaps, to out (n_cells values of
 * 4 bytes per row: int32 when as_int32 != 0, else float32); *n_written = rows written.  The row filter of
 * impute_accessibility (:468-486) is keep = row_nonzero of a first pass (cgs_impute_column_totals), so that the host
 * never holds a chunk it then has to compact.  Product of chunk i + 1 overlaps the copy of chunk i. */
int cgs_impute_rows(int device, const float *topic_region, const float *cell_topic, int32_t n_regions, int32_t n_topics,
                    int32_t n_cells, float impute_scale, int as_int32, int32_t chunk_rows, const uint8_t *keep, void *out,
                    int64_t *n_written);
int cgs_impute_column_totals(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                             int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, int32_t chunk_rows,
                             void *column_totals, uint8_t *row_nonzero);
int cgs_impute_normalized_shard(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                                int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                                int32_t chunk_rows, const void *column_totals, void *out, uint8_t *row_nonzero,
                                int64_t *n_kept);
int cgs_impute_row_stats_shard(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                               int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                               int32_t chunk_rows, const void *column_totals, float *mean, float *var,
                               uint8_t *row_nonzero);
/* The same reduction on normalize_scores(impute_accessibility(...)) (norm_scale > 0: statistics of
 * log1p(x / (colsum(x) / norm_scale)), float64 values for the int32 imputation, float32 otherwise) or on the imputed
 * values themselves (norm_scale == 0), from the two factors: mean / var are float32[n_regions] over ALL regions,
 * row_nonzero (may be NULL) tells which of them impute_accessibility keeps. */
int cgs_impute_row_stats(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                         int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                         int32_t chunk_rows, float *mean, float *var, uint8_t *row_nonzero);

/* ---- count matrix from fragments ------------------------------------------------
 * create_cistopic_object_from_fragments counts fragments in regions with `regions.join(fragments)` (pyranges) and
 * `counts_df.groupby(["Name", "regionID"]).size().unstack()` (cistopic_class.py:868-889).  Here every fragment finds
 * the regions it overlaps on the device and the counts come back as a regions x cells CSR (int32), the layout of
 * CistopicObject.fragment_matrix; binarising it (sklearn binarize, cistopic_class.py:592) gives the binary_matrix
 * cgs_corpus_from_regions_by_cells takes.
 * Regions: n_chrom chromosomes, region_chrom_ptr[n_chrom + 1] the region range of each, region_start / region_end
 * int32[R] sorted by start inside a chromosome (half-open, as BED).  Fragments: int32[n_fragments] chromosome id
 * (< 0 or >= n_chrom: no regions there), start, end, cell id (< 0 or >= n_cells: barcode not kept).  A fragment
 * counts once for every region it overlaps (fragment.start < region.end and region.start < fragment.end). */
typedef struct cgs_fragment_counts cgs_fragment_counts;
int cgs_fragment_counts_create(cgs_fragment_counts **out, int device, int32_t n_chrom, const int32_t *region_chrom_ptr,
                               const int32_t *region_start, const int32_t *region_end, int64_t n_fragments,
                               const int32_t *frag_chrom, const int32_t *frag_start, const int32_t *frag_end,
                               const int32_t *frag_cell, int32_t n_cells);
int cgs_fragment_counts_destroy(cgs_fragment_counts *h);
/* nnz = stored (region, cell) pairs; n_hits = fragment-region overlaps (the sum of the counts). */
int cgs_fragment_counts_dims(const cgs_fragment_counts *h, int32_t *n_regions, int32_t *n_cells, int64_t *nnz,
                             int64_t *n_hits);
/* CSR to host: indptr int64[n_regions + 1], indices int32[nnz] (cells ascending inside a region), counts int32[nnz]. */
int cgs_fragment_counts_export(const cgs_fragment_counts *h, int64_t *indptr, int32_t *indices, int32_t *counts);

/* ---- fragments / regions BED files (host code, no device needed) -------------------
 * The reader in front of cgs_fragment_counts_*: replaces read_fragments_to_pyranges (fragments.py:21-165, called at
 * cistopic_class.py:814-817) and pyranges.read_bed (cistopic_class.py:835, :578) for the columns the path reads.
 * Plain or gzip / bgzip text.  Blank lines and lines starting with '#' BEFORE the first record are skipped and the first
 * record sets the column count (fragments.py:67-85); fewer than min_columns (4 for fragments, 3 for regions) fails with
 * the reference's message (:87-91).  Columns: 1 Chromosome and 4 Name (barcode) as dictionary codes -- the dictionaries
 * SORTED, as the categories of a pandas "category" column are; 2 Start and 3 End as int32; 5 Score as int32 when every
 * value is an integer.  Later columns are ignored; a record with another column count, or a Start / End that is not a
 * 32-bit integer, is an error naming the line.  n_threads <= 0: all host cores. */
typedef struct cgs_bed_table cgs_bed_table;
int cgs_bed_read(cgs_bed_table **out, const char *path, int32_t min_columns, int32_t n_threads);
int cgs_bed_destroy(cgs_bed_table *t);
/* score_state: 0 = no fifth column, 1 = int32 scores, 2 = a fifth column that is not all integers (exported as 0).
 * *_chars = total bytes of the dictionary strings.  Any output may be NULL. */
int cgs_bed_dims(const cgs_bed_table *t, int64_t *n_rows, int32_t *n_columns, int32_t *n_chromosomes,
                 int64_t *chromosome_chars, int32_t *n_names, int64_t *name_chars, int32_t *score_state);
/* Columns int32[n_rows] in file order; dictionaries as concatenated UTF-8 bytes + int64[n + 1] offsets.  Any output may
 * be NULL (name / score must be NULL when the table has no such column). */
int cgs_bed_export(const cgs_bed_table *t, int32_t *chromosome, int32_t *start, int32_t *end, int32_t *name, int32_t *score,
                   char *chromosome_chars, int64_t *chromosome_offsets, char *name_chars, int64_t *name_offsets);
/* collapse_duplicates (utils.py:284-293, applied at cistopic_class.py:819-829 when the file has no Score column): records
 * equal in (Chromosome, Start, End, Name) become one record, Score = how many there were; the first of each group keeps
 * its place. */
int cgs_bed_collapse_duplicates(cgs_bed_table *t, int32_t n_threads, int64_t *n_removed);

#ifdef __cplusplus
}
#endif
#endif /* CGS_B200_H */
