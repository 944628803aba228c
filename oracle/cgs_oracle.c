/*
 * cgs_oracle.c -- TEST INFRASTRUCTURE ONLY (not part of the product path).
 *
 * CPU restatement of the collapsed Gibbs sampler that pycisTopic's
 * run_cgs_model drives through the third-party PyPI package `lda`
 * (call sites: /root/reference/src/pycisTopic/lda_models.py:248-287, fit at :287,
 * attributes consumed at :292-305 and :345-357).  The `lda` sources are NOT
 * vendored in /root/reference and the package is not installable here (no
 * network), so this file restates the published algorithm of lda-project/lda
 * (lda/_lda.pyx: _sample_topics, _loglikelihood, searchsorted; lda/lda.py:
 * _initialize, _fit; lda/utils.py: matrix_to_lists) as recorded in
 * SURVEY.md Appendix B.  PARITY UNPINNED for the sampler: the reference ships no
 * tests or golden vectors for this path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 *
 * Layouts follow the `lda` package exactly: nzw is (K, W) row-major int32,
 * ndz is (D, K) row-major int32, nz is (K) int32, all arithmetic in fp64,
 * tokens visited strictly sequentially in (cell ascending, region ascending)
 * order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* lda/utils.py matrix_to_lists: expand a D x W CSR count matrix
 * into per-token word (WS) and document (DS) ids, row-major, columns ascending.
 * For pycisTopic's binary matrix every count is 1.  Returns the token count. */
int64_t oracle_matrix_to_lists(const int64_t *indptr, const int32_t *indices,
                               const int32_t *data, int32_t D,
                               int32_t *WS, int32_t *DS)
{
    int64_t n = 0;
    for (int32_t d = 0; d < D; ++d) {
        for (int64_t p = indptr[d]; p < indptr[d + 1]; ++p) {
            int32_t c = data ? data[p] : 1;
            for (int32_t r = 0; r < c; ++r) {
                if (WS) { WS[n] = indices[p]; DS[n] = d; }
                ++n;
            }
        }
    }
    return n;
}

/* lda/lda.py LDA._initialize: deterministic z_i = i mod K and
 * the three count tables. */
void oracle_initialize(const int32_t *WS, const int32_t *DS, int32_t *ZS,
                       int64_t N, int32_t K, int32_t W,
                       int32_t *nzw, int32_t *ndz, int32_t *nz)
{
    for (int64_t i = 0; i < N; ++i) {
        int32_t z = (int32_t)(i % K);
        ZS[i] = z;
        ndz[(int64_t)DS[i] * K + z] += 1;
        nzw[(int64_t)z * W + WS[i]] += 1;
        nz[z] += 1;
    }
}

/* Rebuild the count tables from a supplied assignment vector (bookkeeping
 * oracle for the bit-exact count tests). */
void oracle_counts_from_z(const int32_t *WS, const int32_t *DS, const int32_t *ZS,
                          int64_t N, int32_t K, int32_t W,
                          int32_t *nzw, int32_t *ndz, int32_t *nz)
{
    for (int64_t i = 0; i < N; ++i) {
        int32_t z = ZS[i];
        ndz[(int64_t)DS[i] * K + z] += 1;
        nzw[(int64_t)z * W + WS[i]] += 1;
        nz[z] += 1;
    }
}

/* lda/_lda.pyx searchsorted: smallest index with arr[idx] >= v (bisect left). */
static inline int32_t searchsorted_left(const double *arr, int32_t length, double value)
{
    int32_t imin = 0, imax = length;
    while (imin < imax) {
        int32_t imid = imin + ((imax - imin) >> 1);
        if (value > arr[imid]) imin = imid + 1;
        else imax = imid;
    }
    return imin;
}

/* lda/_lda.pyx _sample_topics: one full sequential sweep.
 * alpha has K entries, eta has W entries (the package passes symmetric
 * vectors), rands has n_rand entries and token i uses rands[i % n_rand].
 * Returns 0, or -1 if the scratch allocation fails (MemoryError upstream). */
int oracle_sample_topics(const int32_t *WS, const int32_t *DS, int32_t *ZS,
                         int64_t N, int32_t K, int32_t W,
                         int32_t *nzw, int32_t *ndz, int32_t *nz,
                         const double *alpha, const double *eta,
                         const double *rands, int32_t n_rand)
{
    double *dist_sum = (double *)malloc(sizeof(double) * (size_t)K);
    if (!dist_sum) return -1;
    double eta_sum = 0.0;
    for (int32_t w = 0; w < W; ++w) eta_sum += eta[w];

    for (int64_t i = 0; i < N; ++i) {
        int32_t w = WS[i], d = DS[i], z = ZS[i];
        nzw[(int64_t)z * W + w] -= 1;
        ndz[(int64_t)d * K + z] -= 1;
        nz[z] -= 1;

        double cum = 0.0;
        for (int32_t k = 0; k < K; ++k) {
            cum += (nzw[(int64_t)k * W + w] + eta[w]) / (nz[k] + eta_sum)
                   * (ndz[(int64_t)d * K + k] + alpha[k]);
            dist_sum[k] = cum;
        }
        double r = rands[i % n_rand] * cum;
        int32_t z_new = searchsorted_left(dist_sum, K, r);
        if (z_new >= K) z_new = K - 1;   /* unreachable for r <= cum; guard only */

        ZS[i] = z_new;
        nzw[(int64_t)z_new * W + w] += 1;
        ndz[(int64_t)d * K + z_new] += 1;
        nz[z_new] += 1;
    }
    free(dist_sum);
    return 0;
}

/* AD-LDA variant used to check the multi-GPU semantics:
 * cells are split into `n_parts` contiguous ranges [part_ptr[p], part_ptr[p+1]);
 * every part samples against its own replica of nzw/nz taken at sweep start and
 * the replicas' deltas are summed afterwards.  n_parts == 1 is identical to
 * oracle_sample_topics.  cell_tok_ptr is the D+1 token offset of every cell. */
int oracle_sample_topics_adlda(const int32_t *WS, const int32_t *DS, int32_t *ZS,
                               int64_t N, int32_t K, int32_t W, int32_t D,
                               const int64_t *cell_tok_ptr,
                               int32_t *nzw, int32_t *ndz, int32_t *nz,
                               const double *alpha, const double *eta,
                               const double *rands, int32_t n_rand,
                               const int32_t *part_ptr, int32_t n_parts)
{
    (void)N; (void)D;
    size_t kw = (size_t)K * (size_t)W;
    int32_t *base = (int32_t *)malloc(sizeof(int32_t) * kw);
    int32_t *rep = (int32_t *)malloc(sizeof(int32_t) * kw);
    int32_t *base_nz = (int32_t *)malloc(sizeof(int32_t) * (size_t)K);
    int32_t *rep_nz = (int32_t *)malloc(sizeof(int32_t) * (size_t)K);
    if (!base || !rep || !base_nz || !rep_nz) { free(base); free(rep); free(base_nz); free(rep_nz); return -1; }
    memcpy(base, nzw, sizeof(int32_t) * kw);
    memcpy(base_nz, nz, sizeof(int32_t) * (size_t)K);
    int rc = 0;
    for (int32_t p = 0; p < n_parts && rc == 0; ++p) {
        memcpy(rep, base, sizeof(int32_t) * kw);
        memcpy(rep_nz, base_nz, sizeof(int32_t) * (size_t)K);
        int64_t t0 = cell_tok_ptr[part_ptr[p]], t1 = cell_tok_ptr[part_ptr[p + 1]];
        /* rands index keeps the global token position, like a single sweep */
        double *dist_sum = (double *)malloc(sizeof(double) * (size_t)K);
        if (!dist_sum) { rc = -1; break; }
        double eta_sum = 0.0;
        for (int32_t w = 0; w < W; ++w) eta_sum += eta[w];
        for (int64_t i = t0; i < t1; ++i) {
            int32_t w = WS[i], d = DS[i], z = ZS[i];
            rep[(int64_t)z * W + w] -= 1; ndz[(int64_t)d * K + z] -= 1; rep_nz[z] -= 1;
            double cum = 0.0;
            for (int32_t k = 0; k < K; ++k) {
                cum += (rep[(int64_t)k * W + w] + eta[w]) / (rep_nz[k] + eta_sum)
                       * (ndz[(int64_t)d * K + k] + alpha[k]);
                dist_sum[k] = cum;
            }
            double r = rands[i % n_rand] * cum;
            int32_t z_new = searchsorted_left(dist_sum, K, r);
            if (z_new >= K) z_new = K - 1;
            ZS[i] = z_new;
            rep[(int64_t)z_new * W + w] += 1; ndz[(int64_t)d * K + z_new] += 1; rep_nz[z_new] += 1;
        }
        free(dist_sum);
        for (size_t j = 0; j < kw; ++j) nzw[j] += rep[j] - base[j];
        for (int32_t k = 0; k < K; ++k) nz[k] += rep_nz[k] - base_nz[k];
    }
    free(base); free(rep); free(base_nz); free(rep_nz);
    return rc;
}

/* lda/_lda.pyx _loglikelihood: the Griffiths-Steyvers joint
 * log p(w, z) with the model's (already scaled) symmetric alpha and eta.
 * nd[d] = sum_k ndz[d, k]. */
double oracle_loglikelihood(const int32_t *nzw, const int32_t *ndz, const int32_t *nz,
                            const int32_t *nd, int32_t D, int32_t K, int32_t W,
                            double alpha, double eta)
{
    double ll = 0.0;
    double lgamma_eta = lgamma(eta), lgamma_alpha = lgamma(alpha);
    ll += K * lgamma(eta * W);
    for (int32_t k = 0; k < K; ++k) {
        ll -= lgamma(eta * W + nz[k]);
        for (int32_t w = 0; w < W; ++w) {
            int32_t c = nzw[(int64_t)k * W + w];
            if (c > 0) ll += lgamma(eta + c) - lgamma_eta;
        }
    }
    for (int32_t d = 0; d < D; ++d) {
        ll += lgamma(alpha * K) - lgamma(alpha * K + nd[d]);
        for (int32_t k = 0; k < K; ++k) {
            int32_t c = ndz[(int64_t)d * K + k];
            if (c > 0) ll += lgamma(alpha + c) - lgamma_alpha;
        }
    }
    return ll;
}

/* Timed loop for the CPU baseline: n_sweeps back-to-back sweeps reusing one
 * rands vector (the per-sweep shuffle is host Python in the `lda` package and
 * costs O(131072), not part of the O(N K) hot loop). */
int oracle_run_sweeps(const int32_t *WS, const int32_t *DS, int32_t *ZS,
                      int64_t N, int32_t K, int32_t W,
                      int32_t *nzw, int32_t *ndz, int32_t *nz,
                      double alpha_scalar, double eta_scalar,
                      const double *rands, int32_t n_rand, int32_t n_sweeps)
{
    double *alpha = (double *)malloc(sizeof(double) * (size_t)K);
    double *eta = (double *)malloc(sizeof(double) * (size_t)W);
    if (!alpha || !eta) { free(alpha); free(eta); return -1; }
    for (int32_t k = 0; k < K; ++k) alpha[k] = alpha_scalar;
    for (int32_t w = 0; w < W; ++w) eta[w] = eta_scalar;
    int rc = 0;
    for (int32_t s = 0; s < n_sweeps && rc == 0; ++s)
        rc = oracle_sample_topics(WS, DS, ZS, N, K, W, nzw, ndz, nz, alpha, eta, rands, n_rand);
    free(alpha); free(eta);
    return rc;
}
