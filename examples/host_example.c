/* host_example.c -- a host WITHOUT Python on the C ABI (include/cgs_b200.h): what a maintainer of a compiled caller
 * would write where pycisTopic's run_cgs_model drives lda.LDA (reference src/pycisTopic/lda_models.py:247-305).
 *
 *   gcc -std=c99 -I include examples/host_example.c -L pycistopic_b200 -lcgs_b200 -Wl,-rpath,$PWD/pycistopic_b200 -lm
 *   ./a.out [fragments.tsv[.gz]]
 *
 * 1. (optional) reads a fragments file with the native reader -- host code, works without a GPU;
 * 2. builds a small planted-topic binary matrix (cells x regions CSR), the token lists on the device, one model,
 *    z = i mod K, 150 sweeps, the log-likelihood before and after, the count tables back on the host;
 * 3. checks the invariants lda keeps: sum(nz) = sum(ndz) = sum(nzw) = number of tokens, ndz rows = cell sizes.
 * Exit code 0 = everything it could run was right (without a GPU: step 1 only, and it says so). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "cgs_b200.h"

#define CHECK(call)                                                        \
    do {                                                                   \
        if ((call) != 0) {                                                 \
            fprintf(stderr, "%s failed: %s\n", #call, cgs_last_error());   \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static uint64_t rng_state = 88172645463325252ull;
static double uniform(void) { /* xorshift64 */
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

int main(int argc, char **argv) {
    printf("cgs_b200 ABI version %d, at most %d topics\n", cgs_version(), cgs_max_topics());
    if (argc > 1) { /* ---- 1. the file reader ---- */
        cgs_bed_table *t = NULL;
        int64_t rows = 0, removed = 0, chrom_chars = 0, name_chars = 0;
        int32_t columns = 0, n_chrom = 0, n_names = 0, score = 0;
        CHECK(cgs_bed_read(&t, argv[1], 4, 0));
        CHECK(cgs_bed_dims(t, &rows, &columns, &n_chrom, &chrom_chars, &n_names, &name_chars, &score));
        printf("%s: %lld records, %d columns, %d chromosomes, %d barcodes\n", argv[1], (long long)rows, columns, n_chrom, n_names);
        if (columns == 4) { /* no Score column: collapse duplicates, as cistopic_class.py:819-829 */
            CHECK(cgs_bed_collapse_duplicates(t, 0, &removed));
            printf("  %lld duplicate records collapsed\n", (long long)removed);
        }
        CHECK(cgs_bed_destroy(t));
    }

    int devices = 0;
    if (cgs_device_count(&devices) != 0 || devices < 1) {
        printf("no CUDA device: the sampler part is skipped (there is no CPU path)\n");
        return 0;
    }

    /* ---- 2. a planted-topic binary matrix: cell d prefers the regions of block d mod 4 ---- */
    enum { D = 300, W = 2000, K = 4, SWEEPS = 150 }; /* the reference's default n_iter */
    int64_t *indptr = (int64_t *)malloc(sizeof(int64_t) * (D + 1));
    int32_t *indices = (int32_t *)malloc(sizeof(int32_t) * (size_t)D * W);
    int64_t nnz = 0;
    for (int d = 0; d < D; ++d) {
        indptr[d] = nnz;
        for (int w = 0; w < W; ++w) {
            double p = (w / (W / 4) == d % 4) ? 0.12 : 0.01;
            if (uniform() < p) indices[nnz++] = w;
        }
    }
    indptr[D] = nnz;

    cgs_corpus *corpus = NULL;
    cgs_model *model = NULL;
    double ll0 = 0, ll1 = 0;
    CHECK(cgs_corpus_from_cells_by_regions(&corpus, 0, indptr, indices, D, W, nnz, 0, D));
    CHECK(cgs_model_create(&model, corpus, K, 50.0 / K, 0.1, 555, NULL)); /* alpha / K, eta: lda_models.py:256-264 */
    CHECK(cgs_model_init(model));
    CHECK(cgs_model_loglik(model, 0, &ll0));
    CHECK(cgs_model_sweep(model, SWEEPS));
    CHECK(cgs_model_loglik(model, 0, &ll1));
    int32_t *nzw = (int32_t *)malloc(sizeof(int32_t) * K * W), *ndz = (int32_t *)malloc(sizeof(int32_t) * D * K), nz[K];
    CHECK(cgs_model_export_counts(model, nzw, ndz, nz));
    CHECK(cgs_model_destroy(model));
    CHECK(cgs_corpus_destroy(corpus));

    /* ---- 3. the bookkeeping invariants ---- */
    int64_t s_nz = 0, s_nzw = 0, s_ndz = 0, bad_rows = 0;
    for (int k = 0; k < K; ++k) s_nz += nz[k];
    for (int i = 0; i < K * W; ++i) s_nzw += nzw[i];
    for (int d = 0; d < D; ++d) {
        int64_t row = 0;
        for (int k = 0; k < K; ++k) row += ndz[d * K + k];
        s_ndz += row;
        bad_rows += row != indptr[d + 1] - indptr[d];
    }
    printf("%d cells x %d regions, %lld tokens, K = %d: log-likelihood %.1f -> %.1f after %d sweeps\n", D, W, (long long)nnz, K,
           ll0, ll1, SWEEPS);
    printf("sum(nz) = %lld, sum(nzw) = %lld, sum(ndz) = %lld, cells whose ndz row differs from their size: %lld\n",
           (long long)s_nz, (long long)s_nzw, (long long)s_ndz, (long long)bad_rows);
    int ok = s_nz == nnz && s_nzw == nnz && s_ndz == nnz && bad_rows == 0 && ll1 > ll0 && isfinite(ll1);
    printf(ok ? "OK\n" : "FAILED\n");
    free(indptr); free(indices); free(nzw); free(ndz);
    return ok ? 0 : 1;
}
