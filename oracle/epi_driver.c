/*
 * epi_driver.c -- CPU ORACLE DRIVER (test infrastructure, NOT product code).
 *
 * One source, two builds (see oracle/Makefile):
 *   -DDRV_ORACLE : linked with epi_oracle.c        -> oracle/liboracle.so, symbols oracle_*
 *   -DDRV_REF    : linked with the reference's own -> oracle/_ref/libhpgref.so, symbols ref_*
 *                  model.c / mdr.c / dataset.c / cross_validation.c / epistasis.c ...
 *
 * The driver owns what the reference's runner cannot be trusted with for
 * parity work: fold assignment is INJECTED as a
 * per-sample fold id, combinations are enumerated exhaustively in
 * lexicographic order (or given explicitly), and the per-fold top-N uses the
 * canonical total order (BA descending, SNP tuple ascending).  Everything
 * else -- padded rows, byte masks, per-fold counts, the float32 high-risk
 * rule, the confusion matrix, BA -- is done by calling the leaf functions in
 * the same sequence as process_set_of_combinations (epistasis.c:4-93), in
 * batches of 16 combinations (model.h:44).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(DRV_REF)
#include "cross_validation.h"
#include "dataset.h"
#include "epistasis.h"
#include "model.h"
#include "epistasis_runner.h"
#define DRV(name) ref_##name
#elif defined(DRV_ORACLE)
#include "epi_oracle.h"
#define DRV(name) oracle_##name
#else
#error "build with -DDRV_REF or -DDRV_ORACLE"
#endif

#define ROW 16   /* COMBINATIONS_ROW_SSE, model.h:44 */

/* Same layout as hpgv_epi_model_t in include/hpgv_epi.h (40 bytes). */
typedef struct {
    double ba;
    int32_t snp[3];
    uint32_t risky_mask;     /* bit c = cell c is high-risk, c = sum g_j * 3^(order-1-j) */
    uint32_t conf[4];        /* TP, FN, FP, TN */
} epi_model_rec;

/* ---------------- combinatorics (driver-owned enumeration) ---------------- */

static uint64_t choose_u64(uint64_t n, int k) {
    if (k < 0 || (uint64_t) k > n) return 0;
    uint64_t r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - (uint64_t) k + (uint64_t) i) / (uint64_t) i;
    return r;
}

uint64_t DRV(num_combinations)(int nv, int order) { return choose_u64((uint64_t) nv, order); }

/* idx-th k-subset of {0..n-1} in lexicographic order */
void DRV(unrank)(int n, int k, uint64_t idx, int32_t *comb) {
    int x = 0;
    for (int i = 0; i < k; i++) {
        for (;; x++) {
            uint64_t c = choose_u64((uint64_t) (n - x - 1), k - i - 1);
            if (idx < c) break;
            idx -= c;
        }
        comb[i] = x++;
    }
}

static int next_comb(int n, int k, int32_t *comb) {
    int i = k - 1;
    while (i >= 0 && comb[i] == n - k + i) i--;
    if (i < 0) return 0;
    comb[i]++;
    for (int j = i + 1; j < k; j++) comb[j] = comb[j - 1] + 1;
    return 1;
}

/* ---------------- shared per-call state ---------------- */

typedef struct {
    int nv, A, U, order, F, C;
    masks_info info;
    uint8_t *rows;            /* nv padded rows of S_pad bytes */
    uint8_t *fold_masks;      /* F x S_pad, 1 = training */
    unsigned int *test_sizes; /* 3F: total, cases, controls IN the fold */
    int *train_sizes;         /* 3F */
    int *test_sizes_i;        /* 3F, int copy */
    uint8_t **cells;
    int **folds;
} drv_state;

static int cmp_int_drv(const void *a, const void *b) { return *(const int *) a - *(const int *) b; }

static int drv_init(drv_state *st, const uint8_t *geno, int nv, int A, int U, int order, int F,
                    const int32_t *fold_of_sample) {
    memset(st, 0, sizeof(*st));
    st->nv = nv; st->A = A; st->U = U; st->order = order; st->F = F;
    masks_info_init(order, ROW, A, U, &st->info);
    st->C = st->info.num_cell_counts_per_combination;
    const int S = st->info.num_samples_with_padding;
    if (posix_memalign((void **) &st->rows, 16, (size_t) nv * S) != 0) return -1;
    /* stride = nv, block 0: every variant, padded (cross_validation.c:160-195) */
    get_genotypes_of_block_coord(nv, A + U, st->info, nv, 0, (uint8_t *) geno, st->rows);

    /* folds as the reference stores them: sorted sample ids, cases first */
    st->folds = malloc((size_t) F * sizeof(int *));
    st->test_sizes = calloc(3 * (size_t) F, sizeof(unsigned int));
    for (int s = 0; s < A + U; s++) {
        int f = fold_of_sample[s];
        if (f < 0 || f >= F) return -2;
        st->test_sizes[3 * f]++;
        st->test_sizes[3 * f + (s < A ? 1 : 2)]++;
    }
    int *fill = calloc((size_t) F, sizeof(int));
    for (int f = 0; f < F; f++) st->folds[f] = malloc((size_t) (st->test_sizes[3 * f] ? st->test_sizes[3 * f] : 1) * sizeof(int));
    for (int s = 0; s < A + U; s++) st->folds[fold_of_sample[s]][fill[fold_of_sample[s]]++] = s;
    for (int f = 0; f < F; f++) qsort(st->folds[f], st->test_sizes[3 * f], sizeof(int), cmp_int_drv);
    free(fill);
    st->fold_masks = get_k_folds_masks((unsigned) A, (unsigned) U, (unsigned) F, st->folds, st->test_sizes);

    /* singlenode/epistasis_runner.c:100-105 */
    st->train_sizes = calloc(3 * (size_t) F, sizeof(int));
    st->test_sizes_i = calloc(3 * (size_t) F, sizeof(int));
    for (int f = 0; f < F; f++) {
        st->train_sizes[3 * f] = A + U - (int) st->test_sizes[3 * f];
        st->train_sizes[3 * f + 1] = A - (int) st->test_sizes[3 * f + 1];
        st->train_sizes[3 * f + 2] = U - (int) st->test_sizes[3 * f + 2];
        for (int j = 0; j < 3; j++) st->test_sizes_i[3 * f + j] = (int) st->test_sizes[3 * f + j];
    }
    int ncells;
    st->cells = get_genotype_combinations(order, &ncells);
    return 0;
}

static void drv_free(drv_state *st) {
    free(st->rows); free(st->fold_masks); free(st->test_sizes); free(st->train_sizes); free(st->test_sizes_i);
    for (int f = 0; f < st->F; f++) free(st->folds[f]);
    free(st->folds);
    for (int c = 0; c < st->C; c++) free(st->cells[c]);
    free(st->cells);
}

typedef struct {
    uint8_t *masks;
    int *counts_aff, *counts_unaff;
} drv_scratch;

static int scratch_init(const drv_state *st, drv_scratch *sc) {
    size_t ncounts = (size_t) st->C * ROW * st->F + 16;
    if (posix_memalign((void **) &sc->masks, 16, (size_t) ROW * st->info.num_masks) != 0) return -1;
    if (posix_memalign((void **) &sc->counts_aff, 16, ncounts * sizeof(int)) != 0) return -1;
    if (posix_memalign((void **) &sc->counts_unaff, 16, ncounts * sizeof(int)) != 0) return -1;
    return 0;
}
static void scratch_free(drv_scratch *sc) { free(sc->masks); free(sc->counts_aff); free(sc->counts_unaff); }

/* Which of evaluate_model's functions (model.c:462-479) scores a model; BA is what the runner uses (model.c:331).
 * Code 5 is the documented classification accuracy, which evaluate_model itself never reaches (code 0 becomes BA,
 * model.c:465-467): written out here so that the product's extension has a checker. */
static int g_eval_fn = BA;
void DRV(set_eval_function)(int fn) { g_eval_fn = fn; }
static double drv_evaluate(unsigned int *m) {
    if (g_eval_fn == 5) {
        double TP = m[0], FN = m[1], FP = m[2], TN = m[3];
        return (TP + TN) / (TP + FN + TN + FP);
    }
    return evaluate_model(m, (enum eval_function) g_eval_fn);
}

/* One batch of n <= 16 combinations through the reference pipeline
 * (epistasis.c:4-93).  Outputs are indexed [comb][fold]. */
static void eval_batch(const drv_state *st, drv_scratch *sc, int n, const int32_t *combs, int subset,
                       int32_t *counts_aff, int32_t *counts_unaff, uint32_t *risky_mask, uint32_t *conf, double *ba) {
    const int order = st->order, C = st->C, F = st->F, S = st->info.num_samples_with_padding;
    uint8_t *rowptr[ROW * 3];
    for (int c = 0; c < ROW; c++) {
        int src = c < n ? c : n - 1;   /* idle slots repeat the last combination */
        for (int s = 0; s < order; s++) rowptr[c * order + s] = st->rows + (size_t) combs[src * order + s] * S;
    }
    set_genotypes_masks(order, rowptr, ROW, sc->masks, st->info);
    combination_counts_all_folds(order, st->fold_masks, F, st->cells, sc->masks, st->info, sc->counts_aff, sc->counts_unaff);

    for (int f = 0; f < F; f++) {
        unsigned int num_risky[ROW];
        memset(num_risky, 0, sizeof(num_risky));
        void *aux = NULL;
        int *risky_idx = choose_high_risk_combinations2(
            (unsigned int *) sc->counts_aff + (size_t) f * ROW * C, (unsigned int *) sc->counts_unaff + (size_t) f * ROW * C,
            ROW, (unsigned) C, (unsigned) st->info.num_affected, (unsigned) st->info.num_unaffected,   /* epistasis.c:37: dataset-level A, U */
            num_risky, &aux, mdr_high_risk_combinations2);
        int begin = 0;
        for (int rc = 0; rc < n; rc++) {
            size_t o = (size_t) rc * F + f;
            if (counts_aff) {
                memcpy(counts_aff + o * C, sc->counts_aff + (size_t) f * ROW * C + (size_t) rc * C, (size_t) C * sizeof(int));
                memcpy(counts_unaff + o * C, sc->counts_unaff + (size_t) f * ROW * C + (size_t) rc * C, (size_t) C * sizeof(int));
            }
            uint32_t mask = 0;
            for (unsigned r = 0; r < num_risky[rc]; r++) mask |= 1u << risky_idx[begin + (int) r];
            unsigned int m[4];
            if (num_risky[rc] > 0) {
                int comb[3];
                for (int s = 0; s < order; s++) comb[s] = combs[rc * order + s];
                risky_combination *rcomb = risky_combination_new(order, comb, st->cells, (int) num_risky[rc],
                                                                 risky_idx + begin, NULL, st->info);
                confusion_matrix(order, rcomb, rowptr + rc * order, st->fold_masks + (size_t) f * S,
                                 (enum evaluation_subset) subset, st->train_sizes + 3 * f + 1, st->test_sizes_i + 3 * f + 1,
                                 st->info, m);
                risky_combination_free(rcomb);
            } else {
                /* the reference reads a zero-length array here; defined as "predict nobody". */
                const int *sz = (subset == TRAINING) ? st->train_sizes + 3 * f + 1 : st->test_sizes_i + 3 * f + 1;
                m[0] = 0; m[1] = (unsigned) sz[0]; m[2] = 0; m[3] = (unsigned) sz[1];
            }
            begin += (int) num_risky[rc];
            if (risky_mask) risky_mask[o] = mask;
            if (conf) memcpy(conf + o * 4, m, sizeof(m));
            if (ba) ba[o] = drv_evaluate(m);
        }
        free(risky_idx);
    }
}

/* -------- public: evaluate an explicit list of combinations -------- */
int DRV(eval)(const uint8_t *geno, int nv, int A, int U, int order, int F, const int32_t *fold_of_sample, int subset,
              int64_t ncomb, const int32_t *combs,
              int32_t *counts_aff, int32_t *counts_unaff, uint32_t *risky_mask, uint32_t *conf, double *ba) {
    drv_state st;
    if (order < 2 || order > 3) return -3;
    int rc = drv_init(&st, geno, nv, A, U, order, F, fold_of_sample);
    if (rc) return rc;
    drv_scratch sc;
    if (scratch_init(&st, &sc)) return -1;
    for (int64_t b = 0; b < ncomb; b += ROW) {
        int n = (int) ((ncomb - b) < ROW ? (ncomb - b) : ROW);
        size_t o = (size_t) b * F;
        eval_batch(&st, &sc, n, combs + b * order, subset,
                   counts_aff ? counts_aff + o * st.C : NULL, counts_unaff ? counts_unaff + o * st.C : NULL,
                   risky_mask ? risky_mask + o : NULL, conf ? conf + o * 4 : NULL, ba ? ba + o : NULL);
    }
    scratch_free(&sc);
    drv_free(&st);
    return 0;
}

/* -------- canonical top-N -------- */

/* returns 1 when a ranks strictly before b: BA descending (NaN last), then SNP tuple ascending */
static int rec_before(const epi_model_rec *a, const epi_model_rec *b, int order) {
    int an = isnan(a->ba), bn = isnan(b->ba);
    if (an != bn) return bn;
    if (!an && a->ba != b->ba) return a->ba > b->ba;
    for (int s = 0; s < order; s++) if (a->snp[s] != b->snp[s]) return a->snp[s] < b->snp[s];
    return 0;
}

typedef struct { epi_model_rec *v; int n, cap; } toplist;

static void top_insert(toplist *t, const epi_model_rec *r, int order) {
    if (t->n == t->cap && !rec_before(r, &t->v[t->n - 1], order)) return;
    int pos = t->n < t->cap ? t->n : t->n - 1;
    while (pos > 0 && rec_before(r, &t->v[pos - 1], order)) { t->v[pos] = t->v[pos - 1]; pos--; }
    t->v[pos] = *r;
    if (t->n < t->cap) t->n++;
}

/* -------- public: exhaustive search over linear indices [first, last) -------- */
int DRV(search)(const uint8_t *geno, int nv, int A, int U, int order, int F, const int32_t *fold_of_sample, int subset,
                int topn, uint64_t first, uint64_t last, int num_threads,
                epi_model_rec *out /*[F][topn]*/, int32_t *n_out /*[F]*/) {
    drv_state st;
    if (order < 2 || order > 3) return -3;
    int rc = drv_init(&st, geno, nv, A, U, order, F, fold_of_sample);
    if (rc) return rc;
    uint64_t total = choose_u64((uint64_t) nv, order);
    if (last > total) last = total;
    if (first > last) first = last;
    if (num_threads < 1) num_threads = 1;

    toplist *all = calloc((size_t) num_threads * F, sizeof(toplist));
    for (int i = 0; i < num_threads * F; i++) { all[i].cap = topn; all[i].v = malloc((size_t) topn * sizeof(epi_model_rec)); }

    const uint64_t WORK_ITEM = 4096;   /* combinations per work item (multiple of ROW) */
    uint64_t nchunks = (last - first + WORK_ITEM - 1) / WORK_ITEM;
    int failed = 0;
#pragma omp parallel num_threads(num_threads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        drv_scratch sc;
        int ok = scratch_init(&st, &sc) == 0;
        if (!ok) {
#pragma omp atomic write
            failed = 1;
        }
        uint32_t mask[ROW * 64];
        uint32_t conf[ROW * 64 * 4];
        double ba[ROW * 64];
        int32_t combs[ROW * 3];
#pragma omp for schedule(dynamic, 1)
        for (uint64_t ch = 0; ch < nchunks; ch++) {
            if (!ok) continue;
            uint64_t lo = first + ch * WORK_ITEM, hi = lo + WORK_ITEM < last ? lo + WORK_ITEM : last;
            int32_t cur[3];
            DRV(unrank)(nv, order, lo, cur);
            uint64_t idx = lo;
            while (idx < hi) {
                int n = 0;
                while (n < ROW && idx < hi) {
                    memcpy(combs + n * order, cur, (size_t) order * sizeof(int32_t));
                    n++; idx++;
                    next_comb(nv, order, cur);
                }
                eval_batch(&st, &sc, n, combs, subset, NULL, NULL, mask, conf, ba);
                for (int c = 0; c < n; c++) for (int f = 0; f < F; f++) {
                    epi_model_rec r;
                    memset(&r, 0, sizeof(r));
                    r.ba = ba[c * F + f];
                    for (int s = 0; s < order; s++) r.snp[s] = combs[c * order + s];
                    for (int s = order; s < 3; s++) r.snp[s] = -1;
                    r.risky_mask = mask[c * F + f];
                    memcpy(r.conf, conf + (size_t) (c * F + f) * 4, 4 * sizeof(uint32_t));
                    top_insert(&all[tid * F + f], &r, order);
                }
            }
        }
        if (ok) scratch_free(&sc);
    }
    if (F > 64) failed = 1;
    for (int f = 0; f < F && !failed; f++) {
        toplist fin = { out + (size_t) f * topn, 0, topn };
        for (int t = 0; t < num_threads; t++)
            for (int i = 0; i < all[t * F + f].n; i++) top_insert(&fin, &all[t * F + f].v[i], order);
        n_out[f] = fin.n;
        for (int i = fin.n; i < topn; i++) { memset(&out[(size_t) f * topn + i], 0, sizeof(epi_model_rec)); out[(size_t) f * topn + i].ba = NAN; out[(size_t) f * topn + i].snp[0] = out[(size_t) f * topn + i].snp[1] = out[(size_t) f * topn + i].snp[2] = -1; }
    }
    for (int i = 0; i < num_threads * F; i++) free(all[i].v);
    free(all);
    drv_free(&st);
    return failed ? -1 : 0;
}

/* -------- public: the leaf helpers, flat, for direct cross-checks -------- */

/* fold id per sample from get_k_folds (cross_validation.c:4-100); sizes_out = 3k */
int DRV(k_folds)(int A, int U, int k, long seed, int32_t *fold_of_sample, uint32_t *sizes_out);

int DRV(fold_masks)(int A, int U, int F, const int32_t *fold_of_sample, uint8_t *out /*[F][S_pad]*/) {
    drv_state st;
    uint8_t dummy = 0;
    /* nv = 0 rows: only the fold bookkeeping of drv_init is used */
    int rc = drv_init(&st, &dummy, 0, A, U, 2, F, fold_of_sample);
    if (rc) return rc;
    memcpy(out, st.fold_masks, (size_t) F * st.info.num_samples_with_padding);
    drv_free(&st);
    return 0;
}

/* float32 high-risk flags for explicit count pairs (mdr.c:45-75) */
int DRV(high_risk)(const int32_t *ca, const int32_t *cu, int n, unsigned A, unsigned U, int32_t *flags) {
    int padded = 16 * ((n + 15) / 16) + 16;
    int *a = calloc((size_t) padded, sizeof(int)), *u = calloc((size_t) padded, sizeof(int));
    memcpy(a, ca, (size_t) n * sizeof(int)); memcpy(u, cu, (size_t) n * sizeof(int));
    void *aux = NULL;
    int *r = mdr_high_risk_combinations2(a, u, n, A, U, &aux);
    for (int i = 0; i < n; i++) flags[i] = r[i] != 0;
    free(r); free(a); free(u);
    return 0;
}

double DRV(evaluate)(const uint32_t *m, int function) {
    unsigned int mm[4] = { m[0], m[1], m[2], m[3] };
    return evaluate_model(mm, (enum eval_function) function);
}

/* blocked enumeration exactly as singlenode/epistasis_runner.c:114-125,241-258 walks it;
 * returns the number of combinations written (capacity `cap`), duplicates and all. */
int64_t DRV(enumerate_blocked)(int nv, int order, int stride, int32_t *out, int64_t cap) {
    int nblocks = (int) ceil((double) nv / stride);
    int bc[3] = { 0, 0, 0 };
    int64_t n = 0;
    do {
        int comb[3];
        get_first_combination_in_block(order, comb, bc, stride);
        do {
            if (n < cap) for (int s = 0; s < order; s++) out[n * order + s] = comb[s];
            n++;
        } while (get_next_combination_in_block(order, comb, bc, stride, nv));
    } while (get_next_block(nblocks, order, bc));
    return n;
}

#if defined(DRV_ORACLE)
int DRV(k_folds)(int A, int U, int k, long seed, int32_t *fold_of_sample, uint32_t *sizes_out) {
    oracle_set_shuffle_seed(seed);
    unsigned int *sizes = NULL;
    int **folds = get_k_folds((unsigned) A, (unsigned) U, (unsigned) k, &sizes);
    for (int f = 0; f < k; f++) {
        for (unsigned j = 0; j < sizes[3 * f]; j++) fold_of_sample[folds[f][j]] = f;
        free(folds[f]);
    }
    memcpy(sizes_out, sizes, 3 * (size_t) k * sizeof(uint32_t));
    free(folds); free(sizes);
    return 0;
}
#endif

#if defined(DRV_REF)
/* The reference reseeds srand48 from gettimeofday().tv_usec before every
 * shuffle (lib/c/src/math/data/array_utils.c:173-188).  libhpgref.so is
 * linked with -Bsymbolic, so this definition is the one array_utils.o binds
 * to, which makes the reference's own get_k_folds reproducible. */
#include <sys/time.h>
static long g_fake_usec = -1;
int gettimeofday(struct timeval *tv, void *tz) {
    (void) tz;
    if (tv) {
        if (g_fake_usec >= 0) { tv->tv_sec = 0; tv->tv_usec = g_fake_usec; }
        else {
            struct timespec ts;
            clock_gettime(CLOCK_REALTIME, &ts);
            tv->tv_sec = ts.tv_sec; tv->tv_usec = ts.tv_nsec / 1000;
        }
    }
    return 0;
}

int DRV(k_folds)(int A, int U, int k, long seed, int32_t *fold_of_sample, uint32_t *sizes_out) {
    g_fake_usec = seed;
    unsigned int *sizes = NULL;
    int **folds = get_k_folds((unsigned) A, (unsigned) U, (unsigned) k, &sizes);
    g_fake_usec = -1;
    for (int f = 0; f < k; f++) {
        for (unsigned j = 0; j < sizes[3 * f]; j++) fold_of_sample[folds[f][j]] = f;
        free(folds[f]);
    }
    memcpy(sizes_out, sizes, 3 * (size_t) k * sizeof(uint32_t));
    free(folds); free(sizes);
    return 0;
}

/* compare_int (src/hpg_variant_utils.c:351-353) and get_output_file
 * (src/hpg_variant_utils.c:302-314): the two externals of the hot-path
 * objects that live in files dragging in the whole VCF stack. */
int compare_int(const void *a, const void *b) { return *(const int *) a - *(const int *) b; }

FILE *get_output_file(shared_options_data_t *shared, char *default_name, char **path) {
    const char *dir = (shared->output_directory && strlen(shared->output_directory) > 0) ? shared->output_directory : ".";
    const char *name = (shared->output_filename && strlen(shared->output_filename) > 0) ? shared->output_filename : default_name;
    *path = malloc(strlen(dir) + strlen(name) + 2);
    sprintf(*path, "%s/%s", dir, name);
    return fopen(*path, "w");
}

/* The reference's own end-to-end runner (singlenode/epistasis_runner.c:24),
 * with its OpenMP loop, on a dataset file with the current 12-byte header.
 * This is the CPU baseline (`cpu_baseline.kind = "reference"`). */
int DRV(run_epistasis)(const char *dataset, const char *outdir, int order, int stride, int num_folds,
                       int num_cv_repetitions, int max_ranking_size, int eval_subset, int eval_mode, int num_threads) {
    static int log_ready = 0;
    if (!log_ready) { init_log_custom(LOG_WARN_LEVEL, 1, NULL, "w"); log_ready = 1; }
    shared_options_data_t shared;
    memset(&shared, 0, sizeof(shared));
    shared.output_directory = (char *) outdir;
    shared.output_filename = (char *) "";
    shared.num_threads = num_threads;
    epistasis_options_data_t opts;
    memset(&opts, 0, sizeof(opts));
    opts.dataset_filename = (char *) dataset;
    opts.order = order;
    opts.stride = stride;
    opts.num_folds = num_folds;
    opts.num_cv_repetitions = num_cv_repetitions;
    opts.max_ranking_size = max_ranking_size;
    opts.eval_subset = (enum evaluation_subset) eval_subset;
    opts.eval_mode = (enum evaluation_mode) eval_mode;
    return run_epistasis(&shared, &opts);
}
/* -------- reference only: merge_rankings (epistasis.c:96-153) and epistasis_report (epistasis_report.c:28-82) --------
 * Per-fold rankings are given as model records (fold-major, `rank` slots per fold, snp[0] < 0 = empty slot).  They are
 * turned into the reference's own `struct heap` of risky_combination exactly the way the runner fills them
 * (risky_combination_new + add_to_model_ranking with the min comparator of the mode, singlenode/epistasis_runner.c:269-287),
 * then handed to the reference's merge_rankings.  Used by tests/golden/make_golden.py to pin a16/a17. */
static struct heap **heaps_from_models(int order, int F, int rank, const epi_model_rec *models, int eval_mode,
                                       compare_risky_heap_func *min_out, compare_risky_heap_func *max_out) {
    compare_risky_heap_func hmin = eval_mode == CV_A ? compare_risky_heap_accuracy_min : compare_risky_heap_count_min;
    compare_risky_heap_func hmax = eval_mode == CV_A ? compare_risky_heap_accuracy_max : compare_risky_heap_count_max;
    int ncells;
    uint8_t **cells = get_genotype_combinations(order, &ncells);
    masks_info info;
    masks_info_init(order, ROW, 16, 16, &info);             /* only num_cell_counts_per_combination is read */
    struct heap **h = malloc((size_t) F * sizeof(struct heap *));
    for (int f = 0; f < F; f++) {
        h[f] = malloc(sizeof(struct heap));
        heap_init(h[f]);
        for (int r = 0; r < rank; r++) {
            const epi_model_rec *m = &models[(size_t) f * rank + r];
            if (m->snp[0] < 0) continue;
            int comb[3] = { m->snp[0], m->snp[1], m->snp[2] };
            int idx[27], n = 0;
            for (int c = 0; c < ncells; c++) if ((m->risky_mask >> c) & 1u) idx[n++] = c;
            risky_combination *rc = risky_combination_new(order, comb, cells, n, idx, NULL, info);
            rc->accuracy = m->ba;
            if (add_to_model_ranking(rc, rank, h[f], hmin) < 0) risky_combination_free(rc);
        }
    }
    for (int c = 0; c < ncells; c++) free(cells[c]);
    free(cells);
    *min_out = hmin; *max_out = hmax;
    return h;
}

/* rows out (capacity cap): snp [cap][3], cv_count [cap], cv_accuracy [cap], num_risky [cap], genotypes [cap][27*3];
 * returns the number of distinct combinations, in the order the reference's report would list them */
int DRV(merge_rankings)(int order, int F, int rank, const epi_model_rec *models, int eval_mode, int cap,
                        int32_t *snp, int32_t *cv_count, double *cv_accuracy, int32_t *num_risky, uint8_t *genotypes) {
    compare_risky_heap_func hmin, hmax;
    struct heap **h = heaps_from_models(order, F, rank, models, eval_mode, &hmin, &hmax);
    struct heap *best = merge_rankings(F, h, hmin, hmax);
    int n = 0;
    while (!heap_empty(best)) {
        struct heap_node *hn = heap_take(hmax, best);
        risky_combination *e = (risky_combination *) hn->value;
        if (n < cap) {
            for (int s = 0; s < 3; s++) snp[n * 3 + s] = s < order ? e->combination[s] : -1;
            cv_count[n] = e->cross_validation_count;
            cv_accuracy[n] = e->accuracy;
            num_risky[n] = e->num_risky_genotypes;
            memcpy(genotypes + (size_t) n * 81, e->genotypes, (size_t) e->num_risky_genotypes * order);
        }
        n++;
        risky_combination_free(e);
        free(hn);
    }
    free(best);
    for (int f = 0; f < F; f++) free(h[f]);
    free(h);
    return n;
}

int DRV(report)(int order, int F, int rank, const epi_model_rec *models, int eval_mode, int eval_subset, int cv_repetition,
                int max_ranking_size, const char *path) {
    compare_risky_heap_func hmin, hmax;
    struct heap **h = heaps_from_models(order, F, rank, models, eval_mode, &hmin, &hmax);
    struct heap *best = merge_rankings(F, h, hmin, hmax);
    FILE *fd = fopen(path, "w");
    if (!fd) return -1;
    epistasis_report(order, cv_repetition, (enum evaluation_mode) eval_mode, (enum evaluation_subset) eval_subset, best,
                     max_ranking_size, hmax, fd);
    fclose(fd);
    while (!heap_empty(best)) {
        struct heap_node *hn = heap_take(hmax, best);
        risky_combination_free((risky_combination *) hn->value);
        free(hn);
    }
    free(best);
    for (int f = 0; f < F; f++) free(h[f]);
    free(h);
    return 0;
}
#endif
