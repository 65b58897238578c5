/*
 * epi_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Scalar, byte-at-a-time restatement of hpg-variant's epistasis (MDR + k-fold
 * CV) leaf functions.  The reference works on 16-byte SSE registers over
 * one-byte-per-sample masks; this file keeps the same data layouts (so the
 * reference's unit tests link against it) but spells every step out per
 * sample, which is what makes it usable as an independent check.
 *
 * Build flags matter for mdr_high_risk_combinations2: it must be compiled
 * with -ffp-contract=off (no FMA contraction) so the six float32 operations
 * round exactly like the reference's _mm_*_ps sequence.
 *
 * See epi_oracle.h for the usage policy (tests / smoke / cpu_baseline only)
 * and the parity pin.
 */
#include "epi_oracle.h"

#include <assert.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int round_up16(int n) { return 16 * ((n + 15) / 16); }

static int ipow3(int order) {
    int c = 1;
    for (int i = 0; i < order; i++) c *= NUM_GENOTYPES;
    return c;
}

static int popcount8(unsigned v) { return __builtin_popcount(v & 0xFFu); }

/* ------------------------------------------------------------------ *
 * Sizes.  model.c:208-219: each class is padded to a multiple of 16
 * samples; a combination owns 3*order byte rows of S_pad bytes.
 * ------------------------------------------------------------------ */
void masks_info_init(int order, int num_combinations_in_a_row, int num_affected, int num_unaffected, masks_info *info) {
    info->num_affected = num_affected;
    info->num_unaffected = num_unaffected;
    info->num_affected_with_padding = round_up16(num_affected);
    info->num_unaffected_with_padding = round_up16(num_unaffected);
    info->num_combinations_in_a_row = num_combinations_in_a_row;
    info->num_cell_counts_per_combination = ipow3(order);
    info->num_samples_with_padding = info->num_affected_with_padding + info->num_unaffected_with_padding;
    info->num_masks = NUM_GENOTYPES * order * info->num_samples_with_padding;
    assert(info->num_affected_with_padding);
    assert(info->num_unaffected_with_padding);
}

/* ------------------------------------------------------------------ *
 * Byte masks.  model.c:28-74: layout [combination][snp][genotype][S_pad],
 * byte = 0xFF where the sample's genotype equals g, 0 elsewhere; the two
 * padding stretches (after the cases, after the controls) are forced to 0.
 * A genotype byte outside {0,1,2} (255 = missing) matches no mask.
 * ------------------------------------------------------------------ */
void set_genotypes_masks(int order, uint8_t **genotypes, int num_combinations, uint8_t *masks, masks_info info) {
    const int S = info.num_samples_with_padding;
    const int a_end = info.num_affected, a_pad = info.num_affected_with_padding;
    const int u_end = a_pad + info.num_unaffected;
    for (int c = 0; c < num_combinations; c++) {
        for (int j = 0; j < order; j++) {
            const uint8_t *row = genotypes[c * order + j];
            for (int g = 0; g < NUM_GENOTYPES; g++) {
                uint8_t *out = masks + (size_t) c * info.num_masks + (size_t) (j * NUM_GENOTYPES + g) * S;
                for (int s = 0; s < S; s++) {
                    bool padding = (s >= a_end && s < a_pad) || (s >= u_end);
                    out[s] = (!padding && row[s] == g) ? 0xFF : 0x00;
                }
            }
        }
    }
}

/* AND of the `order` genotype masks selected by one cell, at sample s. */
static inline unsigned cell_byte(const uint8_t *rc_masks, const uint8_t *cell, int order, int S, int s) {
    unsigned v = 0xFF;
    for (int j = 0; j < order; j++) {
        v &= rc_masks[(size_t) (j * NUM_GENOTYPES + cell[j]) * S + s];
    }
    return v;
}

/* ------------------------------------------------------------------ *
 * Whole-dataset cell counts.  model.c:76-129: popcount of the ANDed 0xFF
 * masks divided by 8, cases over [0, A_pad), controls over [A_pad, S_pad).
 * ------------------------------------------------------------------ */
void combination_counts(int order, uint8_t *masks, uint8_t **genotype_combinations, int num_genotype_combinations,
                        int *counts_aff, int *counts_unaff, masks_info info) {
    const int S = info.num_samples_with_padding;
    const int a_pad = info.num_affected_with_padding;
    const int a_scan = round_up16(info.num_affected), u_scan = round_up16(info.num_unaffected);
    for (int rc = 0; rc < info.num_combinations_in_a_row; rc++) {
        const uint8_t *rc_masks = masks + (size_t) rc * order * NUM_GENOTYPES * S;
        for (int c = 0; c < num_genotype_combinations; c++) {
            const uint8_t *cell = genotype_combinations[c];
            int bits = 0;
            for (int s = 0; s < a_scan; s++) bits += popcount8(cell_byte(rc_masks, cell, order, S, s));
            counts_aff[rc * info.num_cell_counts_per_combination + c] = bits / 8;
            bits = 0;
            for (int s = 0; s < u_scan; s++) bits += popcount8(cell_byte(rc_masks, cell, order, S, a_pad + s));
            counts_unaff[rc * info.num_cell_counts_per_combination + c] = bits / 8;
        }
    }
}

/* ------------------------------------------------------------------ *
 * Per-fold TRAINING cell counts.  model.c:131-206: the ANDed mask is ANDed
 * once more with fold f's byte mask (0x01 = sample is in the training part
 * of fold f) and the surviving bits are counted -- no division, because a
 * fold-mask byte has a single bit.  Output index [f][rc][cell].
 * ------------------------------------------------------------------ */
void combination_counts_all_folds(int order, uint8_t *fold_masks, int num_folds,
                                  uint8_t **genotype_permutations, uint8_t *masks, masks_info info,
                                  int *counts_aff, int *counts_unaff) {
    const int S = info.num_samples_with_padding;
    const int C = info.num_cell_counts_per_combination;
    const int R = info.num_combinations_in_a_row;
    const int a_pad = info.num_affected_with_padding;
    const int a_scan = round_up16(info.num_affected), u_scan = round_up16(info.num_unaffected);
    for (int rc = 0; rc < R; rc++) {
        const uint8_t *rc_masks = masks + (size_t) rc * order * NUM_GENOTYPES * S;
        for (int c = 0; c < C; c++) {
            const uint8_t *cell = genotype_permutations[c];
            for (int f = 0; f < num_folds; f++) {
                const uint8_t *fm = fold_masks + (size_t) f * S;
                int na = 0, nu = 0;
                for (int s = 0; s < a_scan; s++) na += popcount8(cell_byte(rc_masks, cell, order, S, s) & fm[s]);
                for (int s = 0; s < u_scan; s++) nu += popcount8(cell_byte(rc_masks, cell, order, S, a_pad + s) & fm[a_pad + s]);
                counts_aff[(size_t) f * R * C + rc * C + c] = na;
                counts_unaff[(size_t) f * R * C + rc * C + c] = nu;
            }
        }
    }
}

/* ------------------------------------------------------------------ *
 * High-risk rule.
 * ------------------------------------------------------------------ */

/* mdr.c:22-42: the scalar double form (kept for completeness; the runner
 * does not call it). */
bool mdr_high_risk_combinations(unsigned int count_affected, unsigned int count_unaffected,
                                unsigned int samples_affected, unsigned int samples_unaffected, void **aux_return_values) {
    (void) aux_return_values;
    if (count_affected == 0 && count_unaffected == 0) return false;
    int total = (int) (count_affected + count_unaffected);
    double ratio = (double) samples_affected / samples_unaffected;
    double prop_unaff = count_unaffected * ratio;
    double reduction = total / (prop_unaff + count_affected);
    double norm_unaff = prop_unaff * reduction;
    double norm_aff = total - norm_unaff;
    return norm_aff >= norm_unaff;
}

/* mdr.c:45-75: six float32 operations per cell, each rounded to nearest,
 * never fused.  `volatile` pins every intermediate to a float32 store so no
 * compiler may keep excess precision or contract a*b+c.  A true flag is
 * returned as INT_MIN, the value _mm_cvtps_epi32 makes of an all-ones mask
 * (mdr.c:70); callers only test it for non-zero (model.c:241).  An empty
 * cell gives 0/0 = NaN and compares false. */
int *mdr_high_risk_combinations2(int *counts_affected, int *counts_unaffected, int num_counts,
                                 unsigned int num_affected, unsigned int num_unaffected, void **aux_return_values) {
    (void) aux_return_values;
    int padded = round_up16(num_counts);
    if (padded == 0) padded = 16;
    int *high_risk = NULL;
    if (posix_memalign((void **) &high_risk, 16, (size_t) padded * sizeof(int)) != 0) return NULL;
    memset(high_risk, 0, (size_t) padded * sizeof(int));

    volatile float ratio = (float) num_affected / (float) num_unaffected;   /* mdr.c:52 */
    for (int i = 0; i < num_counts; i++) {
        volatile float ca = (float) counts_affected[i];
        volatile float cu = (float) counts_unaffected[i];
        volatile float total = ca + cu;                 /* mdr.c:62 */
        volatile float prop_unaff = cu * ratio;         /* mdr.c:64 */
        volatile float denom = prop_unaff + ca;         /* mdr.c:65 */
        volatile float reduction = total / denom;       /* mdr.c:65 */
        volatile float norm_unaff = prop_unaff * reduction;  /* mdr.c:67 */
        volatile float norm_aff = total - norm_unaff;   /* mdr.c:68 */
        high_risk[i] = (norm_aff >= norm_unaff) ? INT_MIN : 0;  /* mdr.c:70 */
    }
    return high_risk;
}

/* model.c:226-255: flatten flags into a list of risky cell indices, grouped
 * by combination; num_risky[c] is incremented (caller zeroes it). */
int *choose_high_risk_combinations2(unsigned int *counts_aff, unsigned int *counts_unaff,
                                    unsigned int num_combinations, unsigned int num_counts_per_combination,
                                    unsigned int num_affected, unsigned int num_unaffected,
                                    unsigned int *num_risky, void **aux_ret,
                                    int *(*test_func)(int *, int *, int, unsigned int, unsigned int, void **)) {
    (void) aux_ret;
    int num_counts = (int) (num_combinations * num_counts_per_combination);
    void *test_aux = NULL;
    int *flags = test_func((int *) counts_aff, (int *) counts_unaff, num_counts, num_affected, num_unaffected, &test_aux);
    int *risky = malloc((size_t) (num_counts > 0 ? num_counts : 1) * sizeof(int));
    int total = 0;
    for (int i = 0; i < num_counts; i++) {
        if (flags[i]) {
            risky[total++] = i % (int) num_counts_per_combination;
            num_risky[i / (int) num_counts_per_combination]++;
        }
    }
    free(flags);
    return risky;
}

/* model.c:278-296 */
risky_combination *risky_combination_new(int order, int comb[], uint8_t **possible_genotypes_combinations,
                                         int num_risky, int *risky_idx, void *aux_info, masks_info info) {
    risky_combination *risky = malloc(sizeof(risky_combination));
    risky->order = order;
    risky->combination = malloc((size_t) order * sizeof(int));
    risky->cross_validation_count = 1;
    risky->accuracy = 0.0;
    risky->genotypes = malloc((size_t) info.num_cell_counts_per_combination * order);
    risky->num_risky_genotypes = num_risky;
    risky->auxiliary_info = aux_info;
    memcpy(risky->combination, comb, (size_t) order * sizeof(int));
    for (int i = 0; i < num_risky; i++) {
        memcpy(risky->genotypes + order * i, possible_genotypes_combinations[risky_idx[i]], (size_t) order);
    }
    return risky;
}

/* model.c:313-317 */
void risky_combination_free(risky_combination *combination) {
    free(combination->combination);
    free(combination->genotypes);
    free(combination);
}

/* ------------------------------------------------------------------ *
 * Confusion matrix.  model.c:337-460.  A sample is predicted "case" when,
 * for at least one risky cell, every SNP's genotype byte equals the cell's
 * genotype.  Padding positions never predict.  TRAINING keeps samples whose
 * fold-mask byte has bit 0 set, TESTING keeps the complement of bit 0
 * (model.c:405-412: mask ^ 1).  matrix = {TP, FN, FP, TN} (model.c:445-453).
 * A model with zero risky cells reads a zero-length array in the reference
 * (undefined); here it predicts nobody.
 * ------------------------------------------------------------------ */
void confusion_matrix(int order, risky_combination *combination, uint8_t **genotypes,
                      uint8_t *fold_masks, enum evaluation_subset subset, int training_size[2], int testing_size[2],
                      masks_info info, unsigned int *matrix) {
    const int S = info.num_samples_with_padding;
    const int a_end = info.num_affected, a_pad = info.num_affected_with_padding;
    const int u_end = a_pad + info.num_unaffected;
    int predicted_aff = 0, predicted_unaff = 0;
    for (int s = 0; s < S; s++) {
        bool padding = (s >= a_end && s < a_pad) || (s >= u_end);
        if (padding) continue;
        unsigned any = 0;
        for (int r = 0; r < combination->num_risky_genotypes; r++) {
            unsigned all = 0xFF;
            for (int j = 0; j < order; j++) {
                all &= (genotypes[j][s] == combination->genotypes[r * order + j]) ? 0xFFu : 0u;
            }
            any |= all;
        }
        unsigned keep = (subset == TRAINING) ? fold_masks[s] : (unsigned) (fold_masks[s] ^ 1);
        int bits = popcount8(any & keep);
        if (s < a_pad) predicted_aff += bits; else predicted_unaff += bits;
    }
    matrix[0] = (unsigned) predicted_aff;
    matrix[2] = (unsigned) predicted_unaff;
    if (subset == TRAINING) {
        matrix[1] = (unsigned) (training_size[0] - predicted_aff);
        matrix[3] = (unsigned) (training_size[1] - predicted_unaff);
    } else {
        matrix[1] = (unsigned) (testing_size[0] - predicted_aff);
        matrix[3] = (unsigned) (testing_size[1] - predicted_unaff);
    }
}

/* model.c:462-479.  Note model.c:465: `if (!function) function = BA;` --
 * CA has enum value 0, so asking for CA yields BA.  Kept. */
double evaluate_model(unsigned int *m, enum eval_function function) {
    double TP = m[0], FN = m[1], FP = m[2], TN = m[3];
    if (!function) function = BA;
    switch (function) {
        case CA:    return (TP + TN) / (TP + FN + TN + FP);
        case BA:    return ((TP / (TP + FN)) + (TN / (TN + FP))) / 2;
        case GAMMA: return (TP * TN - FP * FN) / (TP * TN + FP * FN);
        case TAU_B: return (TP * TN - FP * FN) / sqrt((TP + FN) * (TN + FP) * (TP + FP) * (TN + FN));
        default:    return NAN;
    }
}

/* model.c:324-335: BA is hard-wired (model.c:331). */
double test_model(int order, risky_combination *risky_comb, uint8_t **genotypes,
                  uint8_t *fold_masks, enum evaluation_subset subset, int training_size[2], int testing_size[2],
                  masks_info info, unsigned int *conf_matrix) {
    confusion_matrix(order, risky_comb, genotypes, fold_masks, subset, training_size, testing_size, info, conf_matrix);
    double eval = evaluate_model(conf_matrix, BA);
    risky_comb->accuracy = eval;
    return eval;
}

/* ------------------------------------------------------------------ *
 * Enumerators.  dataset.c:80-201.
 * ------------------------------------------------------------------ */

/* dataset.c:80-82 */
int get_block_stride(size_t block_operations, int order) {
    return (int) ceil(pow((double) block_operations, 1.0 / order));
}

/* dataset.c:84-104: next multiset of block ids in non-decreasing order. */
int get_next_block(int num_blocks, int order, int bc[]) {
    for (int i = order - 1; i >= 0; i--) {
        if (bc[i] + 1 < num_blocks) {
            bc[i]++;
            for (int j = i + 1; j < order; j++) bc[j] = bc[i];
            return 1;
        }
    }
    return 0;
}

/* dataset.c:106-119 */
void get_first_combination_in_block(int order, int first[], int bc[], int stride) {
    first[0] = bc[0] * stride;
    for (int i = 1; i < order; i++) {
        first[i] = bc[i] * stride;
        if (first[i] <= first[i - 1]) first[i] = first[i - 1] + 1;
    }
}

static int imin(int a, int b) { return a < b ? a : b; }

/* dataset.c:131-171.  Restated as is, including the limit formula of
 * dataset.c:137 that makes order >= 3 incomplete across blocks. */
int get_next_combination_in_block(int order, int comb[], int bc[], int stride, int num_variants) {
    int i = order - 1;
    comb[i]++;
    while (i > 0 && comb[i] >= imin((bc[i] + 1) * stride - order + 1 + i, num_variants)) {
        i--;
        comb[i]++;
    }
    if (comb[0] > (bc[0] + 1) * stride - 1 || comb[0] >= num_variants) return 0;
    for (i = i + 1; i < order; i++) {
        comb[i] = (bc[i - 1] == bc[i]) ? comb[i - 1] + 1 : bc[i] * stride;
    }
    if (comb[order - 1] > (bc[order - 1] + 1) * stride - 1 || comb[order - 1] >= num_variants) return 0;
    return 1;
}

/* dataset.c:188-201: odometer over {0,1,2}^order, last digit fastest. */
uint8_t get_next_genotype_combination(int order, uint8_t comb[]) {
    for (int i = order - 1; i >= 0; i--) {
        comb[i]++;
        if (comb[i] < NUM_GENOTYPES) return 1;
        if (comb[0] >= NUM_GENOTYPES) return 0;
        comb[i] = 0;
    }
    return comb[0] < NUM_GENOTYPES;
}

/* dataset.c:173-186 */
uint8_t **get_genotype_combinations(int order, int *num_combinations) {
    *num_combinations = ipow3(order);
    uint8_t **cells = malloc((size_t) *num_combinations * sizeof(uint8_t *));
    cells[0] = calloc((size_t) order, 1);
    int more = 1;
    for (int i = 1; i < *num_combinations && more; i++) {
        cells[i] = malloc((size_t) order);
        memcpy(cells[i], cells[i - 1], (size_t) order);
        more = get_next_genotype_combination(order, cells[i]);
    }
    return cells;
}

/* ------------------------------------------------------------------ *
 * Cross-validation.  cross_validation.c.
 * ------------------------------------------------------------------ */
static long g_shuffle_seed = 0;
void oracle_set_shuffle_seed(long seed) { g_shuffle_seed = seed; }

/* lib/c/src/math/data/array_utils.c:173-188: reseed, then Fisher-Yates from
 * the top with j = (unsigned) (drand48() * (i+1)). */
static void shuffle_ints(int *v, size_t n) {
    if (n <= 1) return;
    srand48(g_shuffle_seed);
    for (size_t i = n - 1; i > 0; i--) {
        size_t j = (unsigned int) (drand48() * (double) (i + 1));
        int t = v[j]; v[j] = v[i]; v[i] = t;
    }
}

static int cmp_int(const void *a, const void *b) { return *(const int *) a - *(const int *) b; }

/* cross_validation.c:4-100.  Shuffle case ids and control ids separately;
 * deal them round-robin (one case and one control per fold per pass, always
 * restarting at fold 0); sort each fold's ids.  sizes[3f..3f+2] =
 * (total, cases, controls) of fold f. */
int **get_k_folds(unsigned int num_aff, unsigned int num_unaff, unsigned int k, unsigned int **sizes) {
    unsigned int n = num_aff + num_unaff;
    int *samples = malloc((size_t) (n ? n : 1) * sizeof(int));
    for (unsigned int i = 0; i < n; i++) samples[i] = (int) i;
    shuffle_ints(samples, num_aff);
    shuffle_ints(samples + num_aff, num_unaff);

    int **folds = malloc(k * sizeof(int *));
    unsigned int *fs = calloc(3 * (size_t) k, sizeof(unsigned int));
    int **aff = malloc(k * sizeof(int *)), **unaff = malloc(k * sizeof(int *));
    for (unsigned int f = 0; f < k; f++) {
        unsigned int cap = n / k + 1;
        aff[f] = malloc(cap * sizeof(int));
        unaff[f] = malloc(cap * sizeof(int));
    }
    unsigned int done_aff = 0, done_unaff = 0;
    while (done_aff + done_unaff < n) {
        for (unsigned int f = 0; f < k && done_aff + done_unaff < n; f++) {
            if (done_aff < num_aff)     aff[f][fs[3 * f + 1]++] = samples[done_aff++];
            if (done_unaff < num_unaff) unaff[f][fs[3 * f + 2]++] = samples[num_aff + done_unaff++];
        }
    }
    for (unsigned int f = 0; f < k; f++) {
        fs[3 * f] = fs[3 * f + 1] + fs[3 * f + 2];
        folds[f] = malloc((size_t) (fs[3 * f] ? fs[3 * f] : 1) * sizeof(int));
        memcpy(folds[f], aff[f], fs[3 * f + 1] * sizeof(int));
        memcpy(folds[f] + fs[3 * f + 1], unaff[f], fs[3 * f + 2] * sizeof(int));
        qsort(folds[f], fs[3 * f], sizeof(int), cmp_int);
        free(aff[f]); free(unaff[f]);
    }
    free(aff); free(unaff); free(samples);
    *sizes = fs;
    return folds;
}

/* cross_validation.c:102-132: byte 1 = sample belongs to the TRAINING part
 * of the fold, 0 = sample is in the fold (testing) or is padding.  Controls
 * are shifted by the case padding. */
uint8_t *get_k_folds_masks(unsigned int num_aff, unsigned int num_unaff, unsigned int k,
                           int **folds, unsigned int *sizes) {
    unsigned int a_pad = (unsigned) round_up16((int) num_aff), u_pad = (unsigned) round_up16((int) num_unaff);
    unsigned int S = a_pad + u_pad, shift = a_pad - num_aff;
    uint8_t *fm = NULL;
    if (posix_memalign((void **) &fm, 16, (size_t) S * k) != 0) return NULL;
    memset(fm, 1, (size_t) S * k);
    for (unsigned int f = 0; f < k; f++) {
        uint8_t *row = fm + (size_t) f * S;
        for (unsigned int j = 0; j < sizes[3 * f + 1]; j++) row[folds[f][j]] = 0;
        for (unsigned int j = sizes[3 * f + 1]; j < sizes[3 * f]; j++) row[folds[f][j] + shift] = 0;
        memset(row + num_aff, 0, a_pad - num_aff);
        memset(row + S - (u_pad - num_unaff), 0, u_pad - num_unaff);
    }
    return fm;
}

/* cross_validation.c:160-195: copy up to `stride` SNP rows of the block into
 * padded rows (cases, zero pad, controls, zero pad).  Rows past the last
 * variant are left untouched, as in the reference (cross_validation.c:167). */
uint8_t *get_genotypes_of_block_coord(int num_variants, int num_samples, masks_info info,
                                      int stride, int block_coord, uint8_t *block_start, uint8_t *genotypes) {
    for (int i = 0; i < stride && block_coord * stride + i < num_variants; i++) {
        uint8_t *dst = genotypes + (size_t) i * info.num_samples_with_padding;
        const uint8_t *src = block_start + (size_t) i * num_samples;
        memcpy(dst, src, (size_t) info.num_affected);
        memset(dst + info.num_affected, 0, (size_t) (info.num_affected_with_padding - info.num_affected));
        memcpy(dst + info.num_affected_with_padding, src + info.num_affected, (size_t) info.num_unaffected);
        memset(dst + info.num_affected_with_padding + info.num_unaffected, 0,
               (size_t) (info.num_unaffected_with_padding - info.num_unaffected));
    }
    return genotypes;
}
