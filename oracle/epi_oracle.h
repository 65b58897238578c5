/*
 * epi_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the algorithms on hpg-variant's `hpg-var-gwas epi`
 * hot path (MDR + k-fold cross-validation).  Every function cites the
 * reference file:line whose behaviour it restates (paths relative to the
 * reference checkout, src/gwas/epistasis unless noted).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product path (hpg_variant_b200)
 * never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  The reference's own golden vectors
 * (test/test_epistasis_model.c, 7 tests) are compiled against this library
 * by oracle/Makefile (target check-oracle) and pass; the same driver
 * (epi_driver.c) is linked once against this file and once against the
 * reference's own sources (oracle/_ref/libhpgref.so) and tests/ compare the
 * two on seeded inputs bit for bit.
 *
 * The struct layouts and signatures below intentionally equal the
 * reference's model.h:49-70, so that the reference's unit tests link against
 * this library unchanged.
 */
#ifndef EPI_ORACLE_H
#define EPI_ORACLE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#define NUM_GENOTYPES 3

/* model.h:49-57 */
typedef struct {
    double accuracy;
    int order;
    int num_risky_genotypes;
    int cross_validation_count;
    uint8_t *genotypes;
    int *combination;
    void *auxiliary_info;
} risky_combination;

/* model.h:60-70 */
typedef struct {
    int num_affected;
    int num_unaffected;
    int num_affected_with_padding;
    int num_unaffected_with_padding;
    int num_samples_with_padding;
    int num_masks;
    int num_combinations_in_a_row;
    int num_cell_counts_per_combination;
    uint8_t *masks;
} masks_info;

/* model.h:72-84 */
enum evaluation_subset { TESTING, TRAINING };
enum evaluation_mode { CV_C, CV_A };
enum eval_function { CA, BA, wBA, GAMMA, TAU_B };

/* model.c:208-219 */
void masks_info_init(int order, int num_combinations_in_a_row, int num_affected, int num_unaffected, masks_info *info);
/* model.c:28-74 */
void set_genotypes_masks(int order, uint8_t **genotypes, int num_combinations, uint8_t *masks, masks_info info);
/* model.c:76-129 */
void combination_counts(int order, uint8_t *masks, uint8_t **genotype_combinations, int num_genotype_combinations,
                        int *counts_aff, int *counts_unaff, masks_info info);
/* model.c:131-206 */
void combination_counts_all_folds(int order, uint8_t *fold_masks, int num_folds,
                                  uint8_t **genotype_permutations, uint8_t *masks, masks_info info,
                                  int *counts_aff, int *counts_unaff);

/* mdr.c:22-42 (scalar double rule, unused by the runner) */
bool mdr_high_risk_combinations(unsigned int count_affected, unsigned int count_unaffected,
                                unsigned int samples_affected, unsigned int samples_unaffected, void **aux_return_values);
/* mdr.c:45-75 (float32 rule, the one the runner uses) */
int *mdr_high_risk_combinations2(int *counts_affected, int *counts_unaffected, int num_counts,
                                 unsigned int num_affected, unsigned int num_unaffected, void **aux_return_values);
/* model.c:226-255, with the pointer-correct callback prototype */
int *choose_high_risk_combinations2(unsigned int *counts_aff, unsigned int *counts_unaff,
                                    unsigned int num_combinations, unsigned int num_counts_per_combination,
                                    unsigned int num_affected, unsigned int num_unaffected,
                                    unsigned int *num_risky, void **aux_ret,
                                    int *(*test_func)(int *, int *, int, unsigned int, unsigned int, void **));

/* model.c:278-296, 313-317 */
risky_combination *risky_combination_new(int order, int comb[], uint8_t **possible_genotypes_combinations,
                                         int num_risky, int *risky_idx, void *aux_info, masks_info info);
void risky_combination_free(risky_combination *combination);

/* model.c:324-335, 337-460, 462-479 */
double test_model(int order, risky_combination *risky_comb, uint8_t **genotypes,
                  uint8_t *fold_masks, enum evaluation_subset subset, int training_size[2], int testing_size[2],
                  masks_info info, unsigned int *conf_matrix);
void confusion_matrix(int order, risky_combination *combination, uint8_t **genotypes,
                      uint8_t *fold_masks, enum evaluation_subset subset, int training_size[2], int testing_size[2],
                      masks_info info, unsigned int *matrix);
double evaluate_model(unsigned int *confusion_matrix, enum eval_function function);

/* dataset.c:80-201 */
int get_block_stride(size_t block_operations, int order);
int get_next_block(int num_blocks, int order, int block_coordinates[]);
void get_first_combination_in_block(int order, int init_coordinates[], int block_coordinates[], int stride);
int get_next_combination_in_block(int order, int comb[], int block_coordinates[], int stride, int num_variants);
uint8_t **get_genotype_combinations(int order, int *num_combinations);
uint8_t get_next_genotype_combination(int order, uint8_t comb[]);

/* cross_validation.c:4-100, 102-132, 160-195.  get_k_folds draws its shuffles
 * from srand48(seed)+drand48 exactly like lib/c/src/math/data/array_utils.c:173-188,
 * where `seed` is the microsecond clock in the reference; the
 * oracle takes it from oracle_set_shuffle_seed() so runs are reproducible. */
void oracle_set_shuffle_seed(long seed);
int **get_k_folds(unsigned int num_samples_affected, unsigned int num_samples_unaffected, unsigned int k, unsigned int **sizes);
uint8_t *get_k_folds_masks(unsigned int num_samples_affected, unsigned int num_samples_unaffected, unsigned int k,
                           int **folds, unsigned int *sizes);
uint8_t *get_genotypes_of_block_coord(int num_variants, int num_samples, masks_info info,
                                      int stride, int block_coord, uint8_t *block_start, uint8_t *genotypes);

#endif
