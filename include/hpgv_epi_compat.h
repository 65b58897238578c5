/*
 * hpgv_epi_compat.h -- the reference's own C host API for the `hpg-var-gwas epi` path,
 * re-provided on top of the CUDA engine (hpgv_epi.h).  A maintainer links
 * libhpgv_epi_host.so + libhpgv_epi.so instead of the objects of
 * src/gwas/epistasis/{model,mdr,dataset,cross_validation,epistasis,epistasis_report}.c and
 * singlenode/epistasis_runner.c; main_gwas.c:71 (`epistasis(argc-1, argv+1, config)`) and
 * main_epistasis.c:103 (`run_epistasis(shared, opts)`) keep compiling and linking unchanged.
 *
 * Same names, argument meaning and error behaviour as the reference; every declaration
 * cites the reference declaration it replaces (paths relative to the reference checkout).
 * Struct layouts equal the reference's so objects compiled against its headers interoperate.
 */
#ifndef HPGV_EPI_COMPAT_H
#define HPGV_EPI_COMPAT_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/gwas/epistasis/model.h:72-73 */
enum evaluation_subset { TESTING, TRAINING };
enum evaluation_mode { CV_C, CV_A };

/* src/shared_options.h:95-115 -- only output_directory, output_filename and num_threads are read on
 * this path (singlenode/epistasis_runner.c:39,128; src/hpg_variant_utils.c:302-314). */
typedef struct shared_options_data {
    char *vcf_filename;
    char *ped_filename;
    char *output_filename;
    char *output_directory;
    char *host_url;
    char *version;
    char *species;
    int max_batches;
    int batch_lines;
    int batch_bytes;
    int num_threads;
    int entries_per_thread;
    int compression;
    void *chain;            /* filter_chain*, unused here */
    int log_level;
} shared_options_data_t;

/* src/gwas/epistasis/epistasis.h:78-87 */
typedef struct epistasis_options_data {
    char *dataset_filename;
    int order;
    int stride;                 /* tiling hint only: results do not depend on it */
    int num_folds;
    int num_cv_repetitions;
    int max_ranking_size;
    enum evaluation_subset eval_subset;
    enum evaluation_mode eval_mode;
} epistasis_options_data_t;

/* src/error.h:44-50 and src/error.h CANT_READ_CONFIG_FILE */
#define EPISTASIS_DATASET_NOT_SPECIFIED     210
#define EPISTASIS_ORDER_NOT_SPECIFIED       211
#define EPISTASIS_FOLDS_NOT_SPECIFIED       212
#define EPISTASIS_CV_RUNS_NOT_SPECIFIED     213
#define EPISTASIS_EVAL_SUBSET_NOT_SPECIFIED 214
#define EPISTASIS_EVAL_MODE_NOT_SPECIFIED   215
#define EPISTASIS_STRIDE_NOT_SPECIFIED      216

/* src/gwas/epistasis/epistasis_runner.h:49 (singlenode/epistasis_runner.c:24-363).
 * Loads the dataset, then for each CV repetition: draws stratified folds, runs the exhaustive search on
 * the GPU(s), merges the per-fold rankings (CV-C / CV-A) and writes <outdir>/hpg-variant.cv<r>.epi.
 * Returns what the reference returns: the result of creating the output directory (0, or -1 when it
 * already exists -- singlenode/epistasis_runner.c:39-42,362).  Fatal conditions (missing dataset,
 * uncreatable directory, no GPU) print the reference's message and exit(1) like LOG_FATAL does.
 * Extensions, read from the environment so the struct layouts stay untouched:
 *   HPGV_EPI_SEED=<n>  deterministic folds (repetition r uses seed n + r); default: microsecond clock
 *   HPGV_EPI_GPUS=<n>  shard the combination space over n GPUs of this box (default 1), one host thread per GPU
 *   HPGV_EPI_EVAL_FUNCTION=<code>  HPGV_EVAL_* of hpgv_epi.h that scores the models (default BA, model.c:331) */
int run_epistasis(shared_options_data_t *shared_options_data, epistasis_options_data_t *options_data);

/* src/gwas/main_gwas.h:46 (src/gwas/epistasis/main_epistasis.c:24-118): config file, then the command
 * line (-d/--dataset, --order, --stride, --num-folds, --num-cv-runs, --rank-size, --eval-subset,
 * --eval-mode, --outdir, --config, --num-threads -- the table of epistasis_options_parsing.c:115-140; plus --seed, --gpus,
 * --eval-function and the shared --out / --log-level, which the reference's epi table leaves out), verification with the
 * reference's error codes, then run_epistasis.  Returns 0 like the reference (it ignores run_epistasis's code).
 * configuration_file == NULL applies the defaults the reference ships in etc/hpg-variant/hpg-variant.conf:36-45. */
int epistasis(int argc, char *argv[], const char *configuration_file);

/* src/gwas/epistasis/dataset.h:53-55 (dataset.c:54-72): mmap of the whole file.  Accepts the current
 * 12-byte header and the legacy 16-byte header of test/epistasis_dataset.bin. */
uint8_t *epistasis_dataset_load(int *num_affected, int *num_unaffected, size_t *num_variants, size_t *file_len,
                                size_t *genotypes_offset, char *filename);
int epistasis_dataset_close(uint8_t *contents, size_t file_len);

/* src/gwas/epistasis/dataset.h:61-73 (dataset.c:80-201), host-only enumerators */
int get_block_stride(size_t block_operations, int order);
int get_next_block(int num_blocks, int order, int block_coordinates[]);
void get_first_combination_in_block(int order, int init_coordinates[], int block_coordinates[], int stride);
int get_next_combination_in_block(int order, int comb[], int block_coordinates[], int stride, int num_variants);
uint8_t **get_genotype_combinations(int order, int *num_combinations);
uint8_t get_next_genotype_combination(int order, uint8_t comb[]);

/* src/gwas/epistasis/cross_validation.h:14-16 (cross_validation.c:4-132).  get_k_folds seeds its two
 * shuffles from HPGV_EPI_SEED when set, else from the microsecond clock like the reference.
 * Ownership as in the reference: caller frees every fold, the array and sizes; fold masks with free(). */
int **get_k_folds(unsigned int samples_affected, unsigned int samples_unaffected, unsigned int k, unsigned int **sizes);
uint8_t *get_k_folds_masks(unsigned int num_samples_affected, unsigned int num_samples_unaffected, unsigned int k,
                           int **folds, unsigned int *sizes);

/* ---- leaf functions of the hot path, with the reference's signatures, for function-level parity tests -------------
 * (src/gwas/epistasis/model.h:91-155, mdr.h:37-39, cross_validation.h:14-23).  With them the reference's own
 * unit tests (test/test_epistasis_model.c) link against this library exactly as they link against the reference's
 * objects.  They are ADAPTERS: every count, risk flag, confusion matrix and accuracy below comes from the CUDA engine
 * (hpgv_epi_unpack_masks / hpgv_epi_eval / hpgv_epi_high_risk / hpgv_epi_confusion / hpgv_epi_evaluate on GPU 0), one
 * small upload per call -- they are for tests, the search never goes through them. */

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

/* model.h:84 */
enum eval_function { CA, BA, wBA, GAMMA, TAU_B };

/* model.c:208-219 */
void masks_info_init(int order, int num_combinations_in_a_row, int num_affected, int num_unaffected, masks_info *info);
/* model.c:28-74: byte masks [combination][snp][genotype][S_pad], 0xFF where the sample has the genotype (GPU: pack + unpack) */
void set_genotypes_masks(int order, uint8_t **genotypes, int num_combinations, uint8_t *masks, masks_info info);
/* model.c:76-129 (whole data set) and model.c:131-206 (per fold, fold_masks = 1 for a training sample); counts come back as
 * [fold][combination in the row][cell] like the reference's */
void combination_counts(int order, uint8_t *masks, uint8_t **genotype_combinations, int num_genotype_combinations,
                        int *counts_aff, int *counts_unaff, masks_info info);
void combination_counts_all_folds(int order, uint8_t *fold_masks, int num_folds, uint8_t **genotype_permutations, uint8_t *masks,
                                  masks_info info, int *counts_aff, int *counts_unaff);
/* mdr.c:45-75: 0 / -1 flags like _mm_cmpge_ps leaves them, 16-byte aligned (free with free()) */
int *mdr_high_risk_combinations2(int *counts_affected, int *counts_unaffected, int num_counts, unsigned int num_affected,
                                 unsigned int num_unaffected, void **aux_return_values);
/* model.c:226-255 with the pointer-correct callback type; the device rule decides whatever test_func is */
int *choose_high_risk_combinations2(unsigned int *counts_aff, unsigned int *counts_unaff, unsigned int num_combinations,
                                    unsigned int num_counts_per_combination, unsigned int num_affected, unsigned int num_unaffected,
                                    unsigned int *num_risky, void **aux_ret,
                                    int *(*test_func)(int *, int *, int, unsigned int, unsigned int, void **));
/* model.c:278-296, 313-317 */
risky_combination *risky_combination_new(int order, int comb[], uint8_t **possible_genotypes_combinations, int num_risky, int *risky_idx,
                                         void *aux_info, masks_info info);
void risky_combination_free(risky_combination *combination);
/* model.c:337-460, 462-479, 324-335 */
void confusion_matrix(int order, risky_combination *combination, uint8_t **genotypes, uint8_t *fold_masks, enum evaluation_subset subset,
                      int training_size[2], int testing_size[2], masks_info info, unsigned int *matrix);
double evaluate_model(unsigned int *confusion_matrix, enum eval_function function);
double test_model(int order, risky_combination *risky_comb, uint8_t **genotypes, uint8_t *fold_masks, enum evaluation_subset subset,
                  int training_size[2], int testing_size[2], masks_info info, unsigned int *conf_matrix);
/* cross_validation.c:160-195: the padded copy of a block's rows (the GPU engine packs bit planes instead and never calls it) */
uint8_t *get_genotypes_of_block_coord(int num_variants, int num_samples, masks_info info, int stride, int block_coord,
                                      uint8_t *block_start, uint8_t *genotypes);

/* vcf-tools/vcf2epi/dataset_creator.c:172-223: writes a data set in the current format (3 x uint32 header, then
 * variant-major genotype bytes, cases first); returns 0, or -1 when the file cannot be written */
int epistasis_dataset_write(const char *filename, const uint8_t *genotypes, size_t num_variants, int num_affected, int num_unaffected);
/* dataset_creator.c:255-265: the byte a sample's GT becomes (0 hom-ref, 1 het, 2 hom-alt, 255 when the alleles are missing) */
uint8_t epistasis_dataset_encode_genotype(int allele1, int allele2, int alleles_missing);
/* dataset_creator.c:302-320: column of every sample in the data set, cases first then controls, input order kept inside a class
 * (phenotypes[i] != 0 = affected); malloc'd, caller frees */
int *group_individuals_by_phenotype(uint8_t *phenotypes, int num_affected, int num_unaffected);

/* One ranked row of a repetition's report = what merge_rankings (epistasis.c:96-153) leaves in its heap. */
typedef struct {
    double cv_accuracy;          /* sum of the fold accuracies / num_folds (epistasis.c:142,148) */
    int cv_count;                /* number of folds whose top-N held the combination */
    int order;
    int snp[3];
    int num_risky;
    uint8_t risky_genotypes[27][3];
} hpgv_epi_report_row_t;

/* merge_rankings + the ordering of epistasis_report (epistasis.c:96-153, epistasis_report.c:49-82) on the
 * per-fold top-N lists returned by hpgv_epi_search: rows sorted for eval_mode (CV_A: accuracy descending;
 * CV_C: count descending then accuracy descending; ties by SNP tuple ascending).  Returns the number of
 * rows written (<= capacity). */
int hpgv_epi_merge_rankings(int order, int num_folds, int rank_size, const void *models /* hpgv_epi_model_t[F][rank] */,
                            enum evaluation_mode mode, hpgv_epi_report_row_t *rows, int capacity);

/* epistasis_report (epistasis_report.c:28-82), same text format, at most max_ranking_size rows */
void hpgv_epi_write_report(int order, int cv_repetition, enum evaluation_mode mode, enum evaluation_subset subset,
                           const hpgv_epi_report_row_t *rows, int num_rows, int max_ranking_size, FILE *fd);

/* The reference logs to stdout/stderr and to hpg-var-gwas.log (main_gwas.c:32).  Opens (path) or closes (NULL) the log file. */
void hpgv_epi_host_open_log(const char *path);

#ifdef __cplusplus
}
#endif
#endif
