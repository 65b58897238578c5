/*
 * hpgv_epi.h -- flat C-ABI of the B200-native epistasis (MDR + k-fold CV) engine.
 *
 * This is the layer the reference's C host code binds to (there is no plugin
 * interface in hpg-variant; the boundary is link-level).  Each
 * entry point names the reference code it replaces; paths are relative to the
 * reference checkout, directory src/gwas/epistasis unless noted.  The
 * reference-signature wrappers (run_epistasis, epistasis, get_k_folds, ...)
 * built on top of this header are declared in hpgv_epi_compat.h.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function
 * returns HPGV_OK (0) or a negative HPGV_E_* code; hpgv_epi_last_error()
 * gives the message.  Nothing in here calls exit() (the reference does:
 * LOG_FATAL, lib/c/src/commons/log.h:103-109).  There is NO CPU fallback: if
 * no CUDA device is usable hpgv_epi_create() fails.
 */
#ifndef HPGV_EPI_H
#define HPGV_EPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPGV_OK              0
#define HPGV_E_CUDA         -1   /* CUDA runtime error (message has the cudaError string) */
#define HPGV_E_ARG          -2   /* invalid argument */
#define HPGV_E_STATE        -3   /* call order violated (e.g. run before set_folds) */
#define HPGV_E_UNSUPPORTED  -4   /* shape outside what the kernels were built for */
#define HPGV_E_IO           -5   /* dataset file missing / malformed */
#define HPGV_E_NOMEM        -6

#define HPGV_MAX_FOLDS      32
#define HPGV_MAX_RANK       4096

/* enum evaluation_subset, model.h:72 (TESTING = 0, TRAINING = 1) */
#define HPGV_SUBSET_TESTING  0
#define HPGV_SUBSET_TRAINING 1

/* enum eval_function, model.h:84 { CA, BA, wBA, GAMMA, TAU_B } and evaluate_model, model.c:462-479.
 * HPGV_EVAL_CA is what the reference EXECUTES for code 0: `if (!function) function = BA` (model.c:465-467) turns it
 * into BA, which is also why test/test_epistasis_model.c:518-519 passes.  HPGV_EVAL_CA_TRUE is the documented formula
 * (TP+TN)/(TP+FN+TN+FP) as an extension.  wBA is a TODO in the reference and is rejected (HPGV_E_UNSUPPORTED). */
#define HPGV_EVAL_CA       0
#define HPGV_EVAL_BA       1
#define HPGV_EVAL_WBA      2
#define HPGV_EVAL_GAMMA    3
#define HPGV_EVAL_TAU_B    4
#define HPGV_EVAL_CA_TRUE  5

/* One ranked model = one (SNP combination, fold) evaluation.  Replaces
 * `risky_combination` (model.h:49-57): accuracy, combination[], and the risky
 * genotype tuples, here as a bit mask over the 3^order cells in the order
 * get_genotype_combinations() produces them (dataset.c:173-186: last SNP
 * fastest, cell = sum g_j * 3^(order-1-j)).  conf = {TP, FN, FP, TN}
 * (model.c:445-453).  Unused slots have accuracy = NaN and snp = -1. */
typedef struct {
    double   accuracy;      /* balanced accuracy, model.c:473 */
    int32_t  snp[3];        /* ascending 0-based variant indices; snp[2] = -1 for order 2 */
    uint32_t risky_mask;
    uint32_t conf[4];
} hpgv_epi_model_t;         /* 40 bytes */

typedef struct hpgv_epi_ctx hpgv_epi_ctx;

/* ---- context ---------------------------------------------------------------- */

/* One context = one GPU.  device < 0 selects the current CUDA device. */
int  hpgv_epi_create(int device, hpgv_epi_ctx **out);
void hpgv_epi_destroy(hpgv_epi_ctx *ctx);
const char *hpgv_epi_last_error(const hpgv_epi_ctx *ctx);   /* ctx may be NULL: last create() error */

/* Work is enqueued on `cuda_stream` (a cudaStream_t; NULL = legacy default
 * stream).  Host-pointer outputs are synchronised before the call returns. */
int hpgv_epi_set_stream(hpgv_epi_ctx *ctx, void *cuda_stream);

/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t hpgv_epi_launch_count(const hpgv_epi_ctx *ctx);

/* Evaluation function that ranks the models (and fills `accuracy`) in every later search / eval of this context.
 * Default HPGV_EVAL_BA: the only one the reference's runner ever uses (model.c:331 hard-wires BA); the others are
 * SURVEY 8(f)3.  A model whose value is NaN (0/0) ranks last, like in add_to_model_ranking (model.c:491). */
int hpgv_epi_set_eval_function(hpgv_epi_ctx *ctx, int eval_function);

/* ---- dataset: replaces epistasis_dataset_load/close (dataset.c:54-72) ---------
 * Genotypes are variant-major bytes, cases (affected) first then controls:
 * genotypes[v * (A+U) + s] in {0,1,2}, anything else (255) = missing
 * (vcf-tools/vcf2epi/dataset_creator.c:255-265,302-320). */
int hpgv_epi_load_dataset_host(hpgv_epi_ctx *ctx, const uint8_t *genotypes,
                               int64_t num_variants, int num_affected, int num_unaffected);
/* same, bytes already in device memory of ctx's GPU (not copied, must outlive the dataset) */
int hpgv_epi_load_dataset_device(hpgv_epi_ctx *ctx, const uint8_t *d_genotypes,
                                 int64_t num_variants, int num_affected, int num_unaffected);
/* file with the current 12-byte header (3 x uint32: variants, affected, unaffected;
 * dataset.c:58-63) or the legacy 16-byte header of the shipped fixture. */
int hpgv_epi_load_dataset_file(hpgv_epi_ctx *ctx, const char *path);
int hpgv_epi_dataset_dims(const hpgv_epi_ctx *ctx, int64_t *num_variants, int *num_affected, int *num_unaffected);

/* ---- folds: replaces get_k_folds_masks (cross_validation.c:102-132) and the
 * per-combination set_genotypes_masks (model.c:28-74).  fold_of_sample[s] in
 * [0, num_folds) is the fold whose TESTING part holds sample s (s in dataset
 * column order).  Builds the (class, fold)-segmented bit planes on the GPU.
 * Asynchronous on the context's stream: the call returns once the pack kernel is
 * enqueued (fold_of_sample is copied before it returns); errors of the kernel
 * surface at the next synchronising call. */
int hpgv_epi_set_folds(hpgv_epi_ctx *ctx, int num_folds, const int32_t *fold_of_sample);

/* Stratified k-fold assignment with the reference's algorithm
 * (cross_validation.c:4-100 + lib/c/src/math/data/array_utils.c:173-188) and an
 * explicit seed instead of the microsecond clock.  Host only.
 * sizes (may be NULL) receives 3*k values: total, cases, controls per fold. */
int hpgv_epi_k_folds(int num_affected, int num_unaffected, int num_folds, long seed,
                     int32_t *fold_of_sample, uint32_t *sizes);

/* ---- the search: replaces the block loop of run_epistasis
 * (singlenode/epistasis_runner.c:128-307) = process_set_of_combinations
 * (epistasis.c:4-93) over every combination, i.e. combination_counts_all_folds,
 * choose_high_risk_combinations2 / mdr_high_risk_combinations2, confusion_matrix,
 * evaluate_model(BA) and add_to_model_ranking, for all folds.
 *
 * Combinations are numbered in lexicographic order of their ascending SNP
 * tuples; [first, last) selects a contiguous range (multi-GPU sharding), last
 * = UINT64_MAX means "to the end".  out receives num_folds x rank_size models,
 * fold-major, each fold's list sorted best first by (accuracy descending, SNP
 * tuple ascending) -- the canonical order that replaces the reference's
 * heap-internal tie-breaking.  out may be a host pointer
 * (hpgv_epi_search) or a device pointer (hpgv_epi_search_device). */
int hpgv_epi_search(hpgv_epi_ctx *ctx, int order, int eval_subset, int rank_size,
                    uint64_t first, uint64_t last, hpgv_epi_model_t *out);
int hpgv_epi_search_device(hpgv_epi_ctx *ctx, int order, int eval_subset, int rank_size,
                           uint64_t first, uint64_t last, hpgv_epi_model_t *d_out);

/* Merge `num_lists` per-rank results (each num_folds x rank_size, as produced by
 * hpgv_epi_search_device and all-gathered) into one num_folds x rank_size result
 * on the GPU: the cross-GPU step that replaces the MPI tree merge
 * (mpi/epistasis_runner.c:410-452).  d_lists is [num_lists][num_folds][rank_size]. */
int hpgv_epi_merge_device(hpgv_epi_ctx *ctx, int order, int eval_subset, int num_lists, int num_folds, int rank_size,
                          const hpgv_epi_model_t *d_lists, hpgv_epi_model_t *d_out);

/* The same with HOST pointers in and out (a single-process caller that drives several GPUs, e.g. run_epistasis with
 * HPGV_EPI_GPUS > 1, hands the per-GPU results of hpgv_epi_search to the context that merges). */
int hpgv_epi_merge_host(hpgv_epi_ctx *ctx, int order, int eval_subset, int num_lists, int num_folds, int rank_size,
                        const hpgv_epi_model_t *lists, hpgv_epi_model_t *out);

/* Number of usable CUDA devices (0 when there is none or the runtime fails: hpgv_epi_last_error(NULL) says why). */
int hpgv_epi_device_count(void);

uint64_t hpgv_epi_num_combinations(int64_t num_variants, int order);

/* ---- parity hooks: per-combination dump for an explicit list of combinations.
 * combs is [num_combs][order] ascending indices.  Any output may be NULL.
 *   counts_aff/unaff [num_combs][F][3^order]  TRAINING counts = combination_counts_all_folds (model.c:131-206)
 *   risky_mask       [num_combs][F]           mdr_high_risk_combinations2 (mdr.c:45-75)
 *   conf             [num_combs][F][4]        confusion_matrix (model.c:337-460) for eval_subset
 *   accuracy         [num_combs][F]           evaluate_model(BA) (model.c:473)
 * All pointers are host pointers. */
int hpgv_epi_eval(hpgv_epi_ctx *ctx, int order, int eval_subset, int64_t num_combs, const int32_t *combs,
                  int32_t *counts_aff, int32_t *counts_unaff, uint32_t *risky_mask, uint32_t *conf, double *accuracy);

/* confusion_matrix (model.c:337-460) for risky cells GIVEN by the caller (the reference passes them in a
 * risky_combination): risky_mask_in is [num_combs][F], conf [num_combs][F][4], accuracy [num_combs][F]. */
int hpgv_epi_confusion(hpgv_epi_ctx *ctx, int order, int eval_subset, int64_t num_combs, const int32_t *combs,
                       const uint32_t *risky_mask_in, uint32_t *conf, double *accuracy);

/* The device's high-risk rule (the function the search kernels call; mdr_high_risk_combinations2, mdr.c:45-75) on n
 * explicit (affected, unaffected) count pairs with dataset sizes (num_affected, num_unaffected): flags[i] = 0/1. */
int hpgv_epi_high_risk(hpgv_epi_ctx *ctx, const int32_t *counts_aff, const int32_t *counts_unaff, int64_t n,
                       int num_affected, int num_unaffected, int32_t *flags);

/* evaluate_model (model.c:462-479) on the device, the function the search kernels call: conf is [n][4] = {TP, FN, FP, TN}. */
int hpgv_epi_evaluate(hpgv_epi_ctx *ctx, int eval_function, int64_t n, const uint32_t *conf, double *values);

/* Unpack the GPU bit planes of one variant back to the reference's byte masks:
 * out is [3][S_pad] with 0xFF where genotype == g (set_genotypes_masks layout,
 * model.c:28-74; S_pad = 16*ceil(A/16) + 16*ceil(U/16)).  Parity hook for the packer. */
int hpgv_epi_unpack_masks(hpgv_epi_ctx *ctx, int64_t variant, uint8_t *out);

/* Whole path behind one call, HOST buffers in, HOST result out (what bench.py's
 * `e2e` times): load_dataset_host + set_folds + search. */
int hpgv_epi_run_host(hpgv_epi_ctx *ctx, const uint8_t *genotypes, int64_t num_variants,
                      int num_affected, int num_unaffected, int num_folds, const int32_t *fold_of_sample,
                      int order, int eval_subset, int rank_size, uint64_t first, uint64_t last,
                      hpgv_epi_model_t *out);

/* Introspection for benches/docs: layout chosen by set_folds. */
typedef struct {
    int num_folds, num_segments, num_blocks, block_words;   /* block = block_words 32-bit words of one (class, fold) segment;
                                                              * 3 = tri layout: three words plus a 4-bit tail shared eight to a word */
    int count_bits;                                          /* 8 or 16: per-segment counter width in shared memory */
    int64_t plane_bytes;                                     /* bytes of the packed planes in HBM */
    int words_per_class_row;                                 /* W of SURVEY 8(d): ceil(A/32)+ceil(U/32) */
    int num_chunks, chunk_blocks, row_words;                 /* planes are [chunk][snp][row_words], chunk = chunk_blocks blocks */
} hpgv_epi_layout_t;
int hpgv_epi_layout(const hpgv_epi_ctx *ctx, hpgv_epi_layout_t *out);

/* Device time (CUDA events on the context's stream) of the most recent search kernel launch --
 * the dominant kernel -- and its grid size.  Blocks until that launch has finished. */
int hpgv_epi_last_search_ms(hpgv_epi_ctx *ctx, float *ms, int *grid);

/* The same for the last n search launches (n <= 32), oldest first, without a host synchronisation per
 * launch: a caller that times a loop of steps enqueues them back to back and reads the durations once.
 * Returns the number of durations written (fewer than n when fewer launches were made). */
int hpgv_epi_search_times(hpgv_epi_ctx *ctx, int n, float *ms);

/* Development counters of the most recent search launched with HPGV_DEBUG_COUNTERS=1 in the environment (candidates that
 * passed the pre-filter / reached a list / were stored, lock spins, per-CTA start and end clocks; epi_types.h SearchArgs::dbg).
 * Returns the number of values written. */
int hpgv_epi_debug_counters(hpgv_epi_ctx *ctx, uint64_t *out, int n);

/* POPC / LOP3 pipe micro-benchmark (roofline denominator):
 * runs `iters` dependent-free rounds per thread on every SM and returns
 * measured 32-bit ops per second.  kind: 0 = POPC, 1 = LOP3, 2 = POPC+LOP3 mix. */
int hpgv_epi_pipe_peak(hpgv_epi_ctx *ctx, int kind, int iters, double *ops_per_second);

#ifdef __cplusplus
}
#endif
#endif
