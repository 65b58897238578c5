// epi_host.cpp -- the reference's C host API for `hpg-var-gwas epi` (include/hpgv_epi_compat.h),
// re-provided on top of the CUDA engine (include/hpgv_epi.h).  Host-side plumbing only: option
// handling, the dataset mmap, fold drawing, the repetition loop, the CV-C / CV-A merge of the
// per-fold rankings and the .epi report.  Every count, risk flag, accuracy and per-fold ranking
// comes from the GPU; nothing in this file evaluates a SNP combination.
#include "../../include/hpgv_epi_compat.h"
#include "../../include/hpgv_epi.h"

#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <thread>
#include <vector>

// -------------------------------------------------------------------------------------
// logging in the reference's format (lib/c/src/commons/log.c:19-39): "<ctime>\t<LEVEL>\t<file> [<line>] in <func>(): <msg>",
// INFO and below to stdout, WARNING and above to stderr, everything also to hpg-var-gwas.log when it could be opened
// -------------------------------------------------------------------------------------
namespace {

enum { LV_DEBUG = 1, LV_INFO = 2, LV_WARN = 3, LV_ERROR = 4, LV_FATAL = 5 };
int g_log_level = LV_INFO;
FILE *g_log_file = nullptr;

void log_msg(int level, const char *word, const char *file, int line, const char *func, const char *fmt, ...) {
    if (level < g_log_level) return;
    time_t raw;
    time(&raw);
    char stamp[64];
    snprintf(stamp, sizeof stamp, "%s", ctime(&raw));
    stamp[strcspn(stamp, "\n")] = 0;
    char body[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(body, sizeof body, fmt, ap);
    va_end(ap);
    FILE *os = level < LV_WARN ? stdout : stderr;
    fprintf(os, "%s\t%s\t%s [%i] in %s(): %s", stamp, word, file, line, func, body);
    if (g_log_file) fprintf(g_log_file, "%s\t%s\t%s [%i] in %s(): %s", stamp, word, file, line, func, body);
}
#define LOGI(...) log_msg(LV_INFO, "INFO", "epi_host.cpp", __LINE__, __func__, __VA_ARGS__)
#define LOGW(...) log_msg(LV_WARN, "WARNING", "epi_host.cpp", __LINE__, __func__, __VA_ARGS__)
#define LOGE(...) log_msg(LV_ERROR, "ERROR", "epi_host.cpp", __LINE__, __func__, __VA_ARGS__)
#define LOGF(...) do { log_msg(LV_FATAL, "FATAL", "epi_host.cpp", __LINE__, __func__, __VA_ARGS__); exit(1); } while (0)

long env_long(const char *name, long dflt) {
    const char *v = getenv(name);
    return (v && *v) ? strtol(v, nullptr, 10) : dflt;
}

long clock_seed() {      // what array_shuffle_int seeds with (lib/c/src/math/data/array_utils.c:176-179)
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return (long) tv.tv_usec;
}

}  // namespace

// -------------------------------------------------------------------------------------
// dataset: dataset.c:54-72
// -------------------------------------------------------------------------------------
extern "C" uint8_t *epistasis_dataset_load(int *num_affected, int *num_unaffected, size_t *num_variants, size_t *file_len,
                                           size_t *genotypes_offset, char *filename) {
    const int fd = open(filename, O_RDONLY);
    if (fd < 0) return nullptr;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 12) { close(fd); return nullptr; }
    void *map = mmap(nullptr, (size_t) sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return nullptr;
    const uint8_t *p = static_cast<const uint8_t *>(map);
    const long long len = sb.st_size;
    uint32_t h[4] = {0, 0, 0, 0};
    memcpy(h, p, len >= 16 ? 16 : 12);
    auto fits = [&](long long n, long long a, long long u, long long off, long long slack) {
        return n >= 1 && a >= 1 && u >= 1 && len >= off + n * (a + u) && len <= off + n * (a + u) + slack;
    };
    // current header: 3 x uint32 (dataset.c:58-63); the shipped fixture has size_t + 2 x uint32
    if (fits(h[0], h[1], h[2], 12, 0)) { *num_variants = h[0]; *num_affected = (int) h[1]; *num_unaffected = (int) h[2]; *genotypes_offset = 12; }
    else if (len >= 16 && h[1] == 0 && fits(h[0], h[2], h[3], 16, 7)) { *num_variants = h[0]; *num_affected = (int) h[2]; *num_unaffected = (int) h[3]; *genotypes_offset = 16; }
    else { *num_variants = h[0]; *num_affected = (int) h[1]; *num_unaffected = (int) h[2]; *genotypes_offset = 12; }   // what the reference would read
    *file_len = (size_t) len;
    return static_cast<uint8_t *>(map);
}

extern "C" int epistasis_dataset_close(uint8_t *contents, size_t file_len) {
    if (!contents || file_len == 0) return -1;
    return munmap(contents, file_len);
}

// -------------------------------------------------------------------------------------
// enumerators: dataset.c:80-201 (host only; the GPU search enumerates tuples itself)
// -------------------------------------------------------------------------------------
extern "C" int get_block_stride(size_t block_operations, int order) {
    return (int) std::ceil(std::pow((double) block_operations, 1.0 / order));
}

// next multiset of block ids in non-decreasing order, e.g. (0,1,3) -> (0,2,2) when num_blocks = 4
extern "C" int get_next_block(int num_blocks, int order, int block_coordinates[]) {
    for (int pos = order - 1; pos >= 0; pos--) {
        if (block_coordinates[pos] + 1 >= num_blocks) continue;
        const int v = ++block_coordinates[pos];
        for (int q = pos + 1; q < order; q++) block_coordinates[q] = v;
        return 1;
    }
    return 0;
}

extern "C" void get_first_combination_in_block(int order, int init_coordinates[], int block_coordinates[], int stride) {
    for (int pos = 0; pos < order; pos++) {
        int v = block_coordinates[pos] * stride;
        if (pos > 0 && v <= init_coordinates[pos - 1]) v = init_coordinates[pos - 1] + 1;   // same block: consecutive SNPs
        init_coordinates[pos] = v;
    }
}

// Successor of comb inside the block tuple, with the reference's limits (dataset.c:131-171): exact for
// order 2 (apart from num_variants % stride == 1), incomplete across mixed blocks for order >= 3.
extern "C" int get_next_combination_in_block(int order, int comb[], int block_coordinates[], int stride, int num_variants) {
    int pos = order - 1;
    comb[pos]++;
    for (;;) {
        if (pos == 0) break;
        const int limit = std::min((block_coordinates[pos] + 1) * stride - order + 1 + pos, num_variants);
        if (comb[pos] < limit) break;
        pos--;
        comb[pos]++;
    }
    if (comb[0] > (block_coordinates[0] + 1) * stride - 1 || comb[0] >= num_variants) return 0;
    for (int q = pos + 1; q < order; q++)
        comb[q] = (block_coordinates[q - 1] == block_coordinates[q]) ? comb[q - 1] + 1 : block_coordinates[q] * stride;
    const int last = comb[order - 1];
    if (last > (block_coordinates[order - 1] + 1) * stride - 1 || last >= num_variants) return 0;
    return 1;
}

extern "C" uint8_t get_next_genotype_combination(int order, uint8_t comb[]) {
    // base-3 increment, last SNP fastest; 0 once the first digit overflows
    for (int pos = order - 1; pos >= 0; pos--) {
        if (++comb[pos] < 3) return 1;
        if (pos == 0) return 0;
        comb[pos] = 0;
    }
    return 0;
}

extern "C" uint8_t **get_genotype_combinations(int order, int *num_combinations) {
    int n = 1;
    for (int o = 0; o < order; o++) n *= 3;
    *num_combinations = n;
    uint8_t **cells = (uint8_t **) malloc((size_t) n * sizeof(uint8_t *));
    for (int c = 0; c < n; c++) {
        cells[c] = (uint8_t *) calloc((size_t) order, 1);
        int rem = c;
        for (int pos = order - 1; pos >= 0; pos--) { cells[c][pos] = (uint8_t) (rem % 3); rem /= 3; }
    }
    return cells;
}

// -------------------------------------------------------------------------------------
// folds: cross_validation.c:4-132
// -------------------------------------------------------------------------------------
static int **folds_from_assignment(unsigned A, unsigned U, unsigned k, const int32_t *fos, const uint32_t *sz, unsigned **sizes) {
    int **folds = (int **) malloc(k * sizeof(int *));
    unsigned *out = (unsigned *) calloc(3 * (size_t) k, sizeof(unsigned));
    std::vector<unsigned> fill(k, 0);
    for (unsigned f = 0; f < k; f++) {
        out[3 * f] = sz[3 * f]; out[3 * f + 1] = sz[3 * f + 1]; out[3 * f + 2] = sz[3 * f + 2];
        folds[f] = (int *) malloc(std::max<size_t>(1, sz[3 * f]) * sizeof(int));
    }
    for (unsigned s = 0; s < A + U; s++) folds[fos[s]][fill[fos[s]]++] = (int) s;    // ascending ids = the reference's qsort
    *sizes = out;
    return folds;
}

extern "C" int **get_k_folds(unsigned int A, unsigned int U, unsigned int k, unsigned int **sizes) {
    if (A < k) LOGW("There are less affected samples than folds and they won't be properly distributed\n");
    if (U < k) LOGW("There are less unaffected samples than folds and they won't be properly distributed\n");
    std::vector<int32_t> fos((size_t) A + U);
    std::vector<uint32_t> sz(3 * (size_t) k);
    const long seed = getenv("HPGV_EPI_SEED") ? env_long("HPGV_EPI_SEED", 0) : clock_seed();
    hpgv_epi_k_folds((int) A, (int) U, (int) k, seed, fos.data(), sz.data());
    return folds_from_assignment(A, U, k, fos.data(), sz.data(), sizes);
}

// byte masks, 1 = training sample, 0 = sample of this fold or padding; controls start at A rounded up to 16
extern "C" uint8_t *get_k_folds_masks(unsigned int A, unsigned int U, unsigned int k, int **folds, unsigned int *sizes) {
    const size_t a_pad = 16 * ((A + 15) / 16), u_pad = 16 * ((U + 15) / 16), s_pad = a_pad + u_pad;
    uint8_t *masks = nullptr;
    if (posix_memalign((void **) &masks, 16, std::max<size_t>(16, k * s_pad)) != 0) return nullptr;
    for (unsigned f = 0; f < k; f++) {
        uint8_t *m = masks + f * s_pad;
        memset(m, 0, s_pad);
        memset(m, 1, A);
        memset(m + a_pad, 1, U);
        for (unsigned x = 0; x < sizes[3 * f]; x++) {
            const unsigned s = (unsigned) folds[f][x];
            m[s < A ? s : a_pad + (s - A)] = 0;
        }
    }
    return masks;
}

// -------------------------------------------------------------------------------------
// merge_rankings (epistasis.c:96-153) + report (epistasis_report.c:28-82)
// -------------------------------------------------------------------------------------
extern "C" int hpgv_epi_merge_rankings(int order, int num_folds, int rank_size, const void *models_v, enum evaluation_mode mode,
                                       hpgv_epi_report_row_t *rows, int capacity) {
    const hpgv_epi_model_t *models = static_cast<const hpgv_epi_model_t *>(models_v);
    struct Key {
        int s[3];
        bool operator<(const Key &o) const { return s[0] != o.s[0] ? s[0] < o.s[0] : (s[1] != o.s[1] ? s[1] < o.s[1] : s[2] < o.s[2]); }
    };
    std::map<Key, hpgv_epi_report_row_t> acc;
    for (int f = 0; f < num_folds; f++) {                     // ascending fold: the risky genotypes of the first fold are kept
        for (int r = 0; r < rank_size; r++) {
            const hpgv_epi_model_t &m = models[(size_t) f * rank_size + r];
            if (m.snp[0] < 0) continue;
            const Key key{{m.snp[0], m.snp[1], order == 3 ? m.snp[2] : -1}};
            auto it = acc.find(key);
            if (it == acc.end()) {
                hpgv_epi_report_row_t row;
                memset(&row, 0, sizeof row);
                row.order = order;
                row.snp[0] = m.snp[0]; row.snp[1] = m.snp[1]; row.snp[2] = order == 3 ? m.snp[2] : -1;
                const int cells = order == 2 ? 9 : 27;
                for (int c = 0; c < cells; c++) {
                    if (!((m.risky_mask >> c) & 1u)) continue;
                    int rem = c;
                    for (int pos = order - 1; pos >= 0; pos--) { row.risky_genotypes[row.num_risky][pos] = (uint8_t) (rem % 3); rem /= 3; }
                    row.num_risky++;
                }
                it = acc.emplace(key, row).first;
            }
            it->second.cv_accuracy += m.accuracy;             // NaN propagates like in the reference's sum
            it->second.cv_count += 1;
        }
    }
    std::vector<hpgv_epi_report_row_t> all;
    all.reserve(acc.size());
    for (auto &kv : acc) {
        kv.second.cv_accuracy /= num_folds;                   // divided by the fold count even when cv_count < folds (epistasis.c:142,148)
        all.push_back(kv.second);
    }
    auto nan_last = [](double a, double b) {                  // descending, NaN after every number
        if (std::isnan(a)) return false;
        if (std::isnan(b)) return true;
        return a > b;
    };
    std::stable_sort(all.begin(), all.end(), [&](const hpgv_epi_report_row_t &a, const hpgv_epi_report_row_t &b) {
        if (mode == CV_C && a.cv_count != b.cv_count) return a.cv_count > b.cv_count;
        if (a.cv_accuracy != b.cv_accuracy && !(std::isnan(a.cv_accuracy) && std::isnan(b.cv_accuracy))) return nan_last(a.cv_accuracy, b.cv_accuracy);
        return false;                                         // map order = SNP tuple ascending, kept by the stable sort
    });
    const int n = std::min<int>((int) all.size(), capacity);
    for (int x = 0; x < n; x++) rows[x] = all[x];
    return n;
}

extern "C" void hpgv_epi_write_report(int order, int cv_repetition, enum evaluation_mode mode, enum evaluation_subset subset,
                                      const hpgv_epi_report_row_t *rows, int num_rows, int max_ranking_size, FILE *fd) {
    fprintf(fd, "#CROSS VALIDATION %d\n", cv_repetition + 1);
    fprintf(fd, "#COMBINATIONS OF: %d SNPs\n", order);
    if (mode == CV_C) fprintf(fd, "#EVALUATION MODE: Cross-validation consistency\n");
    else if (mode == CV_A) fprintf(fd, "#EVALUATION MODE: Cross-validation accuracy\n");
    if (subset == TRAINING) fprintf(fd, "#EVALUATION PARTITION: Training\n");
    else if (subset == TESTING) fprintf(fd, "#EVALUATION PARTITION: Testing\n");
    fprintf(fd, "#POSITION\tSNPs\tGENOTYPES\tCV-C\tCV-A\n");
    for (int pos = 0; pos < num_rows && pos < max_ranking_size; pos++) {
        const hpgv_epi_report_row_t &r = rows[pos];
        fprintf(fd, "%d\t(", pos + 1);
        for (int s = 0; s < order - 1; s++) fprintf(fd, " %d,", r.snp[s]);
        fprintf(fd, " %d )\t", r.snp[order - 1]);
        for (int g = 0; g < r.num_risky; g++) {
            fprintf(fd, "(%d-", r.risky_genotypes[g][0]);
            for (int s = 1; s < order - 1; s++) fprintf(fd, "%d, ", r.risky_genotypes[g][s]);
            fprintf(fd, "%d), ", r.risky_genotypes[g][order - 1]);
        }
        fprintf(fd, "%d\t%.3f\n", r.cv_count, r.cv_accuracy);
    }
}

// -------------------------------------------------------------------------------------
// run_epistasis: singlenode/epistasis_runner.c:24-363
// -------------------------------------------------------------------------------------
extern "C" int run_epistasis(shared_options_data_t *shared, epistasis_options_data_t *opt) {
    int num_affected = 0, num_unaffected = 0;
    size_t num_variants = 0, file_len = 0, genotypes_offset = 0;
    uint8_t *input_file = epistasis_dataset_load(&num_affected, &num_unaffected, &num_variants, &file_len, &genotypes_offset, opt->dataset_filename);
    if (!input_file) LOGF("File %s does not exist!\n", opt->dataset_filename);
    const uint8_t *genotypes = input_file + genotypes_offset;
    if (file_len < genotypes_offset + num_variants * (size_t) (num_affected + num_unaffected))
        LOGF("Dataset %s is shorter than its header announces (%zu variants, %d + %d samples)\n", opt->dataset_filename, num_variants, num_affected, num_unaffected);

    const char *outdir = (shared->output_directory && *shared->output_directory) ? shared->output_directory : ".";
    int ret_code = mkdir(outdir, S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH);
    if (ret_code != 0 && errno != EEXIST) LOGF("Can't create output directory: %s\n", outdir);

    const int order = opt->order, num_folds = opt->num_folds;
    const int stride = opt->stride > 0 ? opt->stride : 1;
    LOGI("Combinations of order %d, %d variants per block\n", order, stride);
    LOGI("%zu variants, %d blocks per dimension\n", num_variants, (int) ((num_variants + stride - 1) / stride));
    if (opt->eval_mode == CV_A) LOGI("Using CV-a as ranking criteria\n");
    else if (opt->eval_mode == CV_C) LOGI("Using CV-c as ranking criteria\n");
    else LOGF("Rank criteria not specified! Must be 'count' or 'accu'\n");
    if (order != 2 && order != 3) LOGF("Combinations of order %d are not supported by the GPU engine (2 or 3)\n", order);

    // one engine context per GPU; the combination index space is cut into contiguous ranges.  Everything that touches
    // CUDA goes through the C-ABI of libhpgv_epi.so: this library links no CUDA runtime of its own.
    int ngpu = (int) env_long("HPGV_EPI_GPUS", 1);
    const int have = hpgv_epi_device_count();
    if (have < 1) LOGF("No CUDA device is available: the epistasis engine has no CPU path (%s)\n", hpgv_epi_last_error(nullptr));
    ngpu = std::max(1, std::min(ngpu, have));
    std::vector<hpgv_epi_ctx *> ctx((size_t) ngpu, nullptr);
    const int eval_fn = (int) env_long("HPGV_EPI_EVAL_FUNCTION", HPGV_EVAL_BA);
    for (int g = 0; g < ngpu; g++) {
        if (hpgv_epi_create(g, &ctx[g]) != HPGV_OK) LOGF("GPU %d: %s\n", g, hpgv_epi_last_error(nullptr));
        if (hpgv_epi_set_eval_function(ctx[g], eval_fn) != HPGV_OK) LOGF("GPU %d: %s\n", g, hpgv_epi_last_error(ctx[g]));
        if (hpgv_epi_load_dataset_host(ctx[g], genotypes, (int64_t) num_variants, num_affected, num_unaffected) != HPGV_OK)
            LOGF("GPU %d: %s\n", g, hpgv_epi_last_error(ctx[g]));
    }
    const uint64_t total = hpgv_epi_num_combinations((int64_t) num_variants, order);
    const int rank = std::max(1, std::min(opt->max_ranking_size, HPGV_MAX_RANK));
    const size_t nrec = (size_t) num_folds * rank;
    const bool seeded = getenv("HPGV_EPI_SEED") != nullptr;
    const long seed0 = env_long("HPGV_EPI_SEED", 0);

    std::vector<int32_t> fos((size_t) num_affected + num_unaffected);
    std::vector<hpgv_epi_model_t> models(nrec), part(nrec * (size_t) ngpu);
    std::vector<hpgv_epi_report_row_t> rows(nrec);
    for (int r = 0; r < opt->num_cv_repetitions; r++) {
        LOGI("Running cross-validation #%d...\n", r + 1);
        if (hpgv_epi_k_folds(num_affected, num_unaffected, num_folds, seeded ? seed0 + r : clock_seed(), fos.data(), nullptr) != HPGV_OK)
            LOGF("Cannot draw %d folds\n", num_folds);
        if (ngpu == 1) {
            if (hpgv_epi_set_folds(ctx[0], num_folds, fos.data()) != HPGV_OK) LOGF("GPU 0: %s\n", hpgv_epi_last_error(ctx[0]));
            if (hpgv_epi_search(ctx[0], order, opt->eval_subset, rank, 0, total, models.data()) != HPGV_OK)
                LOGF("GPU 0: %s\n", hpgv_epi_last_error(ctx[0]));
        } else {
            // One host thread per GPU (the role of the MPI ranks of mpi/epistasis_runner.c:129-157): pack, search the GPU's
            // range, bring its F x N models back.  hpgv_epi_search synchronises its own stream before it returns, so
            // nothing here depends on how streams of different contexts order.  GPU 0 then merges (mpi/...:410-452).
            std::vector<int> rcs((size_t) ngpu, HPGV_OK);
            std::vector<std::thread> workers;
            for (int g = 0; g < ngpu; g++) {
                workers.emplace_back([&, g]() {
                    const uint64_t lo = total / ngpu * g + std::min<uint64_t>(g, total % ngpu);
                    const uint64_t hi = total / ngpu * (g + 1) + std::min<uint64_t>(g + 1, total % ngpu);
                    int rc = hpgv_epi_set_folds(ctx[g], num_folds, fos.data());
                    if (rc == HPGV_OK) rc = hpgv_epi_search(ctx[g], order, opt->eval_subset, rank, lo, hi, part.data() + nrec * g);
                    rcs[g] = rc;
                });
            }
            for (auto &w : workers) w.join();
            for (int g = 0; g < ngpu; g++) {
                if (rcs[g] != HPGV_OK) LOGF("GPU %d: %s\n", g, hpgv_epi_last_error(ctx[g]));
                LOGI("Range finished: GPU %d\n", g);
            }
            if (hpgv_epi_merge_host(ctx[0], order, opt->eval_subset, ngpu, num_folds, rank, part.data(), models.data()) != HPGV_OK)
                LOGF("GPU 0: %s\n", hpgv_epi_last_error(ctx[0]));
        }
        const int nrows = hpgv_epi_merge_rankings(order, num_folds, rank, models.data(), opt->eval_mode, rows.data(), (int) rows.size());

        char default_name[32];
        snprintf(default_name, sizeof default_name, "hpg-variant.cv%d.epi", r + 1);
        const char *fname = (shared->output_filename && *shared->output_filename) ? shared->output_filename : default_name;
        const std::string path = std::string(outdir) + "/" + fname;
        LOGI("Output file will be saved in path %s\n", path.c_str());
        FILE *fd = fopen(path.c_str(), "w");
        if (!fd) LOGF("Can't open output file %s\n", path.c_str());
        hpgv_epi_write_report(order, r, opt->eval_mode, opt->eval_subset, rows.data(), nrows, opt->max_ranking_size, fd);
        fclose(fd);
    }
    for (int g = 0; g < ngpu; g++) hpgv_epi_destroy(ctx[g]);
    epistasis_dataset_close(input_file, file_len);
    return ret_code;
}

// -------------------------------------------------------------------------------------
// epistasis(): main_epistasis.c:24-118 + epistasis_options_parsing.c:24-182
// -------------------------------------------------------------------------------------
namespace {

struct EpiOptions {
    std::string dataset, outdir, outfile, eval_subset, eval_mode, eval_function;
    bool has_dataset = false, has_order = false, has_subset = false, has_mode = false;
    long order = 0, stride = 0, num_folds = 0, num_cv = 0, rank = 0, threads = 0, seed = 0, gpus = 0;
    bool has_seed = false;
};

// The subset of the libconfig grammar hpg-variant.conf uses: nested groups `name : { ... } ;`, settings
// `key = value ;` with integer or "string" values, # and // comments.  Collects gwas.epistasis.* settings.
bool read_epistasis_config(const char *path, std::map<std::string, std::string> &out) {
    FILE *fp = fopen(path, "r");
    if (!fp) return false;
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, n);
    fclose(fp);
    std::vector<std::string> scope;
    std::string token, pending_name;
    size_t i = 0;
    auto skip = [&]() {
        for (;;) {
            while (i < text.size() && isspace((unsigned char) text[i])) i++;
            if (i < text.size() && text[i] == '#') { while (i < text.size() && text[i] != '\n') i++; continue; }
            if (i + 1 < text.size() && text[i] == '/' && text[i + 1] == '/') { while (i < text.size() && text[i] != '\n') i++; continue; }
            if (i + 1 < text.size() && text[i] == '/' && text[i + 1] == '*') { i += 2; while (i + 1 < text.size() && !(text[i] == '*' && text[i + 1] == '/')) i++; i += 2; continue; }
            break;
        }
    };
    while (true) {
        skip();
        if (i >= text.size()) break;
        const char c = text[i];
        if (isalpha((unsigned char) c) || c == '_' || c == '*') {
            size_t j = i;
            while (j < text.size() && (isalnum((unsigned char) text[j]) || text[j] == '_' || text[j] == '-' || text[j] == '*')) j++;
            pending_name = text.substr(i, j - i);
            i = j;
            skip();
            if (i < text.size() && (text[i] == ':' || text[i] == '=')) {
                i++;
                skip();
                if (i < text.size() && text[i] == '{') { scope.push_back(pending_name); i++; continue; }
                // scalar value
                std::string value;
                if (i < text.size() && text[i] == '"') {
                    size_t j2 = text.find('"', i + 1);
                    if (j2 == std::string::npos) return false;
                    value = text.substr(i + 1, j2 - i - 1);
                    i = j2 + 1;
                } else {
                    size_t j2 = i;
                    while (j2 < text.size() && text[j2] != ';' && text[j2] != ',' && !isspace((unsigned char) text[j2])) j2++;
                    value = text.substr(i, j2 - i);
                    i = j2;
                }
                std::string full;
                for (auto &s : scope) full += s + ".";
                out[full + pending_name] = value;
            } else {
                return false;
            }
        } else if (c == '}') {
            if (scope.empty()) return false;
            scope.pop_back();
            i++;
        } else if (c == ';' || c == ',') {
            i++;
        } else {
            return false;
        }
    }
    return scope.empty();
}

void usage() {
    printf("Usage: hpg-var-gwas epi -d|--dataset=<file> [--outdir=<str>] --order=<int> [--num-folds=<int>] [--num-cv-runs=<int>]\n"
           "                        [--rank-size=<int>] [--eval-subset=<str>] [--eval-mode=<str>] [--stride=<int>] [-c|--config=<file>]\n"
           "                        [--num-threads=<int>] [--seed=<int>] [--gpus=<int>] [--eval-function=<str>] [--out=<file>] [-l|--log-level=<str>]\n"
           "  -d, --dataset=<file>   Binary dataset used as input\n"
           "  --outdir=<str>         Directory where the output files will be stored\n"
           "  --order=<int>          Number of SNPs to be combined at the same time\n"
           "  --num-folds=<int>      Number of folds in a k-fold cross-validation\n"
           "  --num-cv-runs=<int>    Number of times the k-fold cross-validation process is run\n"
           "  --rank-size=<int>      Number of best models saved\n"
           "  --eval-subset=<str>    Whether to used training (default) or testing partitions when evaluating the best models\n"
           "  --eval-mode=<str>      Whether to rank risky combinations by their CV-C or CV-A (values can be 'count' or 'accu')\n"
           "  --stride=<int>         Number of SNPs per block partition of the dataset (tiling hint; results do not depend on it)\n"
           "  -c, --config=<file>    File that contains the parameters for configuring the application\n"
           "  --num-threads=<int>    Number of threads when a task runs in parallel (host side only)\n"
           "  --seed=<int>           Draw reproducible folds (repetition r uses seed + r); default: microsecond clock\n"
           "  --gpus=<int>           Number of GPUs of this box to shard the combination space over (default 1)\n"
           "  --eval-function=<str>  Function that scores a model: ba (default, what the reference uses), ca, gamma, tau-b\n"
           "  --out=<file>           Name of the report file inside --outdir (default hpg-variant.cv<N>.epi)\n"
           "  -l, --log-level=<str>  Level of the messages to log (debug, info, warn, error, fatal, nothing)\n");
}

}  // namespace

extern "C" int epistasis(int argc, char *argv[], const char *configuration_file) {
    if (argc == 1 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        usage();
        return 0;
    }
    EpiOptions o;
    // Step 1: configuration file (keys gwas.epistasis.*, epistasis_options_parsing.c:39-89)
    if (configuration_file) {
        std::map<std::string, std::string> cfg;
        if (!read_epistasis_config(configuration_file, cfg)) {
            LOGE("Configuration file error: cannot parse %s\n", configuration_file);
            LOGE("Configuration file read with errors\n");
            return 2;   /* CANT_READ_CONFIG_FILE */
        }
        auto num = [&](const char *key, long &dst, const char *what) {
            auto it = cfg.find(std::string("gwas.epistasis.") + key);
            if (it == cfg.end()) LOGW("%s not found in configuration file, must be set via command-line\n", what);
            else dst = strtol(it->second.c_str(), nullptr, 10);
        };
        num("num-threads", o.threads, "Number of threads");
        num("stride", o.stride, "Number of SNPs per block partition");
        num("num-folds", o.num_folds, "Number of folds per k-fold cross-validation");
        num("num-cv-repetitions", o.num_cv, "Number of cross-validation repetitions");
        num("max-ranking-size", o.rank, "Maximum number of best models recorded");
        auto it = cfg.find("gwas.epistasis.evaluation-subset");
        if (it == cfg.end()) LOGW("Evaluation subset not found in configuration file, must be set via command-line\n");
        else { o.eval_subset = it->second; o.has_subset = true; }
        it = cfg.find("gwas.epistasis.evaluation-mode");
        if (it == cfg.end()) LOGW("Evaluation mode not found in configuration file, must be set via command-line\n");
        else { o.eval_mode = it->second; o.has_mode = true; }
    } else {
        // no configuration file at all: the defaults the reference ships in etc/hpg-variant/hpg-variant.conf:36-45
        o.stride = 100; o.num_folds = 10; o.num_cv = 10; o.rank = 50; o.threads = 4;
        o.eval_subset = "training"; o.has_subset = true;
        o.eval_mode = "count"; o.has_mode = true;
    }
    // Step 2: command line overrides the configuration file
    int errors = 0;
    for (int a = 1; a < argc; a++) {
        std::string arg = argv[a], name, value;
        bool has_value = false;
        if (arg.rfind("--", 0) == 0) {
            const size_t eq = arg.find('=');
            name = arg.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) { value = arg.substr(eq + 1); has_value = true; }
        } else if (arg == "-d" || arg == "-c" || arg == "-l") {
            name = arg == "-d" ? "dataset" : (arg == "-c" ? "config" : "log-level");
        } else if (arg.rfind("-d", 0) == 0 || arg.rfind("-c", 0) == 0 || arg.rfind("-l", 0) == 0) {
            name = arg[1] == 'd' ? "dataset" : (arg[1] == 'c' ? "config" : "log-level");
            value = arg.substr(2); has_value = true;
        } else {
            printf("hpg-var-gwas: unexpected argument \"%s\"\n", arg.c_str());
            errors++;
            continue;
        }
        if (!has_value) {
            if (a + 1 >= argc) { printf("hpg-var-gwas: option \"%s\" requires an argument\n", arg.c_str()); errors++; continue; }
            value = argv[++a];
        }
        char *end = nullptr;
        auto as_long = [&](long &dst) {
            dst = strtol(value.c_str(), &end, 10);
            if (end == value.c_str() || *end) { printf("hpg-var-gwas: invalid argument \"%s\" to option --%s\n", value.c_str(), name.c_str()); errors++; }
        };
        if (name == "dataset") { o.dataset = value; o.has_dataset = true; }
        else if (name == "outdir") o.outdir = value;
        else if (name == "config") {}                          // read by the front-end before epistasis() is called (main_gwas.c:58-66)
        else if (name == "out") o.outfile = value;             // shared option (shared_options.c:30); not in the reference's epi table, accepted here
        else if (name == "log-level") {                        // shared option (shared_options.c:58), same
            static const char *levels[] = {"debug", "info", "warn", "error", "fatal"};
            bool ok = false;
            for (int l = 0; l < 5; l++) if (value == levels[l] || value == std::to_string(l + 1)) { g_log_level = l + 1; ok = true; }
            if (value == "nothing") { g_log_level = LV_FATAL + 1; ok = true; }
            if (!ok) { printf("hpg-var-gwas: invalid argument \"%s\" to option --log-level\n", value.c_str()); errors++; }
        }
        else if (name == "eval-function") o.eval_function = value;
        else if (name == "order") { as_long(o.order); o.has_order = true; }
        else if (name == "num-folds") as_long(o.num_folds);
        else if (name == "num-cv-runs") as_long(o.num_cv);
        else if (name == "rank-size") as_long(o.rank);
        else if (name == "stride") as_long(o.stride);
        else if (name == "num-threads") as_long(o.threads);
        else if (name == "eval-subset") { o.eval_subset = value; o.has_subset = true; }
        else if (name == "eval-mode") { o.eval_mode = value; o.has_mode = true; }
        else if (name == "seed") { as_long(o.seed); o.has_seed = true; }
        else if (name == "gpus") as_long(o.gpus);
        else { printf("hpg-var-gwas: invalid option \"%s\"\n", arg.c_str()); errors++; }
    }
    // Like the reference, parse errors are printed and verification decides (parse_epistasis_options prints arg_print_errors
    // and returns, epistasis_options_parsing.c:105-112; main_epistasis.c:62-70 goes on to verify_epistasis_options).
    if (errors) LOGW("%d command-line argument(s) could not be understood (see above); continuing with the rest like the reference\n", errors);

    // Step 3: verification, with the reference's messages and codes (epistasis_options_parsing.c:143-182)
    if (!o.has_dataset) { LOGE("Please specify the dataset file.\n"); return EPISTASIS_DATASET_NOT_SPECIFIED; }
    if (!o.has_order || o.order == 0) { LOGE("Please specify the number of SNPs to be combined at the same time.\n"); return EPISTASIS_ORDER_NOT_SPECIFIED; }
    if (o.num_folds == 0) { LOGE("Please specify the number of folds in a k-fold cross-validation.\n"); return EPISTASIS_FOLDS_NOT_SPECIFIED; }
    if (o.num_cv == 0) { LOGE("Please specify the times the cross-validation will be run.\n"); return EPISTASIS_CV_RUNS_NOT_SPECIFIED; }
    if (!o.has_subset || (o.eval_subset != "training" && o.eval_subset != "testing")) {
        LOGE("Please specify the dataset partition for evaluating the best models (training/testing).\n");
        return EPISTASIS_EVAL_SUBSET_NOT_SPECIFIED;
    }
    if (o.stride == 0) { LOGE("Please specify the number of SNPs per block partition of the dataset.\n"); return EPISTASIS_STRIDE_NOT_SPECIFIED; }

    // Step 4: option structures (main_epistasis.c:134-146: anything but "testing" is TRAINING, anything but "count" is CV_A)
    shared_options_data_t shared;
    memset(&shared, 0, sizeof shared);
    std::string outdir = o.outdir;
    shared.output_directory = outdir.empty() ? nullptr : &outdir[0];
    std::string outfile = o.outfile;
    shared.output_filename = outfile.empty() ? nullptr : &outfile[0];
    shared.num_threads = (int) o.threads;
    epistasis_options_data_t opt;
    memset(&opt, 0, sizeof opt);
    std::string dataset = o.dataset;
    opt.dataset_filename = &dataset[0];
    opt.order = (int) o.order;
    opt.stride = (int) o.stride;
    opt.num_folds = (int) o.num_folds;
    opt.num_cv_repetitions = (int) o.num_cv;
    opt.max_ranking_size = (int) (o.rank > 0 ? o.rank : 50);
    opt.eval_subset = o.eval_subset == "testing" ? TESTING : TRAINING;
    opt.eval_mode = (o.has_mode && o.eval_mode == "count") ? CV_C : CV_A;
    if (o.has_seed) setenv("HPGV_EPI_SEED", std::to_string(o.seed).c_str(), 1);
    if (o.gpus > 0) setenv("HPGV_EPI_GPUS", std::to_string(o.gpus).c_str(), 1);
    if (!o.eval_function.empty()) {
        // extension: which of evaluate_model's functions ranks the models; the reference's runner hard-wires BA (model.c:331)
        static const struct { const char *name; int code; } fns[] = {{"ba", HPGV_EVAL_BA}, {"ca", HPGV_EVAL_CA_TRUE}, {"gamma", HPGV_EVAL_GAMMA}, {"tau-b", HPGV_EVAL_TAU_B}};
        int code = -1;
        for (auto &f : fns) if (o.eval_function == f.name) code = f.code;
        if (code < 0) { LOGE("Unknown evaluation function '%s' (ba, ca, gamma, tau-b)\n", o.eval_function.c_str()); return EPISTASIS_EVAL_MODE_NOT_SPECIFIED; }
        setenv("HPGV_EPI_EVAL_FUNCTION", std::to_string(code).c_str(), 1);
    }

    // Step 5
    run_epistasis(&shared, &opt);
    return 0;
}

// -------------------------------------------------------------------------------------
// leaf functions of model.h / mdr.h / cross_validation.h as adapters over the CUDA engine (function-level parity tests)
// -------------------------------------------------------------------------------------
namespace {

hpgv_epi_ctx *leaf_ctx() {
    static hpgv_epi_ctx *ctx = nullptr;
    if (!ctx && hpgv_epi_create(0, &ctx) != HPGV_OK) LOGF("No CUDA device is available: the epistasis engine has no CPU path (%s)\n", hpgv_epi_last_error(nullptr));
    return ctx;
}
#define LEAF_CK(call) do { if ((call) != HPGV_OK) LOGF("%s: %s\n", #call, hpgv_epi_last_error(leaf_ctx())); } while (0)

// dataset column s (cases first) -> byte position in a padded row
inline int padded_pos(const masks_info &info, int s) { return s < info.num_affected ? s : info.num_affected_with_padding + (s - info.num_affected); }

// uploads `rows` (padded rows, one per SNP) as a data set of rows.size() variants; a single row is doubled (the engine wants >= 2)
void leaf_load(const std::vector<const uint8_t *> &rows, const masks_info &info) {
    const int S = info.num_affected + info.num_unaffected;
    const size_t nv = std::max<size_t>(2, rows.size());
    std::vector<uint8_t> g(nv * (size_t) S);
    for (size_t v = 0; v < nv; v++) {
        const uint8_t *row = rows[std::min(v, rows.size() - 1)];
        for (int s = 0; s < S; s++) g[v * S + s] = row[padded_pos(info, s)];
    }
    LEAF_CK(hpgv_epi_load_dataset_host(leaf_ctx(), g.data(), (int64_t) nv, info.num_affected, info.num_unaffected));
}

// two folds from a byte mask over padded positions: fold 0 = masked out (the fold's own samples), fold 1 = training
void leaf_folds(const uint8_t *fold_mask, const masks_info &info) {
    const int S = info.num_affected + info.num_unaffected;
    std::vector<int32_t> fos((size_t) S, 1);
    if (fold_mask) for (int s = 0; s < S; s++) fos[s] = fold_mask[padded_pos(info, s)] ? 1 : 0;
    LEAF_CK(hpgv_epi_set_folds(leaf_ctx(), 2, fos.data()));
}

// genotype rows back from the byte masks of one SNP ([3][S_pad]): the genotype whose mask is set, 255 where none is
std::vector<uint8_t> row_from_masks(const uint8_t *m, const masks_info &info) {
    const int P = info.num_samples_with_padding;
    std::vector<uint8_t> row((size_t) P, 255);
    for (int p = 0; p < P; p++)
        for (int gt = 0; gt < 3; gt++) if (m[(size_t) gt * P + p]) { row[p] = (uint8_t) gt; break; }
    return row;
}

int cell_of(const uint8_t *perm, int order) {
    int c = 0;
    for (int j = 0; j < order; j++) c = c * 3 + perm[j];
    return c;
}

// TRAINING counts of fold 0 of the current 2-fold partition for the combinations (0..order-1), (order..2 order-1), ...
void leaf_counts(int order, int ncomb, uint8_t **perms, int nperm, int *out_aff, int *out_unaff) {
    const int C = order == 2 ? 9 : 27;
    std::vector<int32_t> combs((size_t) ncomb * order), ca((size_t) ncomb * 2 * C), cu((size_t) ncomb * 2 * C);
    for (int x = 0; x < ncomb * order; x++) combs[x] = x;
    LEAF_CK(hpgv_epi_eval(leaf_ctx(), order, HPGV_SUBSET_TRAINING, ncomb, combs.data(), ca.data(), cu.data(), nullptr, nullptr, nullptr));
    for (int rc = 0; rc < ncomb; rc++)
        for (int p = 0; p < nperm; p++) {
            const int c = cell_of(perms[p], order);
            out_aff[rc * nperm + p] = ca[((size_t) rc * 2 + 0) * C + c];
            out_unaff[rc * nperm + p] = cu[((size_t) rc * 2 + 0) * C + c];
        }
}

}  // namespace

extern "C" void masks_info_init(int order, int num_combinations_in_a_row, int num_affected, int num_unaffected, masks_info *info) {
    info->num_affected = num_affected;
    info->num_unaffected = num_unaffected;
    info->num_affected_with_padding = 16 * ((num_affected + 15) / 16);
    info->num_unaffected_with_padding = 16 * ((num_unaffected + 15) / 16);
    info->num_combinations_in_a_row = num_combinations_in_a_row;
    int cells = 1;
    for (int o = 0; o < order; o++) cells *= 3;
    info->num_cell_counts_per_combination = cells;
    info->num_samples_with_padding = info->num_affected_with_padding + info->num_unaffected_with_padding;
    info->num_masks = 3 * order * info->num_samples_with_padding;
}

extern "C" void set_genotypes_masks(int order, uint8_t **genotypes, int num_combinations, uint8_t *masks, masks_info info) {
    std::vector<const uint8_t *> rows;
    for (int x = 0; x < num_combinations * order; x++) rows.push_back(genotypes[x]);
    leaf_load(rows, info);
    leaf_folds(nullptr, info);
    const size_t P = (size_t) info.num_samples_with_padding;
    for (int x = 0; x < num_combinations * order; x++)     // [combination][snp][genotype][S_pad] is contiguous: 3 * S_pad per SNP
        LEAF_CK(hpgv_epi_unpack_masks(leaf_ctx(), x, masks + (size_t) x * 3 * P));
}

static void leaf_load_from_masks(int order, const uint8_t *masks, const masks_info &info, std::vector<std::vector<uint8_t>> &keep) {
    const size_t P = (size_t) info.num_samples_with_padding;
    const int nrows = info.num_combinations_in_a_row * order;
    keep.clear();
    std::vector<const uint8_t *> rows;
    for (int x = 0; x < nrows; x++) keep.push_back(row_from_masks(masks + (size_t) x * 3 * P, info));
    for (auto &r : keep) rows.push_back(r.data());
    leaf_load(rows, info);
}

extern "C" void combination_counts(int order, uint8_t *masks, uint8_t **genotype_combinations, int num_genotype_combinations,
                                   int *counts_aff, int *counts_unaff, masks_info info) {
    if (order != 2 && order != 3) LOGF("Combinations of order %d are not supported by the GPU engine (2 or 3)\n", order);
    std::vector<std::vector<uint8_t>> keep;
    leaf_load_from_masks(order, masks, info, keep);
    leaf_folds(nullptr, info);                              // nobody is held out: the training table is the whole data set
    leaf_counts(order, info.num_combinations_in_a_row, genotype_combinations, num_genotype_combinations, counts_aff, counts_unaff);
}

extern "C" void combination_counts_all_folds(int order, uint8_t *fold_masks, int num_folds, uint8_t **genotype_permutations, uint8_t *masks,
                                             masks_info info, int *counts_aff, int *counts_unaff) {
    if (order != 2 && order != 3) LOGF("Combinations of order %d are not supported by the GPU engine (2 or 3)\n", order);
    std::vector<std::vector<uint8_t>> keep;
    leaf_load_from_masks(order, masks, info, keep);
    const int C = info.num_cell_counts_per_combination, rows = info.num_combinations_in_a_row;
    for (int f = 0; f < num_folds; f++) {                   // the reference's fold masks need not be a partition: one 2-fold layout each
        leaf_folds(fold_masks + (size_t) f * info.num_samples_with_padding, info);
        leaf_counts(order, rows, genotype_permutations, C, counts_aff + (size_t) f * rows * C, counts_unaff + (size_t) f * rows * C);
    }
}

extern "C" int *mdr_high_risk_combinations2(int *counts_affected, int *counts_unaffected, int num_counts, unsigned int num_affected,
                                            unsigned int num_unaffected, void **aux_return_values) {
    (void) aux_return_values;
    const int padded = 16 * ((num_counts + 15) / 16);
    int *flags = nullptr;
    if (posix_memalign((void **) &flags, 16, std::max<size_t>(16, (size_t) padded * sizeof(int))) != 0) return nullptr;
    memset(flags, 0, (size_t) padded * sizeof(int));
    if (num_counts > 0) {
        LEAF_CK(hpgv_epi_high_risk(leaf_ctx(), counts_affected, counts_unaffected, num_counts, (int) num_affected, (int) num_unaffected, flags));
        for (int x = 0; x < num_counts; x++) flags[x] = flags[x] ? -1 : 0;      // what _mm_cvtps_epi32(_mm_cmpge_ps) leaves: 0x80000000 -> ... any non-zero
    }
    return flags;
}

extern "C" int *choose_high_risk_combinations2(unsigned int *counts_aff, unsigned int *counts_unaff, unsigned int num_combinations,
                                               unsigned int num_counts_per_combination, unsigned int num_affected, unsigned int num_unaffected,
                                               unsigned int *num_risky, void **aux_ret,
                                               int *(*test_func)(int *, int *, int, unsigned int, unsigned int, void **)) {
    (void) aux_ret; (void) test_func;
    const int n = (int) (num_combinations * num_counts_per_combination);
    int *flags = mdr_high_risk_combinations2((int *) counts_aff, (int *) counts_unaff, n, num_affected, num_unaffected, nullptr);
    int *risky = (int *) malloc(std::max<size_t>(1, (size_t) n) * sizeof(int));
    int total = 0;
    for (int x = 0; x < n; x++)
        if (flags[x]) { risky[total++] = x % (int) num_counts_per_combination; num_risky[x / (int) num_counts_per_combination]++; }
    free(flags);
    return risky;
}

extern "C" risky_combination *risky_combination_new(int order, int comb[], uint8_t **possible_genotypes_combinations, int num_risky, int *risky_idx,
                                                    void *aux_info, masks_info info) {
    risky_combination *r = (risky_combination *) malloc(sizeof(risky_combination));
    r->order = order;
    r->combination = (int *) malloc((size_t) order * sizeof(int));
    r->cross_validation_count = 1;
    r->accuracy = 0.0;
    r->genotypes = (uint8_t *) malloc((size_t) info.num_cell_counts_per_combination * order);
    r->num_risky_genotypes = num_risky;
    r->auxiliary_info = aux_info;
    memcpy(r->combination, comb, (size_t) order * sizeof(int));
    for (int x = 0; x < num_risky; x++) memcpy(r->genotypes + (size_t) order * x, possible_genotypes_combinations[risky_idx[x]], (size_t) order);
    return r;
}

extern "C" void risky_combination_free(risky_combination *combination) {
    if (!combination) return;
    free(combination->combination);
    free(combination->genotypes);
    free(combination);
}

extern "C" void confusion_matrix(int order, risky_combination *combination, uint8_t **genotypes, uint8_t *fold_masks, enum evaluation_subset subset,
                                 int training_size[2], int testing_size[2], masks_info info, unsigned int *matrix) {
    if (order != 2 && order != 3) LOGF("Combinations of order %d are not supported by the GPU engine (2 or 3)\n", order);
    std::vector<const uint8_t *> rows;
    for (int j = 0; j < order; j++) rows.push_back(genotypes[j]);
    leaf_load(rows, info);
    leaf_folds(fold_masks, info);
    uint32_t mask[2] = {0, 0};
    for (int x = 0; x < combination->num_risky_genotypes; x++) mask[0] |= 1u << cell_of(combination->genotypes + (size_t) x * order, order);
    mask[1] = mask[0];
    int32_t comb[3] = {0, 1, 2};
    uint32_t conf[8];
    LEAF_CK(hpgv_epi_confusion(leaf_ctx(), order, subset == TRAINING ? HPGV_SUBSET_TRAINING : HPGV_SUBSET_TESTING, 1, comb, mask, conf, nullptr));
    const int *size = subset == TRAINING ? training_size : testing_size;        // fold 0 of the 2-fold layout is "this fold"
    matrix[0] = conf[0]; matrix[2] = conf[2];
    matrix[1] = (unsigned) size[0] - conf[0]; matrix[3] = (unsigned) size[1] - conf[2];
}

extern "C" double evaluate_model(unsigned int *confusion_matrix, enum eval_function function) {
    double v = NAN;
    const uint32_t m[4] = {confusion_matrix[0], confusion_matrix[1], confusion_matrix[2], confusion_matrix[3]};
    LEAF_CK(hpgv_epi_evaluate(leaf_ctx(), (int) function, 1, m, &v));
    return v;
}

extern "C" double test_model(int order, risky_combination *risky_comb, uint8_t **genotypes, uint8_t *fold_masks, enum evaluation_subset subset,
                             int training_size[2], int testing_size[2], masks_info info, unsigned int *conf_matrix) {
    confusion_matrix(order, risky_comb, genotypes, fold_masks, subset, training_size, testing_size, info, conf_matrix);
    return evaluate_model(conf_matrix, BA);                                     // model.c:331
}

extern "C" uint8_t *get_genotypes_of_block_coord(int num_variants, int num_samples, masks_info info, int stride, int block_coord,
                                                 uint8_t *block_start, uint8_t *genotypes) {
    for (int x = 0; x < stride && block_coord * stride + x < num_variants; x++) {
        uint8_t *row = genotypes + (size_t) x * info.num_samples_with_padding;
        const uint8_t *src = block_start + (size_t) x * num_samples;
        memset(row, 0, (size_t) info.num_samples_with_padding);
        memcpy(row, src, (size_t) info.num_affected);
        memcpy(row + info.num_affected_with_padding, src + info.num_affected, (size_t) info.num_unaffected);
    }
    return genotypes;
}

// vcf-tools/vcf2epi/dataset_creator.c:172-223 (the producer's file format, current 12-byte header)
extern "C" int epistasis_dataset_write(const char *filename, const uint8_t *genotypes, size_t num_variants, int num_affected, int num_unaffected) {
    FILE *fp = fopen(filename, "wb");
    if (!fp) return -1;
    const uint32_t hdr[3] = {(uint32_t) num_variants, (uint32_t) num_affected, (uint32_t) num_unaffected};
    const size_t n = num_variants * (size_t) (num_affected + num_unaffected);
    const bool ok = fwrite(hdr, sizeof hdr, 1, fp) == 1 && (n == 0 || fwrite(genotypes, 1, n, fp) == n);
    return (fclose(fp) == 0 && ok) ? 0 : -1;
}

extern "C" uint8_t epistasis_dataset_encode_genotype(int allele1, int allele2, int alleles_missing) {
    if (alleles_missing) return 255;
    if (!allele1 && !allele2) return 0;
    if (allele1 != allele2) return 1;
    return 2;
}

extern "C" int *group_individuals_by_phenotype(uint8_t *phenotypes, int num_affected, int num_unaffected) {
    const int n = num_affected + num_unaffected;
    int *destination = (int *) malloc(std::max<size_t>(1, (size_t) n) * sizeof(int));
    int a = 0, u = num_affected;
    for (int x = 0; x < n; x++) destination[x] = phenotypes[x] ? a++ : u++;
    return destination;
}

extern "C" void hpgv_epi_host_open_log(const char *path) {
    if (g_log_file) fclose(g_log_file);
    g_log_file = path ? fopen(path, "w") : nullptr;
}
