// epi_capi.cu -- implementation of include/hpgv_epi.h on top of the kernels in
// epi_kernels.cuh.  Host side: context, dataset upload, fold layout, work-list
// construction, kernel dispatch.  No CPU compute path exists in this file: every
// count, risk flag, accuracy and ranking decision is made by a CUDA kernel.
#include "../../include/hpgv_epi.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "epi_kernels.cuh"
#include "epi_launch.h"

using namespace hpgv;

static_assert(sizeof(hpgv_epi_model_t) == 40, "hpgv_epi_model_t must be 40 bytes");
static_assert(sizeof(ModelOut) == sizeof(hpgv_epi_model_t), "device/host model records differ");
static_assert(kMaxFolds == HPGV_MAX_FOLDS && kMaxRank == HPGV_MAX_RANK, "limits out of sync with hpgv_epi.h");

#include <unistd.h>

namespace {

thread_local std::string g_create_error;

// cudaGetDeviceCount with a few retries: right after another process of the same box has torn its context down the first
// call of a fresh process occasionally fails with cudaErrorInitializationError (seen on multi-GPU boxes)
cudaError_t device_count_retry(int *n) {
    cudaError_t e = cudaSuccess;
    for (int attempt = 0; attempt < 5; attempt++) {
        e = cudaGetDeviceCount(n);
        if (e == cudaSuccess) return e;
        (void) cudaGetLastError();
        usleep(200 * 1000);
    }
    return e;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct hpgv_epi_ctx {
    int device = 0;
    int num_sms = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int eval_fn = kEvalBA;                // what ranks the models (hpgv_epi_set_eval_function); BA like the reference runner (model.c:331)

    // dataset
    const uint8_t *d_raw = nullptr;
    DevBuf<uint8_t> raw_owned;
    int64_t nv = 0;
    int A = 0, U = 0;

    // fold layout + packed planes
    bool folds_set = false;
    FoldLayout fl{};
    int64_t snp_pad = 0;
    int64_t npos = 0;                     // bit positions per SNP row = nblocks * bw * 32
    std::vector<int32_t> perm;
    std::vector<uint16_t> blk;
    bool tri_suppressed = false;          // the tri layout would fit but 4-word blocks were packed (order 3 / HPGV_NO_TRI)
    std::vector<int32_t> fold_of_sample;  // the assignment of the last set_folds (the order-3 search re-packs without the tri layout)
    DevBuf<FoldLayout> d_fl;
    DevBuf<int32_t> d_perm;
    DevBuf<uint16_t> d_blk;
    DevBuf<uint32_t> d_planes;
    size_t plane_words = 0;

    // search scratch
    DevBuf<Cand> d_lists;
    DevBuf<int> d_list_cnt;
    DevBuf<long long> d_gthr;
    DevBuf<int> d_hist, d_hmax;
    DevBuf<unsigned long long> d_dbg;     // development counters (HPGV_DEBUG_COUNTERS=1)
    DevBuf<int64_t> d_prefix;
    DevBuf<int32_t> d_jt0;
    DevBuf<int2> d_unit_desc;
    bool wl_has_desc = false;
    // search3v2_kernel: per-SNP lists of missing samples (made on the first order-3 search after set_folds)
    DevBuf<uint32_t> d_miss;
    DevBuf<int> d_miss_max;
    DevBuf<uint32_t> d_vmask;             // search3v3_kernel: bit positions of every block that hold a sample
    bool miss_valid = false;
    int miss_cap = 0;
    DevBuf<hpgv_epi_model_t> d_out;
    DevBuf<Cand> d_merge_in;
    // cached work list
    int wl_order = 0, wl_ti = 0;
    uint64_t wl_first = 0, wl_last = 0;
    int64_t wl_nv = -1;
    int wl_it0 = 0, wl_nit = 0;
    int wl_edge_lo = 0, wl_edge_hi = 0;
    int64_t wl_units = 0;
    int wl_band = 0;                      // unit order of the cached list (row-major / band-major, epi_capi.cu build_worklist)
    // device-side timing of the dominant kernel (bench.py's roofline): events around every search launch
    static constexpr int kEvRing = 32;
    cudaEvent_t ev0[kEvRing] = {}, ev1[kEvRing] = {};
    int64_t ev_count = 0;                 // search launches so far; launch k uses slot k % kEvRing
    int last_grid = 0;
    // pinned staging of the fold layout, permutation and block descriptors (set_folds does not synchronise the stream)
    uint8_t *h_stage = nullptr;
    size_t h_stage_cap = 0;
    cudaEvent_t ev_stage = nullptr;
    bool stage_busy = false;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return HPGV_E_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define FAIL(code, msg)       \
    do {                      \
        ctx->err = (msg);     \
        return (code);        \
    } while (0)

// ---------------------------------------------------------------------------------
// reset kernel: global thresholds
// ---------------------------------------------------------------------------------
__global__ void reset_search_kernel(long long *gthr, int *ghmax) {      // ghmax[kMaxFolds] = SearchArgs::gfirst
    if (threadIdx.x < kMaxFolds) { gthr[threadIdx.x] = LLONG_MIN; ghmax[threadIdx.x] = -1; }
    if (threadIdx.x == 0) ghmax[kMaxFolds] = 0;
}

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
extern "C" int hpgv_epi_create(int device, hpgv_epi_ctx **out) {
    if (!out) return HPGV_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = device_count_retry(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        return HPGV_E_CUDA;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= ndev) {
        g_create_error = "device index out of range";
        return HPGV_E_ARG;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return HPGV_E_CUDA;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return HPGV_E_CUDA;
    }
    if (prop.major < 10) {
        g_create_error = "device is not Blackwell (sm_100a kernels only); found sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return HPGV_E_UNSUPPORTED;
    }
    hpgv_epi_ctx *ctx = new hpgv_epi_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem_optin = (int) prop.sharedMemPerBlockOptin;
    for (int k = 0; k < hpgv_epi_ctx::kEvRing; k++) { cudaEventCreate(&ctx->ev0[k]); cudaEventCreate(&ctx->ev1[k]); }
    cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming);
    *out = ctx;
    return HPGV_OK;
}

extern "C" void hpgv_epi_destroy(hpgv_epi_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->raw_owned.release(); ctx->d_fl.release(); ctx->d_perm.release(); ctx->d_blk.release(); ctx->d_planes.release();
    ctx->d_lists.release(); ctx->d_list_cnt.release(); ctx->d_gthr.release(); ctx->d_hist.release(); ctx->d_hmax.release(); ctx->d_dbg.release();
    ctx->d_prefix.release(); ctx->d_jt0.release(); ctx->d_unit_desc.release(); ctx->d_out.release(); ctx->d_merge_in.release(); ctx->d_miss.release(); ctx->d_miss_max.release(); ctx->d_vmask.release();
    for (int k = 0; k < hpgv_epi_ctx::kEvRing; k++) { if (ctx->ev0[k]) cudaEventDestroy(ctx->ev0[k]); if (ctx->ev1[k]) cudaEventDestroy(ctx->ev1[k]); }
    if (ctx->ev_stage) cudaEventDestroy(ctx->ev_stage);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    delete ctx;
}

extern "C" const char *hpgv_epi_last_error(const hpgv_epi_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int hpgv_epi_set_stream(hpgv_epi_ctx *ctx, void *cuda_stream) {
    if (!ctx) return HPGV_E_ARG;
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return HPGV_OK;
}

extern "C" int64_t hpgv_epi_launch_count(const hpgv_epi_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int hpgv_epi_set_eval_function(hpgv_epi_ctx *ctx, int eval_function) {
    if (!ctx) return HPGV_E_ARG;
    if (eval_function == HPGV_EVAL_WBA) FAIL(HPGV_E_UNSUPPORTED, "wBA is declared but not implemented by the reference (model.h:84, no case in model.c:469-478)");
    if (eval_function < HPGV_EVAL_CA || eval_function > HPGV_EVAL_CA_TRUE) FAIL(HPGV_E_ARG, "unknown evaluation function");
    ctx->eval_fn = eval_function;
    return HPGV_OK;
}

// ---------------------------------------------------------------------------------
// dataset
// ---------------------------------------------------------------------------------
static int set_dims(hpgv_epi_ctx *ctx, int64_t nv, int A, int U) {
    if (nv < 2 || nv > (int64_t) INT32_MAX - 4096) FAIL(HPGV_E_ARG, "num_variants must be in [2, 2^31)");
    if (A < 1 || U < 1) FAIL(HPGV_E_ARG, "need at least one affected and one unaffected sample");
    ctx->nv = nv; ctx->A = A; ctx->U = U;
    ctx->folds_set = false;
    // (the cached work list is keyed by nv, order, tile height and range: it survives a reload of the same shape)
    return HPGV_OK;
}

extern "C" int hpgv_epi_load_dataset_host(hpgv_epi_ctx *ctx, const uint8_t *genotypes, int64_t nv, int A, int U) {
    if (!ctx || !genotypes) return HPGV_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = set_dims(ctx, nv, A, U);
    if (rc) return rc;
    const size_t bytes = (size_t) nv * (size_t) (A + U);
    CK(ctx->raw_owned.reserve(bytes));
    CK(cudaMemcpyAsync(ctx->raw_owned.p, genotypes, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->d_raw = ctx->raw_owned.p;
    return HPGV_OK;
}

extern "C" int hpgv_epi_load_dataset_device(hpgv_epi_ctx *ctx, const uint8_t *d_genotypes, int64_t nv, int A, int U) {
    if (!ctx || !d_genotypes) return HPGV_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = set_dims(ctx, nv, A, U);
    if (rc) return rc;
    ctx->d_raw = d_genotypes;
    return HPGV_OK;
}

extern "C" int hpgv_epi_load_dataset_file(hpgv_epi_ctx *ctx, const char *path) {
    if (!ctx || !path) return HPGV_E_ARG;
    FILE *fp = fopen(path, "rb");
    if (!fp) FAIL(HPGV_E_IO, std::string("cannot open dataset file ") + path);
    fseek(fp, 0, SEEK_END);
    const long long len = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    unsigned char hdr[16] = {0};
    const size_t got = fread(hdr, 1, 16, fp);
    if (got < 12) { fclose(fp); FAIL(HPGV_E_IO, "dataset file shorter than its header"); }
    // current format: 3 x uint32 (dataset.c:58-63); legacy fixture: size_t + 2 x uint32
    uint32_t h32[4];
    memcpy(h32, hdr, 16);
    long long nv = 0, A = 0, U = 0, off = 0;
    auto fits = [&](long long n, long long a, long long u, long long o, long long slack) {
        return n >= 1 && a >= 1 && u >= 1 && len >= o + n * (a + u) && len <= o + n * (a + u) + slack;
    };
    if (fits(h32[0], h32[1], h32[2], 12, 0)) { nv = h32[0]; A = h32[1]; U = h32[2]; off = 12; }
    else if (got == 16 && h32[1] == 0 && fits(h32[0], h32[2], h32[3], 16, 7)) { nv = h32[0]; A = h32[2]; U = h32[3]; off = 16; }
    else if (fits(h32[0], h32[1], h32[2], 12, 7)) { nv = h32[0]; A = h32[1]; U = h32[2]; off = 12; }
    else { fclose(fp); FAIL(HPGV_E_IO, "dataset header does not match the file length (neither 12-byte nor legacy 16-byte layout)"); }
    const size_t bytes = (size_t) nv * (size_t) (A + U);
    uint8_t *host = nullptr;
    if (cudaMallocHost(&host, bytes) != cudaSuccess) { fclose(fp); FAIL(HPGV_E_NOMEM, "cannot allocate pinned staging buffer"); }
    fseek(fp, (long) off, SEEK_SET);
    const size_t rd = fread(host, 1, bytes, fp);
    fclose(fp);
    int rc = HPGV_OK;
    if (rd != bytes) { ctx->err = "short read of the genotype matrix"; rc = HPGV_E_IO; }
    if (!rc) rc = hpgv_epi_load_dataset_host(ctx, host, nv, (int) A, (int) U);
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "H2D copy failed"; rc = HPGV_E_CUDA; }
    cudaFreeHost(host);
    return rc;
}

extern "C" int hpgv_epi_dataset_dims(const hpgv_epi_ctx *ctx, int64_t *nv, int *A, int *U) {
    if (!ctx || !ctx->d_raw) return HPGV_E_STATE;
    if (nv) *nv = ctx->nv;
    if (A) *A = ctx->A;
    if (U) *U = ctx->U;
    return HPGV_OK;
}

// ---------------------------------------------------------------------------------
// folds
// ---------------------------------------------------------------------------------
static int apply_folds(hpgv_epi_ctx *ctx, int F, const int32_t *fold_of_sample, bool allow_tri);

extern "C" int hpgv_epi_set_folds(hpgv_epi_ctx *ctx, int F, const int32_t *fold_of_sample) {
    if (!ctx || !fold_of_sample) return HPGV_E_ARG;
    if (!ctx->d_raw) FAIL(HPGV_E_STATE, "set_folds before a dataset was loaded");
    std::vector<int32_t> keep(fold_of_sample, fold_of_sample + (size_t) ctx->A + ctx->U);
    const char *no_tri = getenv("HPGV_NO_TRI");      // A/B switch for benchmarks: keep the 4-word layout
    int rc = apply_folds(ctx, F, keep.data(), !(no_tri && no_tri[0] == '1'));
    if (!rc) ctx->fold_of_sample.swap(keep);
    return rc;
}

// Chooses the sample layout for a fold assignment and packs the planes.  allow_tri = false keeps the 4-word blocks
// where the tri layout would apply (the order-3 kernel has no tri variant).
static int apply_folds(hpgv_epi_ctx *ctx, int F, const int32_t *fold_of_sample, bool allow_tri) {
    if (F < 2 || F > kMaxFolds) FAIL(HPGV_E_ARG, "num_folds must be in [2, 32]");
    CK(cudaSetDevice(ctx->device));
    const int A = ctx->A, U = ctx->U, S = A + U;

    FoldLayout fl{};
    fl.F = F; fl.nseg = 2 * F; fl.A = A; fl.U = U;
    fl.balanced = (A == U);
    fl.ratio = (float) A / (float) U;                 // mdr.c:52
    std::vector<int> seg_size(2 * F, 0);
    for (int s = 0; s < S; s++) {
        const int f = fold_of_sample[s];
        if (f < 0 || f >= F) FAIL(HPGV_E_ARG, "fold_of_sample holds a fold id outside [0, num_folds)");
        seg_size[2 * f + (s < A ? 0 : 1)]++;
    }
    int max_seg = 0;
    fl.eqfolds = 1;
    for (int f = 0; f < F; f++) {
        fl.a_in[f] = seg_size[2 * f]; fl.u_in[f] = seg_size[2 * f + 1];
        if (fl.a_in[f] != fl.u_in[f]) fl.eqfolds = 0;
        max_seg = std::max(max_seg, std::max(fl.a_in[f], fl.u_in[f]));
    }
    if (max_seg > 65535) FAIL(HPGV_E_UNSUPPORTED, "more than 65535 samples of one class in one fold");
    if ((long long) A * U >= (1LL << 62)) FAIL(HPGV_E_UNSUPPORTED, "cohort too large");
    // Layout: a segment is one block when it fits 4 or 8 words (byte counters), else a run of 8-word blocks.
    fl.single = max_seg <= 255;
    fl.bw = max_seg <= 128 ? 4 : 8;
    const int bits_per_block = 32 * fl.bw;
    // 8-word blocks: when 224 sample bits per block need no more blocks than 256 do, the last word of every block stays
    // empty and the search kernels compress 7 words into 3 POPCs instead of 8 into 4
    fl.w7 = 0;
    if (fl.bw == 8) {
        int nb7 = 0, nb8 = 0;
        for (int s = 0; s < 2 * F; s++) { nb7 += std::max(1, (seg_size[s] + 223) / 224); nb8 += std::max(1, (seg_size[s] + 255) / 256); }
        fl.w7 = (nb7 == nb8) ? 1 : 0;
    }
    const int used_bits = fl.w7 ? 224 : bits_per_block;
    std::vector<int> seg_blocks(2 * F), seg_first(2 * F);
    int nb = 0;
    for (int s = 0; s < 2 * F; s++) {
        seg_blocks[s] = fl.single ? 1 : std::max(1, (seg_size[s] + used_bits - 1) / used_bits);
        seg_first[s] = nb;
        nb += seg_blocks[s];
    }
    const int nb_real = nb;
    if (fl.single) nb = (nb + 3) / 4 * 4;              // byte counters are packed four blocks to a word
    if (nb > kMaxBlocks) FAIL(HPGV_E_UNSUPPORTED, "sample axis needs more than 4096 blocks (1M samples)");
    fl.nblocks = nb;
    const bool tri_fits = fl.single && max_seg <= 100 && nb <= 24;
    fl.tri = (allow_tri && tri_fits) ? 1 : 0;
    ctx->tri_suppressed = tri_fits && !allow_tri;
    ctx->npos = (int64_t) nb * bits_per_block;
    fl.marg = 0; fl.marg_off = 0; fl.marg_stride = 4; fl.mlist = 0;
    if (fl.tri) {
        fl.bw = 4;                                       // logical positions: word 3 of a block = its 4-bit tail
        fl.cb = nb; fl.nchunks = 1;
        fl.row_words = tri_row_words(nb);
        fl.marg = 1; fl.marg_off = tri_marg_off(nb, 0);
    } else {
    // chunks: a stage of the search kernels holds (16 + 32 + 1) chunk rows; keep it within 48 KB
        const int block_bytes = 3 * fl.bw * 4;
        int cb_max = std::max(1, (48 * 1024 / (kMaxWarps + kTileJ + 1) - 16) / block_bytes);
        if (fl.single) cb_max = std::max(4, cb_max / 4 * 4);
        fl.nchunks = (nb + cb_max - 1) / cb_max;
        fl.cb = (nb + fl.nchunks - 1) / fl.nchunks;
        if (fl.single) fl.cb = (fl.cb + 3) / 4 * 4;
        fl.nchunks = (nb + fl.cb - 1) / fl.cb;
        // single-block segments: the per-group marginals follow the planes of a chunk row when the staged packer (which
        // writes them) can run and a stage of 49 rows still fits 48 KB
        const char *old_packer = getenv("HPGV_PACK_WARP");
        const char *no_list = getenv("HPGV_MISS_LIST");          // A/B switch: "0" = marginals without the lists of missing samples
        // level 2: quad + list of the group's missing samples, level 1: quad only, level 0: no marginals
        for (int level = (fl.single && !(old_packer && old_packer[0] == '1')) ? ((no_list && no_list[0] == '0') ? 1 : 2) : 0; level >= 0; level--) {
            fl.marg = level > 0;
            fl.mlist = level == 2;
            fl.marg_stride = level == 2 ? 4 + kMissListWords : 4;
            fl.marg_off = fl.cb * 3 * fl.bw;
            fl.row_words = fl.cb * 3 * fl.bw + (level ? (fl.cb / 4) * fl.marg_stride : 0);
            if ((fl.row_words / 4) % 2 == 0) fl.row_words += 4;   // odd number of 16-byte groups per row: conflict-free LDS.128
            if (!level) break;
            if (pack_smem_map(ctx->npos, S, fl).total <= 96 * 1024 && (size_t) (kMaxWarps + kTileJ + 1) * fl.row_words * 4 <= 48 * 1024) break;
        }
    }
    ctx->blk.assign(nb, (uint16_t) 0x7fff);            // padding blocks belong to no segment
    for (int s = 0; s < 2 * F; s++)
        for (int b = 0; b < seg_blocks[s]; b++)
            ctx->blk[seg_first[s] + b] = (uint16_t) (s | (b == seg_blocks[s] - 1 ? 0x8000 : 0));
    (void) nb_real;
    ctx->perm.assign((size_t) ctx->npos, -1);
    {
        std::vector<int> fill(2 * F, 0);
        for (int s = 0; s < S; s++) {       // ascending dataset column inside each segment
            const int seg = 2 * fold_of_sample[s] + (s < A ? 0 : 1);
            const int p = fill[seg]++;
            ctx->perm[(size_t) (seg_first[seg] + p / used_bits) * bits_per_block + p % used_bits] = s;
        }
    }
    ctx->fl = fl;
    ctx->miss_valid = false;
    // rows are padded so that every tile a kernel stages (<= 32 rows past any valid origin) stays inside the buffer
    ctx->snp_pad = ((ctx->nv + kTileJ - 1) / kTileJ) * kTileJ + kTileJ;
    ctx->plane_words = (size_t) fl.nchunks * (size_t) ctx->snp_pad * fl.row_words;

    // a growing device buffer is freed and re-allocated: nothing enqueued earlier may still use it
    if (ctx->perm.size() > ctx->d_perm.cap || ctx->blk.size() > ctx->d_blk.cap || ctx->plane_words > ctx->d_planes.cap || ctx->d_fl.cap < 1)
        CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->d_fl.reserve(1));
    CK(ctx->d_perm.reserve(ctx->perm.size()));
    CK(ctx->d_blk.reserve(ctx->blk.size()));
    CK(ctx->d_planes.reserve(ctx->plane_words));
    {
        // layout, permutation and block descriptors go up from pinned staging memory: truly asynchronous copies, the
        // staging buffer is reused once the copies of the previous call have completed
        const size_t o_perm = align_up(sizeof(FoldLayout), 16), o_blk = o_perm + align_up(ctx->perm.size() * sizeof(int32_t), 16);
        const size_t need = o_blk + align_up(ctx->blk.size() * sizeof(uint16_t), 16);
        if (ctx->stage_busy) CK(cudaEventSynchronize(ctx->ev_stage));
        if (need > ctx->h_stage_cap) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
            CK(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_stage), need));
            ctx->h_stage_cap = need;
        }
        memcpy(ctx->h_stage, &ctx->fl, sizeof(FoldLayout));
        memcpy(ctx->h_stage + o_perm, ctx->perm.data(), ctx->perm.size() * sizeof(int32_t));
        memcpy(ctx->h_stage + o_blk, ctx->blk.data(), ctx->blk.size() * sizeof(uint16_t));
        CK(cudaMemcpyAsync(ctx->d_fl.p, ctx->h_stage, sizeof(FoldLayout), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_perm.p, ctx->h_stage + o_perm, ctx->perm.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_blk.p, ctx->h_stage + o_blk, ctx->blk.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->ev_stage, ctx->stream));
        ctx->stage_busy = true;
    }
    CK(cudaMemsetAsync(ctx->d_planes.p, 0, ctx->plane_words * sizeof(uint32_t), ctx->stream));

    const int threads = 256;
    const PackSmem pm = pack_smem_map(ctx->npos, S, fl);
    const char *old_packer = getenv("HPGV_PACK_WARP");          // A/B switch: the one-warp-per-word packer
    // (the tri layout's marginals are only written by the staged packer; its permutation always fits)
    if (pm.total <= 96 * 1024 && (fl.tri || !(old_packer && old_packer[0] == '1'))) {
        // several CTAs per SM hide the latency of the row loads; each walks SNPs blockIdx.x, + gridDim.x, ...
        if (pm.total > 48 * 1024) {
            CK(cudaFuncSetAttribute(pack_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pm.total));
            CK(cudaFuncSetAttribute(pack_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pm.total));
        }
        const int per_sm = (int) std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / pm.total));
        const int64_t batches = (ctx->nv + kPackRows - 1) / kPackRows;
        int64_t grid = std::min<int64_t>(batches, (int64_t) ctx->num_sms * per_sm);
        grid = (batches + (batches + grid - 1) / grid - 1) / ((batches + grid - 1) / grid);     // every CTA the same number of batches
        const char *gp = getenv("HPGV_PACK_GENERIC");            // A/B switch: "1" packs the tri layout with the generic path
        if (fl.tri && !(gp && gp[0] == '1'))
            pack_rows_kernel<true><<<(unsigned) grid, threads, pm.total, ctx->stream>>>(ctx->d_raw, ctx->nv, S, ctx->d_perm.p, ctx->d_fl.p,
                                                                                        ctx->snp_pad, ctx->npos, ctx->d_planes.p);
        else
            pack_rows_kernel<false><<<(unsigned) grid, threads, pm.total, ctx->stream>>>(ctx->d_raw, ctx->nv, S, ctx->d_perm.p, ctx->d_fl.p,
                                                                                         ctx->snp_pad, ctx->npos, ctx->d_planes.p);
    } else {
        const int64_t warps = ctx->nv * (int64_t) nb * fl.bw;
        const int64_t blocks = (warps * 32 + threads - 1) / threads;
        if (blocks > INT32_MAX) FAIL(HPGV_E_UNSUPPORTED, "dataset too large for the packer grid");
        pack_planes_kernel<<<(unsigned) blocks, threads, 0, ctx->stream>>>(ctx->d_raw, ctx->nv, S, ctx->d_perm.p, ctx->d_fl.p, ctx->snp_pad, ctx->d_planes.p);
    }
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->folds_set = true;
    return HPGV_OK;
}

// drand48-compatible generator (48-bit LCG of POSIX): the reference shuffles with
// srand48(seed) + drand48() (lib/c/src/math/data/array_utils.c:173-188)
namespace {
struct Rand48 {
    uint64_t x;
    explicit Rand48(long seed) : x((((uint64_t) (uint32_t) seed) << 16) | 0x330EULL) {}
    double next() {
        x = (x * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (double) x / 281474976710656.0;
    }
};
void shuffle48(int *v, size_t n, long seed) {
    if (n <= 1) return;
    Rand48 rng(seed);
    for (size_t i = n - 1; i > 0; i--) {
        size_t j = (unsigned int) (rng.next() * (double) (i + 1));
        std::swap(v[i], v[j]);
    }
}
}  // namespace

extern "C" int hpgv_epi_k_folds(int A, int U, int k, long seed, int32_t *fold_of_sample, uint32_t *sizes) {
    if (A < 0 || U < 0 || k < 1 || !fold_of_sample) return HPGV_E_ARG;
    std::vector<int> samples((size_t) A + U);
    for (int i = 0; i < A + U; i++) samples[i] = i;
    shuffle48(samples.data(), (size_t) A, seed);          // cases and controls separately (cross_validation.c:20-21)
    shuffle48(samples.data() + A, (size_t) U, seed);
    std::vector<uint32_t> fs(3 * (size_t) k, 0);
    // round-robin deal, one case and one control per fold per pass (cross_validation.c:45-67)
    for (int i = 0; i < A; i++) { fold_of_sample[samples[i]] = i % k; fs[3 * (i % k) + 1]++; }
    for (int i = 0; i < U; i++) { fold_of_sample[samples[A + i]] = i % k; fs[3 * (i % k) + 2]++; }
    for (int f = 0; f < k; f++) fs[3 * f] = fs[3 * f + 1] + fs[3 * f + 2];
    if (sizes) memcpy(sizes, fs.data(), fs.size() * sizeof(uint32_t));
    return HPGV_OK;
}

extern "C" uint64_t hpgv_epi_num_combinations(int64_t nv, int order) {
    if (nv < order) return 0;
    if (order == 2) return choose2((uint64_t) nv);
    if (order == 3) return choose3((uint64_t) nv);
    return 0;
}

// ---------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------
namespace {

// (i, j) of linear pair index idx
void unrank_pair(uint64_t n, uint64_t idx, int64_t *i_out, int64_t *j_out) {
    // offset(i) = i(2n-i-1)/2 ; find the largest i with offset(i) <= idx
    double nn = (double) n;
    double disc = (2 * nn - 1) * (2 * nn - 1) - 8.0 * (double) idx;
    int64_t i = (int64_t) (((2 * nn - 1) - std::sqrt(std::max(0.0, disc))) / 2.0);
    i = std::max<int64_t>(0, std::min<int64_t>(i, (int64_t) n - 2));
    while (i > 0 && pair_index(n, (uint64_t) i, (uint64_t) i + 1) > idx) i--;
    while (i + 1 <= (int64_t) n - 2 && pair_index(n, (uint64_t) i + 1, (uint64_t) i + 2) <= idx) i++;
    *i_out = i;
    *j_out = (int64_t) (idx - pair_index(n, (uint64_t) i, (uint64_t) i + 1)) + i + 1;
}

// first element of the triple with linear index idx
int64_t unrank_triple_first(uint64_t n, uint64_t idx) {
    int64_t lo = 0, hi = (int64_t) n - 3;     // largest i with (choose3(n) - choose3(n - i)) <= idx
    const uint64_t c3n = choose3(n);
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) / 2;
        if (c3n - choose3(n - (uint64_t) mid) <= idx) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename K>
cudaError_t opt_in_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

}  // namespace

static int build_worklist(hpgv_epi_ctx *ctx, int order, int ti, uint64_t first, uint64_t last) {
    const char *bt_key = getenv("HPGV_BAND_TILES");
    const int band_key = ((ctx->plane_words * sizeof(uint32_t)) > ((size_t) 64 << 20) ? 1 : 0) + (bt_key ? 2 * atoi(bt_key) : 0);
    if (ctx->wl_nv == ctx->nv && ctx->wl_order == order && ctx->wl_ti == ti && ctx->wl_first == first && ctx->wl_last == last && ctx->wl_band == band_key)
        return HPGV_OK;
    ctx->wl_band = band_key;
    const int64_t nv = ctx->nv;
    std::vector<int64_t> prefix;
    std::vector<int32_t> jt0;
    int it0 = 0;
    int64_t units = 0;
    int64_t edge_lo = 0, edge_hi = 0;
    if (first < last) {
        if (order == 2) {
            int64_t i_first, j_first, i_last, j_last;
            unrank_pair((uint64_t) nv, first, &i_first, &j_first);
            unrank_pair((uint64_t) nv, last - 1, &i_last, &j_last);
            edge_lo = i_first; edge_hi = i_last;
            it0 = (int) (i_first / ti);
            const int it1 = (int) (i_last / ti);
            for (int t = it0; t <= it1; t++) {
                int64_t jlo = INT64_MAX, jhi = -1;
                for (int64_t r = std::max<int64_t>((int64_t) t * ti, i_first); r < (int64_t) (t + 1) * ti && r <= i_last && r <= nv - 2; r++) {
                    const int64_t lo = (r == i_first) ? j_first : r + 1;
                    const int64_t hi = (r == i_last) ? j_last : nv - 1;
                    if (lo <= hi) { jlo = std::min(jlo, lo); jhi = std::max(jhi, hi); }
                }
                prefix.push_back(units);
                if (jhi >= 0) { jt0.push_back((int32_t) (jlo / kTileJ)); units += jhi / kTileJ - jlo / kTileJ + 1; }
                else jt0.push_back(0);
            }
        } else {
            // unit = (row i, tile of `ti` j rows, one per warp); the CTA walks the k tiles itself
            const int tj = ti;
            const int64_t i_first = unrank_triple_first((uint64_t) nv, first);
            const int64_t i_last = unrank_triple_first((uint64_t) nv, last - 1);
            edge_lo = i_first; edge_hi = i_last;
            it0 = (int) i_first;
            for (int64_t i = i_first; i <= i_last; i++) {
                prefix.push_back(units);
                const int64_t jlo = i + 1, jhi = nv - 2;
                if (jlo <= jhi) { jt0.push_back((int32_t) (jlo / tj)); units += jhi / tj - jlo / tj + 1; }
                else jt0.push_back(0);
            }
        }
    }
    if (prefix.empty()) { prefix.push_back(0); jt0.push_back(0); }
    // order 2: the tile origins of every unit (8 bytes per unit of 512..640 pairs), while the list stays below 128 MB.
    // When the packed planes do not fit the L2 (c3: 235 MB, c5: 391 MB against 126 MB) the units are listed band by band:
    // a band is a range of j tiles whose rows -- all chunks of them -- take about a quarter of the L2, and every i tile
    // works through the band before the next band is touched.  In row-major order every i tile streams ALL later rows
    // from HBM again (c3: 366 GB per search measured, 1500 x the 245 MB of planes); band-major, a band is read once and
    // the i rows once per band.
    std::vector<int2> desc;
    ctx->wl_has_desc = false;
    if (order == 2 && units > 0 && units <= (int64_t) 16 << 20) {
        desc.reserve((size_t) units);
        const size_t row_bytes_all = (size_t) ctx->fl.nchunks * ctx->fl.row_words * 4;
        const size_t plane_bytes = ctx->plane_words * sizeof(uint32_t);
        const char *bo = getenv("HPGV_BAND_ORDER");              // A/B switch: "0" keeps the row-major order
        int64_t band_tiles = INT64_MAX;
        if (plane_bytes > (size_t) 64 << 20 && !(bo && bo[0] == '0'))
            band_tiles = std::max<int64_t>(1, (int64_t) (((size_t) 32 << 20) / (row_bytes_all * kTileJ)));
        const char *bt = getenv("HPGV_BAND_TILES");              // tests: force a band width (in j tiles) whatever the plane size
        if (bt && atoi(bt) > 0) band_tiles = atoi(bt);
        const int64_t njt = (nv + kTileJ - 1) / kTileJ;
        for (int64_t b0 = 0; b0 < njt; b0 += std::min<int64_t>(band_tiles, njt)) {
            const int64_t b1 = band_tiles == INT64_MAX ? njt : std::min(njt, b0 + band_tiles);
            for (size_t t = 0; t < prefix.size(); t++) {
                const int64_t n = (t + 1 < prefix.size() ? prefix[t + 1] : units) - prefix[t];
                const int64_t lo = std::max<int64_t>(jt0[t], b0), hi = std::min<int64_t>(jt0[t] + n, b1);
                for (int64_t jt = lo; jt < hi; jt++) desc.push_back(make_int2((it0 + (int) t) * ti, (int) (jt * kTileJ)));
            }
            if (band_tiles == INT64_MAX) break;
        }
        CK(ctx->d_unit_desc.reserve(desc.size()));
        CK(cudaMemcpyAsync(ctx->d_unit_desc.p, desc.data(), desc.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
        ctx->wl_has_desc = true;
    }
    CK(ctx->d_prefix.reserve(prefix.size()));
    CK(ctx->d_jt0.reserve(jt0.size()));
    CK(cudaMemcpyAsync(ctx->d_prefix.p, prefix.data(), prefix.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_jt0.p, jt0.data(), jt0.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
    ctx->wl_nv = nv; ctx->wl_order = order; ctx->wl_ti = ti; ctx->wl_first = first; ctx->wl_last = last;
    ctx->wl_it0 = it0; ctx->wl_nit = (int) prefix.size(); ctx->wl_units = units;
    ctx->wl_edge_lo = (int) edge_lo; ctx->wl_edge_hi = (int) edge_hi;
    return HPGV_OK;
}

// Picks the CTA size (16, 8, 4, 2 or 1 warps = tile rows) that fits shared memory, and whether the per-CTA top-N
// lists can live in shared memory during the search.
struct SearchShape {
    int nthreads = 0;
    bool lists_in_smem = false;
    int nstages = 2;
    size_t smem = 0;
};
static SearchShape pick_shape(const hpgv_epi_ctx *ctx, int order, int rank) {
    const FoldLayout &fl = ctx->fl;
    const int ncells = order == 2 ? 9 : 27;
    SearchShape best;
    const char *tw_env = getenv("HPGV_TRI_WARPS");               // A/B switch: "16" keeps the tri kernel at 16 warps
    const bool tri20 = order == 2 && fl.single && !(tw_env && tw_env[0] == '1' && tw_env[1] == '6');
    // 20 (byte-counter order-2 kernels), 16, then for order 2 also 12 and 10 (c5: 90 counter words per thread leave room for ten
    // warps, not for sixteen), 8, 4, 2, 1
    auto next_warps = [&](int w) { return w == kTriWarps ? kMaxWarps : ((order == 2 && w == 16) ? 12 : ((order == 2 && w == 12) ? 10 : (w == 10 ? 8 : w >> 1))); };
    for (int warps = tri20 ? kTriWarps : kMaxWarps; warps >= 1; warps = next_warps(warps)) {
        const int rows = order == 2 ? warps + kTileJ : 1 + warps + kTileJ;
        for (int in_smem = 1; in_smem >= 0; in_smem--) {
            if (in_smem && (size_t) fl.F * rank * sizeof(Cand) > 24 * 1024) continue;
            // a third stage lets warps drift a step further apart before they wait for each other (order-2 kernel)
            const char *ns_env = getenv("HPGV_STAGES");
            const int ns_max = (order == 2 && !(ns_env && ns_env[0] == '2')) ? 3 : 2;
            for (int ns = ns_max; ns >= 2; ns--) {
                const SmemMap m = search_smem_map(fl, rows, ncells, warps * 32, rank, in_smem != 0, ns);
                if (m.total <= (size_t) ctx->max_smem_optin) {
                    best.nthreads = warps * 32; best.lists_in_smem = in_smem != 0; best.nstages = ns; best.smem = m.total;
                    return best;
                }
            }
        }
    }
    return best;
}

static int launch_search(hpgv_epi_ctx *ctx, search_kernel_t kernel, const SearchShape &shape, SearchArgs &args, int F, int rank) {
    CK(opt_in_smem(kernel, shape.smem));
    int64_t grid = ctx->num_sms;                       // one persistent CTA per SM
    grid = std::max<int64_t>(1, std::min<int64_t>(grid, std::max<int64_t>(args.num_units, 1)));
    CK(ctx->d_lists.reserve((size_t) grid * F * rank));
    CK(ctx->d_list_cnt.reserve((size_t) grid * F));
    CK(ctx->d_gthr.reserve(kMaxFolds));
    CK(ctx->d_hmax.reserve(kMaxFolds + 1));
    args.dbg = nullptr;
    {
        const char *dc = getenv("HPGV_DEBUG_COUNTERS");
        if (dc && dc[0] == '1') {
            CK(ctx->d_dbg.reserve(kDbgWords));
            CK(cudaMemsetAsync(ctx->d_dbg.p, 0, (kDbgTrace + 8) * sizeof(unsigned long long), ctx->stream));
            args.dbg = ctx->d_dbg.p;
        }
    }
    {
        const char *fb = getenv("HPGV_FRESH_BOUND");          // A/B switch: "1" = adopt other CTAs' published bounds once per unit
        args.fresh_bound = (fb && fb[0] == '1') ? 1 : 0;       // measured: costs 2 % and buys nothing once the start-up race is gone
        const char *fw = getenv("HPGV_FIRST_WAIT");
        args.first_wait = (fw && fw[0] == '0') ? 0 : 1;
    }
    args.lists = ctx->d_lists.p;
    args.list_cnt = ctx->d_list_cnt.p;
    args.gthr = ctx->d_gthr.p;
    args.ghmax = ctx->d_hmax.p;
    args.gfirst = ctx->d_hmax.p + kMaxFolds;
    args.ghist = nullptr;
    if (args.use_hist) {
        const size_t bins = (size_t) F * (args.hist_bins + hist_coarse_bins(args.hist_bins));
        CK(ctx->d_hist.reserve(bins));
        CK(cudaMemsetAsync(ctx->d_hist.p, 0, bins * sizeof(int), ctx->stream));
        args.ghist = ctx->d_hist.p;
    }
    args.lists_in_smem = shape.lists_in_smem ? 1 : 0;
    args.nstages = shape.nstages;
    reset_search_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_gthr.p, ctx->d_hmax.p);
    CK(cudaGetLastError());
    const int slot = (int) (ctx->ev_count % hpgv_epi_ctx::kEvRing);
    CK(cudaEventRecord(ctx->ev0[slot], ctx->stream));
    kernel<<<(unsigned) grid, shape.nthreads, shape.smem, ctx->stream>>>(args);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1[slot], ctx->stream));
    ctx->ev_count++;
    ctx->last_grid = (int) grid;
    ctx->launches += 2;
    return (int) grid;
}

static int launch_merge(hpgv_epi_ctx *ctx, const MergeArgs &m) {
    const size_t smem = merge_smem_bytes(m.rank_out);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    merge_kernel<<<m.F, kMergeThreads, smem, ctx->stream>>>(m);
    CK(cudaGetLastError());
    ctx->launches++;
    return HPGV_OK;
}

// ---- order 3, resident (j, k) tiles (search3v2_kernel) ---------------------------------------------------------------
struct Shape3 {
    int tj = 0, ni = 0;
    bool lists_in_smem = false;
    size_t smem = 0;
};
static Shape3 pick_shape3(const hpgv_epi_ctx *ctx, int rank, int mcap) {
    const FoldLayout &fl = ctx->fl;
    Shape3 best;
    if (fl.tri || fl.nchunks > 15 || fl.cb * 3 * fl.bw > 2047 || fl.nblocks / 4 > 1023) return best;
    for (int tj = 8; tj >= 4; tj -= 2)
        for (int in_smem = 1; in_smem >= 0; in_smem--) {
            if (in_smem && (size_t) fl.F * rank * sizeof(Cand) > 24 * 1024) continue;
            for (int ni = 4; ni >= 1; ni >>= 1) {
                const Smem3Map m = search3v2_smem_map(fl, tj, ni, 2, mcap, rank, in_smem != 0);
                if (m.total <= (size_t) ctx->max_smem_optin) {
                    best.tj = tj; best.ni = ni; best.lists_in_smem = in_smem != 0; best.smem = m.total;
                    return best;
                }
            }
        }
    return best;
}

// lists of missing samples per SNP; returns the list capacity (entries per SNP incl. the end marker), 0 when some SNP misses too much
static int build_miss_lists(hpgv_epi_ctx *ctx) {
    if (ctx->miss_valid) return ctx->miss_cap;
    const int64_t rows = ctx->snp_pad;
    const int threads = 256;
    const unsigned grid = (unsigned) ((rows * 32 + threads - 1) / threads);
    CK(ctx->d_miss_max.reserve(1));
    CK(cudaMemsetAsync(ctx->d_miss_max.p, 0, sizeof(int), ctx->stream));
    miss_list_kernel<<<grid, threads, 0, ctx->stream>>>(ctx->d_raw, ctx->nv, rows, ctx->A + ctx->U, ctx->d_perm.p, ctx->d_fl.p, ctx->d_blk.p, ctx->npos,
                                                        0, nullptr, ctx->d_miss_max.p);
    CK(cudaGetLastError());
    int mx = 0;
    CK(cudaMemcpyAsync(&mx, ctx->d_miss_max.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->launches++;
    ctx->miss_valid = true;
    ctx->miss_cap = 0;
    if (mx > 250) return 0;                           // that many missing samples per SNP: the plain kernel counts every cell instead
    const int mcap = (mx + 1 + 3) / 4 * 4;            // + end marker, 16-byte multiples (bulk copies)
    CK(ctx->d_miss.reserve((size_t) rows * mcap));
    miss_list_kernel<<<grid, threads, 0, ctx->stream>>>(ctx->d_raw, ctx->nv, rows, ctx->A + ctx->U, ctx->d_perm.p, ctx->d_fl.p, ctx->d_blk.p, ctx->npos,
                                                        mcap, ctx->d_miss.p, ctx->d_miss_max.p);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->miss_cap = mcap;
    return mcap;
}

// search3v3_kernel (balanced cohorts, pre-filter applicable): eight warps, the pair table in registers
static Shape3 pick_shape3b(const hpgv_epi_ctx *ctx, int rank, int mcap) {
    const FoldLayout &fl = ctx->fl;
    Shape3 best;
    const int nwc = fl.single ? fl.nblocks / 4 : fl.F;
    if (fl.tri || fl.nchunks > 15 || fl.cb * 3 * fl.bw > 2047 || fl.nblocks / 4 > 1023 || nwc > 5) return best;   // (built for <= 5 counter words per cell)
    for (int in_smem = 1; in_smem >= 0; in_smem--) {
        if (in_smem && (size_t) fl.F * rank * sizeof(Cand) > 24 * 1024) continue;
        for (int ni = 4; ni >= 1; ni >>= 1) {
            const Smem3bMap m = search3v3_smem_map(fl, 8, ni, mcap, rank, in_smem != 0);
            if (m.total <= (size_t) ctx->max_smem_optin) {
                best.tj = 8; best.ni = ni; best.lists_in_smem = in_smem != 0; best.smem = m.total;
                return best;
            }
        }
    }
    return best;
}

// valid-sample masks per block word (which bit positions hold a sample), for the 1' plane of search3v3_kernel
static int build_vmask(hpgv_epi_ctx *ctx) {
    const int bw = ctx->fl.bw, nwords = ctx->fl.nblocks * bw;
    std::vector<uint32_t> vm((size_t) nwords, 0u);
    for (int64_t pos = 0; pos < ctx->npos; pos++)
        if (ctx->perm[(size_t) pos] >= 0) vm[(size_t) (pos >> 5)] |= 1u << (pos & 31);
    CK(ctx->d_vmask.reserve(vm.size()));
    CK(cudaMemcpyAsync(ctx->d_vmask.p, vm.data(), vm.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HPGV_OK;
}

// units of search3v2_kernel: (j tile, k tile) pairs that hold some j < k, longest i loops first
static int build_worklist3(hpgv_epi_ctx *ctx, int tj, uint64_t first, uint64_t last) {
    if (ctx->wl_nv == ctx->nv && ctx->wl_order == 3 && ctx->wl_ti == -tj && ctx->wl_first == first && ctx->wl_last == last) return HPGV_OK;
    const int64_t nv = ctx->nv;
    std::vector<int2> desc;
    int64_t i_first = 0, i_last = -1;
    if (first < last) {
        i_first = unrank_triple_first((uint64_t) nv, first);
        i_last = unrank_triple_first((uint64_t) nv, last - 1);
        const int64_t njt = (nv - 1 + tj - 1) / tj, nkt = (nv + kTileJ - 1) / kTileJ;
        for (int64_t jt = njt - 1; jt >= 0; jt--) {
            const int64_t j0 = jt * tj;
            if (j0 + tj - 1 <= i_first) break;        // no j of this tile (or of any before it) lies past the range's first i
            for (int64_t kt = (j0 + 1) / kTileJ; kt < nkt; kt++) desc.push_back(make_int2((int) j0, (int) (kt * kTileJ)));
            if ((int64_t) desc.size() > ((int64_t) 4 << 20)) FAIL(HPGV_E_UNSUPPORTED, "order-3 unit list too long");
        }
    }
    if (!desc.empty()) {
        CK(ctx->d_unit_desc.reserve(desc.size()));
        CK(cudaMemcpyAsync(ctx->d_unit_desc.p, desc.data(), desc.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->wl_nv = nv; ctx->wl_order = 3; ctx->wl_ti = -tj; ctx->wl_first = first; ctx->wl_last = last;
    ctx->wl_units = (int64_t) desc.size();
    ctx->wl_edge_lo = (int) i_first; ctx->wl_edge_hi = (int) i_last;
    ctx->wl_has_desc = true;
    return HPGV_OK;
}

static int do_search(hpgv_epi_ctx *ctx, int order, int eval_subset, int rank, uint64_t first, uint64_t last,
                     hpgv_epi_model_t *d_out) {
    if (!ctx->folds_set) FAIL(HPGV_E_STATE, "search before set_folds");
    if (order != 2 && order != 3) FAIL(HPGV_E_ARG, "order must be 2 or 3");
    if (rank < 1 || rank > kMaxRank) FAIL(HPGV_E_ARG, "rank_size must be in [1, 4096]");
    if (eval_subset != HPGV_SUBSET_TESTING && eval_subset != HPGV_SUBSET_TRAINING) FAIL(HPGV_E_ARG, "eval_subset must be 0 (testing) or 1 (training)");
    if (ctx->nv < order) FAIL(HPGV_E_ARG, "fewer variants than the order");
    if (order == 3 && ctx->nv >= (1 << 21)) FAIL(HPGV_E_UNSUPPORTED, "order 3 supports fewer than 2^21 variants");
    CK(cudaSetDevice(ctx->device));
    const uint64_t total = hpgv_epi_num_combinations(ctx->nv, order);
    if (last > total) last = total;
    if (first > last) first = last;

    // the tri layout exists for the order-2 kernel only: re-pack with 4-word blocks for order 3 (and back for order 2)
    if (!ctx->fold_of_sample.empty()) {
        const char *no_tri = getenv("HPGV_NO_TRI");
        const bool want_tri = order == 2 && !(no_tri && no_tri[0] == '1');
        if ((ctx->fl.tri != 0) != want_tri && (ctx->fl.tri || ctx->tri_suppressed)) {
            const std::vector<int32_t> fos = ctx->fold_of_sample;
            int rc = apply_folds(ctx, ctx->fl.F, fos.data(), want_tri);
            if (rc) return rc;
        }
    }
    const FoldLayout &fl = ctx->fl;
    const int F = fl.F;
    SearchArgs args{};
    args.planes = ctx->d_planes.p;
    args.blk_desc = ctx->d_blk.p;
    args.fl = ctx->d_fl.p;
    args.snp_pad = ctx->snp_pad;
    args.nv = (int) ctx->nv;
    args.training = (eval_subset == HPGV_SUBSET_TRAINING);
    args.rank = rank;
    args.first = first;
    args.last = last;
    {
        const char *st = getenv("HPGV_STAGGER");
        args.stagger = st && st[0] >= '0' && st[0] <= '2' ? st[0] - '0' : 1;
        const char *ls = getenv("HPGV_LIST_SCAN");
        args.list_scan = (ls && ls[0] == '0') ? 0 : 1;        // default on; "0" keeps the heap for every list length
        const char *td = getenv("HPGV_TRI_DERIVE");
        args.tri_derive = !(td && td[0] == '0');
    }

    // order 3: the kernel with resident (j, k) tiles when its tables fit shared memory and no SNP misses hundreds of samples
    bool use_v2 = false, use_v3 = false;
    SearchShape shape;
    const bool balanced = fl.balanced && fl.A <= 65535;
    args.eval_fn = ctx->eval_fn == kEvalCA ? kEvalBA : ctx->eval_fn;      // the reference turns code 0 into BA (model.c:465-467)
    // the balanced pre-filter (epilogue_balanced_t) orders by TP - FP: BA on the TRAINING part of equal folds only
    args.prefilter = (balanced && fl.eqfolds && args.training && args.eval_fn == kEvalBA) ? 1 : 0;
    if (order == 3) {
        const char *v2 = getenv("HPGV_SEARCH3_V2");           // A/B switches: "0" keeps the plain order-3 kernel,
        const char *v3 = getenv("HPGV_SEARCH3_V3");           // "1" runs the one-thread-per-triple kernel where it applies (measured slower)
        if (!(v2 && v2[0] == '0')) {
            const int mcap = build_miss_lists(ctx);
            if (mcap < 0) return mcap;
            // one thread per triple, pair table in registers: balanced cohorts whose pre-filter applies (its two-pass epilogue
            // only pays when nearly every fold stops at the pre-filter).  Opt-in: 18 % fewer instructions than the
            // two-threads-per-triple kernel but eight warps per SM instead of 12-16, and slower for it (c4: 12.3 s against 11.4 s)
            const Shape3 s3b = (mcap > 0 && args.prefilter && v3 && v3[0] == '1') ? pick_shape3b(ctx, rank, mcap) : Shape3();
            if (s3b.tj > 0) {
                int rc3 = build_worklist3(ctx, s3b.tj, first, last);
                if (rc3 == HPGV_OK) rc3 = build_vmask(ctx);
                if (rc3 == HPGV_OK) {
                    use_v3 = true;
                    shape.nthreads = s3b.tj * 32; shape.lists_in_smem = s3b.lists_in_smem; shape.nstages = 2; shape.smem = s3b.smem;
                    args.v2_ni = s3b.ni; args.v2_mcap = mcap; args.v2_miss = ctx->d_miss.p; args.v3_vmask = ctx->d_vmask.p;
                } else if (rc3 != HPGV_E_UNSUPPORTED) return rc3;
            }
            const Shape3 s3 = (mcap > 0 && !use_v3) ? pick_shape3(ctx, rank, mcap) : Shape3();
            if (s3.tj > 0) {
                int rc3 = build_worklist3(ctx, s3.tj, first, last);
                if (rc3 == HPGV_OK) {
                    use_v2 = true;
                    shape.nthreads = 2 * s3.tj * 32; shape.lists_in_smem = s3.lists_in_smem; shape.nstages = 2; shape.smem = s3.smem;
                    args.v2_ni = s3.ni; args.v2_mcap = mcap; args.v2_miss = ctx->d_miss.p;
                } else if (rc3 != HPGV_E_UNSUPPORTED) return rc3;
            }
        }
    }
    if (!use_v2 && !use_v3) shape = pick_shape(ctx, order, rank);
    if (shape.nthreads == 0)
        FAIL(HPGV_E_UNSUPPORTED, "fold count x cell count does not fit the shared memory of an SM (" + std::to_string(ctx->max_smem_optin) + " bytes)");
    int rc = (use_v2 || use_v3) ? HPGV_OK : build_worklist(ctx, order, shape.nthreads / 32, first, last);
    if (rc) return rc;
    args.unit_prefix = ctx->d_prefix.p;
    args.unit_jt0 = ctx->d_jt0.p;
    {
        const char *ud = getenv("HPGV_UNIT_DESC");            // A/B switch: walk the prefix table instead
        args.unit_desc = (ctx->wl_has_desc && !(ud && ud[0] == '0')) ? ctx->d_unit_desc.p : nullptr;
    }
    args.it0 = ctx->wl_it0;
    args.n_it = ctx->wl_nit;
    args.num_units = ctx->wl_units;
    args.edge_lo = ctx->wl_edge_lo;
    args.edge_hi = ctx->wl_edge_hi;

    // (balanced: the packed-pair epilogue needs A == U -- r = 1: the float32 rule is exact -- and 16-bit class sizes)
    {
        // score histogram: the pre-filter's conditions and the order-2 kernel's step modes
        const char *hs = getenv("HPGV_HIST");
        args.use_hist = (order == 2 && args.prefilter && !(hs && hs[0] == '0')) ? 1 : 0;
        args.hist_bins = fl.A + 1;
    }
    // the kernel variant of this layout (instantiated in epi_k_*.cu, one translation unit per kernel family)
    const int bwcode = (order == 2 && fl.tri) ? 3 : (fl.bw == 4 ? 4 : (fl.w7 ? 7 : 8));
    search_kernel_t kernel = order == 2 ? kernel_search2(bwcode, fl.single != 0, balanced)
                             : (use_v3 ? kernel_search3v3(bwcode, fl.single != 0)
                                       : (use_v2 ? kernel_search3v2(bwcode, fl.single != 0, balanced) : kernel_search3(bwcode, fl.single != 0, balanced)));
    if (!kernel) FAIL(HPGV_E_UNSUPPORTED, "no search kernel was built for this sample layout");
    const int grid = launch_search(ctx, kernel, shape, args, F, rank);
    if (grid < 0) return grid;

    MergeArgs m{};
    m.lists = ctx->d_lists.p; m.list_cnt = ctx->d_list_cnt.p;
    m.nlists = grid; m.F = F; m.rank_in = rank; m.rank_out = rank; m.training = args.training;
    m.fl = ctx->d_fl.p; m.out = d_out; m.order = order;
    return launch_merge(ctx, m);
}

extern "C" int hpgv_epi_search_device(hpgv_epi_ctx *ctx, int order, int eval_subset, int rank, uint64_t first, uint64_t last,
                                      hpgv_epi_model_t *d_out) {
    if (!ctx || !d_out) return HPGV_E_ARG;
    return do_search(ctx, order, eval_subset, rank, first, last, d_out);
}

extern "C" int hpgv_epi_search(hpgv_epi_ctx *ctx, int order, int eval_subset, int rank, uint64_t first, uint64_t last,
                               hpgv_epi_model_t *out) {
    if (!ctx || !out) return HPGV_E_ARG;
    if (!ctx->folds_set) FAIL(HPGV_E_STATE, "search before set_folds");
    if (rank < 1 || rank > kMaxRank) FAIL(HPGV_E_ARG, "rank_size must be in [1, 4096]");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t) ctx->fl.F * rank;
    CK(ctx->d_out.reserve(n));
    int rc = do_search(ctx, order, eval_subset, rank, first, last, ctx->d_out.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->d_out.p, n * sizeof(hpgv_epi_model_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HPGV_OK;
}

extern "C" int hpgv_epi_merge_device(hpgv_epi_ctx *ctx, int order, int eval_subset, int num_lists, int F, int rank,
                                     const hpgv_epi_model_t *d_lists, hpgv_epi_model_t *d_out) {
    if (!ctx || !d_lists || !d_out) return HPGV_E_ARG;
    if (!ctx->folds_set) FAIL(HPGV_E_STATE, "merge before set_folds (fold sizes are needed)");
    if (F != ctx->fl.F) FAIL(HPGV_E_ARG, "num_folds differs from the layout set by set_folds");
    if (num_lists < 1 || rank < 1 || rank > kMaxRank) FAIL(HPGV_E_ARG, "bad list shape");
    if (order != 2 && order != 3) FAIL(HPGV_E_ARG, "order must be 2 or 3");
    CK(cudaSetDevice(ctx->device));
    const int64_t n = (int64_t) num_lists * F * rank;
    CK(ctx->d_merge_in.reserve((size_t) n));
    models_to_cands_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const ModelOut *>(d_lists), n, ctx->d_merge_in.p);
    CK(cudaGetLastError());
    ctx->launches++;
    MergeArgs m{};
    m.lists = ctx->d_merge_in.p; m.list_cnt = nullptr;
    m.nlists = num_lists; m.F = F; m.rank_in = rank; m.rank_out = rank;
    m.training = (eval_subset == HPGV_SUBSET_TRAINING);
    m.fl = ctx->d_fl.p; m.out = d_out; m.order = order;
    return launch_merge(ctx, m);
}

extern "C" int hpgv_epi_merge_host(hpgv_epi_ctx *ctx, int order, int eval_subset, int num_lists, int F, int rank,
                                   const hpgv_epi_model_t *lists, hpgv_epi_model_t *out) {
    if (!ctx || !lists || !out) return HPGV_E_ARG;
    if (num_lists < 1 || F < 1 || rank < 1 || rank > kMaxRank) FAIL(HPGV_E_ARG, "bad list shape");
    CK(cudaSetDevice(ctx->device));
    const size_t nin = (size_t) num_lists * F * rank, nout = (size_t) F * rank;
    hpgv_epi_model_t *d = nullptr;
    CK(cudaMalloc(&d, (nin + nout) * sizeof(hpgv_epi_model_t)));
    cudaError_t e = cudaMemcpyAsync(d, lists, nin * sizeof(hpgv_epi_model_t), cudaMemcpyHostToDevice, ctx->stream);
    int rc = HPGV_OK;
    if (e == cudaSuccess) rc = hpgv_epi_merge_device(ctx, order, eval_subset, num_lists, F, rank, d, d + nin);
    if (e == cudaSuccess && rc == HPGV_OK) e = cudaMemcpyAsync(out, d + nin, nout * sizeof(hpgv_epi_model_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (rc) return rc;
    if (e != cudaSuccess) FAIL(HPGV_E_CUDA, std::string("merge_host: ") + cudaGetErrorString(e));
    return HPGV_OK;
}

extern "C" int hpgv_epi_device_count(void) {
    int n = 0;
    const cudaError_t e = device_count_retry(&n);
    if (e != cudaSuccess) { g_create_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e); return 0; }
    return n;
}

// ---------------------------------------------------------------------------------
// parity hooks
// ---------------------------------------------------------------------------------
static int eval_impl(hpgv_epi_ctx *ctx, int order, int eval_subset, int64_t ncomb, const int32_t *combs, const uint32_t *risky_in,
                     int32_t *counts_aff, int32_t *counts_unaff, uint32_t *risky_mask, uint32_t *conf, double *accuracy) {
    if (!ctx || !combs || ncomb < 0) return HPGV_E_ARG;
    if (!ctx->folds_set) FAIL(HPGV_E_STATE, "eval before set_folds");
    if (order != 2 && order != 3) FAIL(HPGV_E_ARG, "order must be 2 or 3");
    if (ncomb == 0) return HPGV_OK;
    for (int64_t c = 0; c < ncomb; c++)
        for (int o = 0; o < order; o++) {
            const int32_t v = combs[c * order + o];
            if (v < 0 || v >= ctx->nv || (o > 0 && v <= combs[c * order + o - 1])) FAIL(HPGV_E_ARG, "combination indices must be ascending and inside the dataset");
        }
    CK(cudaSetDevice(ctx->device));
    // the per-combination dump reads logical words: an order-3 dump of a tri-packed dataset is fine (logical_word serves every layout)
    const int F = ctx->fl.F, C = order == 2 ? 9 : 27;
    const size_t nf = (size_t) ncomb * F;
    int32_t *d_combs = nullptr, *d_ca = nullptr, *d_cu = nullptr;
    uint32_t *d_mask = nullptr, *d_mask_in = nullptr, *d_conf = nullptr;
    double *d_acc = nullptr;
    auto cleanup = [&]() { cudaFree(d_combs); cudaFree(d_ca); cudaFree(d_cu); cudaFree(d_mask); cudaFree(d_mask_in); cudaFree(d_conf); cudaFree(d_acc); };
#define CKE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return HPGV_E_CUDA; } } while (0)
    CKE(cudaMalloc(&d_combs, (size_t) ncomb * order * sizeof(int32_t)));
    CKE(cudaMalloc(&d_ca, nf * C * sizeof(int32_t)));
    CKE(cudaMalloc(&d_cu, nf * C * sizeof(int32_t)));
    CKE(cudaMalloc(&d_mask, nf * sizeof(uint32_t)));
    CKE(cudaMalloc(&d_conf, nf * 4 * sizeof(uint32_t)));
    CKE(cudaMalloc(&d_acc, nf * sizeof(double)));
    CKE(cudaMemcpyAsync(d_combs, combs, (size_t) ncomb * order * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (risky_in) {
        CKE(cudaMalloc(&d_mask_in, nf * sizeof(uint32_t)));
        CKE(cudaMemcpyAsync(d_mask_in, risky_in, nf * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    const int warps = 4;
    const size_t smem = (size_t) warps * ctx->fl.nseg * C * sizeof(int);
    const unsigned grid = (unsigned) ((ncomb + warps - 1) / warps);
    const int training = (eval_subset == HPGV_SUBSET_TRAINING);
    eval_kernel<<<grid, warps * 32, smem, ctx->stream>>>(ctx->d_planes.p, ctx->d_blk.p, ctx->d_fl.p, ctx->snp_pad, order, training,
                                                         ctx->eval_fn == kEvalCA ? kEvalBA : ctx->eval_fn, ncomb, d_combs, d_mask_in,
                                                         d_ca, d_cu, d_mask, d_conf, d_acc);
    CKE(cudaGetLastError());
    ctx->launches++;
    if (counts_aff) CKE(cudaMemcpyAsync(counts_aff, d_ca, nf * C * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (counts_unaff) CKE(cudaMemcpyAsync(counts_unaff, d_cu, nf * C * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (risky_mask) CKE(cudaMemcpyAsync(risky_mask, d_mask, nf * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (conf) CKE(cudaMemcpyAsync(conf, d_conf, nf * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (accuracy) CKE(cudaMemcpyAsync(accuracy, d_acc, nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CKE(cudaStreamSynchronize(ctx->stream));
#undef CKE
    cleanup();
    return HPGV_OK;
}

extern "C" int hpgv_epi_eval(hpgv_epi_ctx *ctx, int order, int eval_subset, int64_t ncomb, const int32_t *combs,
                             int32_t *counts_aff, int32_t *counts_unaff, uint32_t *risky_mask, uint32_t *conf, double *accuracy) {
    return eval_impl(ctx, order, eval_subset, ncomb, combs, nullptr, counts_aff, counts_unaff, risky_mask, conf, accuracy);
}

extern "C" int hpgv_epi_confusion(hpgv_epi_ctx *ctx, int order, int eval_subset, int64_t ncomb, const int32_t *combs,
                                  const uint32_t *risky_mask_in, uint32_t *conf, double *accuracy) {
    if (!risky_mask_in) return HPGV_E_ARG;
    return eval_impl(ctx, order, eval_subset, ncomb, combs, risky_mask_in, nullptr, nullptr, nullptr, conf, accuracy);
}

// explicit count pairs / confusion matrices through the device functions the search kernels use
extern "C" int hpgv_epi_high_risk(hpgv_epi_ctx *ctx, const int32_t *counts_aff, const int32_t *counts_unaff, int64_t n,
                                  int num_affected, int num_unaffected, int32_t *flags) {
    if (!ctx || !counts_aff || !counts_unaff || !flags || n < 0) return HPGV_E_ARG;
    if (num_affected < 1 || num_unaffected < 1) FAIL(HPGV_E_ARG, "need at least one affected and one unaffected sample");
    if (n == 0) return HPGV_OK;
    CK(cudaSetDevice(ctx->device));
    int32_t *d = nullptr;
    CK(cudaMalloc(&d, (size_t) n * 3 * sizeof(int32_t)));
    cudaError_t e = cudaMemcpyAsync(d, counts_aff, (size_t) n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, counts_unaff, (size_t) n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        high_risk_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(d, d + n, n, num_affected, num_unaffected, d + 2 * n);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(flags, d + 2 * n, (size_t) n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) FAIL(HPGV_E_CUDA, std::string("high_risk: ") + cudaGetErrorString(e));
    return HPGV_OK;
}

extern "C" int hpgv_epi_evaluate(hpgv_epi_ctx *ctx, int eval_function, int64_t n, const uint32_t *conf, double *values) {
    if (!ctx || !conf || !values || n < 0) return HPGV_E_ARG;
    if (eval_function == HPGV_EVAL_WBA) FAIL(HPGV_E_UNSUPPORTED, "wBA is declared but not implemented by the reference (model.h:84)");
    if (eval_function < HPGV_EVAL_CA || eval_function > HPGV_EVAL_CA_TRUE) FAIL(HPGV_E_ARG, "unknown evaluation function");
    if (n == 0) return HPGV_OK;
    CK(cudaSetDevice(ctx->device));
    uint32_t *d_conf = nullptr;
    double *d_val = nullptr;
    CK(cudaMalloc(&d_conf, (size_t) n * 4 * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_val, (size_t) n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_conf, conf, (size_t) n * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        evaluate_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(eval_function == kEvalCA ? kEvalBA : eval_function, n, d_conf, d_val);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, d_val, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_conf); cudaFree(d_val);
    if (e != cudaSuccess) FAIL(HPGV_E_CUDA, std::string("evaluate: ") + cudaGetErrorString(e));
    return HPGV_OK;
}

extern "C" int hpgv_epi_unpack_masks(hpgv_epi_ctx *ctx, int64_t variant, uint8_t *out) {
    if (!ctx || !out) return HPGV_E_ARG;
    if (!ctx->folds_set) FAIL(HPGV_E_STATE, "unpack before set_folds");
    if (variant < 0 || variant >= ctx->nv) FAIL(HPGV_E_ARG, "variant out of range");
    CK(cudaSetDevice(ctx->device));
    const int a_pad = 16 * ((ctx->A + 15) / 16), u_pad = 16 * ((ctx->U + 15) / 16), s_pad = a_pad + u_pad;
    uint8_t *d_out = nullptr;
    CK(cudaMalloc(&d_out, (size_t) 3 * s_pad));
    cudaMemsetAsync(d_out, 0, (size_t) 3 * s_pad, ctx->stream);
    const unsigned grid = (unsigned) ((ctx->npos + 255) / 256);
    unpack_masks_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_planes.p, variant, ctx->snp_pad, ctx->d_perm.p, ctx->d_fl.p, ctx->npos, ctx->A, a_pad, s_pad, d_out);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, d_out, (size_t) 3 * s_pad, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) FAIL(HPGV_E_CUDA, std::string("unpack: ") + cudaGetErrorString(e));
    return HPGV_OK;
}

extern "C" int hpgv_epi_run_host(hpgv_epi_ctx *ctx, const uint8_t *genotypes, int64_t nv, int A, int U, int F,
                                 const int32_t *fold_of_sample, int order, int eval_subset, int rank,
                                 uint64_t first, uint64_t last, hpgv_epi_model_t *out) {
    int rc = hpgv_epi_load_dataset_host(ctx, genotypes, nv, A, U);
    if (rc) return rc;
    rc = hpgv_epi_set_folds(ctx, F, fold_of_sample);
    if (rc) return rc;
    return hpgv_epi_search(ctx, order, eval_subset, rank, first, last, out);
}

extern "C" int hpgv_epi_last_search_ms(hpgv_epi_ctx *ctx, float *ms, int *grid) {
    if (!ctx || !ms) return HPGV_E_ARG;
    if (ctx->ev_count == 0) FAIL(HPGV_E_STATE, "no search has been launched yet");
    const int slot = (int) ((ctx->ev_count - 1) % hpgv_epi_ctx::kEvRing);
    CK(cudaEventSynchronize(ctx->ev1[slot]));
    CK(cudaEventElapsedTime(ms, ctx->ev0[slot], ctx->ev1[slot]));
    if (grid) *grid = ctx->last_grid;
    return HPGV_OK;
}

extern "C" int hpgv_epi_search_times(hpgv_epi_ctx *ctx, int n, float *ms) {
    if (!ctx || !ms || n < 0) return HPGV_E_ARG;
    n = (int) std::min<int64_t>(std::min<int64_t>(n, hpgv_epi_ctx::kEvRing), ctx->ev_count);
    for (int k = 0; k < n; k++) {
        const int slot = (int) ((ctx->ev_count - n + k) % hpgv_epi_ctx::kEvRing);
        CK(cudaEventSynchronize(ctx->ev1[slot]));
        CK(cudaEventElapsedTime(ms + k, ctx->ev0[slot], ctx->ev1[slot]));
    }
    return n;
}

extern "C" int hpgv_epi_debug_counters(hpgv_epi_ctx *ctx, uint64_t *out, int n) {
    if (!ctx || !out || n < 0) return HPGV_E_ARG;
    if (!ctx->d_dbg.p) FAIL(HPGV_E_STATE, "no search has run with HPGV_DEBUG_COUNTERS=1");
    n = std::min<int>(n, (int) ctx->d_dbg.cap);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->d_dbg.p, (size_t) n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return n;
}

extern "C" int hpgv_epi_layout(const hpgv_epi_ctx *ctx, hpgv_epi_layout_t *out) {
    if (!ctx || !out) return HPGV_E_ARG;
    if (!ctx->folds_set) return HPGV_E_STATE;
    out->num_folds = ctx->fl.F; out->num_segments = ctx->fl.nseg; out->num_blocks = ctx->fl.nblocks; out->block_words = ctx->fl.tri ? 3 : ctx->fl.bw;
    out->num_chunks = ctx->fl.nchunks; out->chunk_blocks = ctx->fl.cb; out->row_words = ctx->fl.row_words;
    out->count_bits = ctx->fl.single ? 8 : 16;
    out->plane_bytes = (int64_t) ctx->plane_words * 4;
    out->words_per_class_row = (ctx->A + 31) / 32 + (ctx->U + 31) / 32;
    return HPGV_OK;
}

extern "C" int hpgv_epi_pipe_peak(hpgv_epi_ctx *ctx, int kind, int iters, double *ops_per_second) {
    if (!ctx || !ops_per_second || iters < 1 || kind < 0 || kind > 2) return HPGV_E_ARG;
    CK(cudaSetDevice(ctx->device));
    uint32_t *d_sink = nullptr;
    CK(cudaMalloc(&d_sink, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 256, grid = ctx->num_sms * 8;
    auto run = [&]() {
        if (kind == 0) pipe_peak_kernel<0><<<grid, threads, 0, ctx->stream>>>(iters, 12345u, d_sink);
        else if (kind == 1) pipe_peak_kernel<1><<<grid, threads, 0, ctx->stream>>>(iters, 12345u, d_sink);
        else pipe_peak_kernel<2><<<grid, threads, 0, ctx->stream>>>(iters, 12345u, d_sink);
    };
    run();   // warm-up
    cudaEventRecord(e0, ctx->stream);
    run();
    cudaEventRecord(e1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_sink);
    ctx->launches += 2;
    if (e != cudaSuccess) FAIL(HPGV_E_CUDA, std::string("pipe_peak: ") + cudaGetErrorString(e));
    // ops counted per inner step: kind 0 -> 1 POPC, kind 1 -> 2 LOP3, kind 2 -> 2 LOP3 + 1 POPC (reported as 3)
    const double per_step = kind == 0 ? 1.0 : (kind == 1 ? 2.0 : 3.0);
    *ops_per_second = per_step * 64.0 * (double) iters * threads * (double) grid / (ms * 1e-3);
    return HPGV_OK;
}
