// epi_types.h -- structures shared by host and device code of the epistasis engine.
#pragma once
#include <stdint.h>

namespace hpgv {

constexpr int kMaxFolds  = 32;            // HPGV_MAX_FOLDS
constexpr int kMaxSegs   = 2 * kMaxFolds; // one segment per (fold, class)
constexpr int kMaxBlocks = 8192;          // sample-axis blocks per SNP row
constexpr int kMaxRank   = 4096;          // HPGV_MAX_RANK
constexpr int kTileJ     = 32;            // one j (or k) SNP per lane

// Per-fold sizes and the segmented sample layout chosen by set_folds().
//
// Samples are permuted to (fold, class) order: segment s = 2*fold + cls
// (cls 0 = affected, 1 = unaffected) is a run of `blocks` of BW 32-bit words;
// every bit position holds one sample of that segment or padding (never set in
// any plane).  Because a sample is in exactly one fold, the per-fold TRAINING
// table of the reference (model.c:131-206) is total - in-fold.
struct FoldLayout {
    int F;                 // folds
    int nseg;              // 2F
    int nblocks;           // total blocks along the sample axis
    int bw;                // words per block (4 or 8)
    int A, U;              // dataset-level class sizes (used by the high-risk rule, epistasis.c:37)
    int balanced;          // A == U: the float32 rule collapses to an integer test
    float ratio;           // (float)A / (float)U, mdr.c:52
    int a_in[kMaxFolds];   // cases in fold f (its testing part)
    int u_in[kMaxFolds];
};

// Candidate / ranked entry as kept on the device (32 bytes, two 16-byte words).
struct __attribute__((aligned(16))) Cand {
    double ba;      // balanced accuracy; -inf encodes NaN (ranks last)
    int32_t i, j, k;
    uint32_t mask;
    int32_t tp, fp;
};
static_assert(sizeof(Cand) == 32, "Cand must be 32 bytes");

// Arguments of the order-2 / order-3 search kernels.
struct SearchArgs {
    const uint32_t *planes;     // [nblocks][snp_pad][3][bw]
    const uint16_t *blk_desc;   // [nblocks] segment id | 0x8000 when last block of its segment
    const FoldLayout *fl;
    int64_t snp_pad;            // padded SNP rows per block
    int nv;                     // real SNP count
    int training;               // evaluate on the training (1) or testing (0) part
    int rank;                   // N
    uint64_t first, last;       // linear combination index range
    // work list: order 2 -> unit = (i-tile, j-tile); prefix[t] = units before i-tile t0+t
    const int64_t *unit_prefix;
    const int32_t *unit_jt0;    // first j-tile of each i-tile
    int it0, n_it;              // first i-tile, number of i-tiles
    int64_t num_units;
    unsigned long long *unit_counter;
    // per-CTA candidate lists
    Cand *lists;                // [grid][F][rank]
    int *list_cnt;              // [grid][F]
    long long *gthr;            // [F] global score threshold (lower bound of the N-th best)
};

}  // namespace hpgv
