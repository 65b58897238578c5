// epi_types.h -- structures shared by host and device code of the epistasis engine.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace hpgv {

constexpr int kMaxFolds  = 32;            // HPGV_MAX_FOLDS
constexpr int kMaxSegs   = 2 * kMaxFolds; // one segment per (fold, class)
constexpr int kMaxBlocks = 4096;          // sample-axis blocks per SNP row (256 samples each when segments span blocks)
constexpr int kMaxRank   = 4096;          // HPGV_MAX_RANK
constexpr int kTileJ     = 32;            // one j (or k) SNP per lane
constexpr int kMaxWarps  = 16;            // consumer warps per CTA = i (or j) rows per tile
constexpr int kTriWarps  = 20;            // the order-2 kernels with byte counters fit 96 registers: five warps per SM sub-partition

// Per-fold sizes and the segmented sample layout chosen by set_folds().
//
// Samples are permuted to (fold, class) order: segment s = 2*fold + cls
// (cls 0 = affected, 1 = unaffected).  A segment is a run of blocks; a block is
// bw 32-bit words of each of the three genotype planes; every bit position
// holds one sample of that segment or padding (never set in any plane).
// Because a sample is in exactly one fold, the per-fold TRAINING table of the
// reference (model.c:131-206) is total - in-fold.
//
// Blocks are grouped into chunks of cb blocks; the packed planes are stored
// chunk-major, [chunk][snp][row_words] with row_words = cb*3*bw (+ padding so
// that row_words/4 is odd: 32 lanes reading 32 consecutive rows with LDS.128
// then hit 8 distinct 16-byte bank groups per quarter-warp).  Inside a row:
// word (bl, g, w) at (bl*3 + g)*bw + w  (tri layout: see tri_word_off / tri_tail_off).
struct FoldLayout {
    int F;                 // folds
    int nseg;              // 2F
    int A, U;              // dataset-level class sizes (used by the high-risk rule, epistasis.c:37)
    int balanced;          // A == U: the float32 rule collapses to an integer test
    int eqfolds;           // every fold holds as many cases as controls (a_in[f] == u_in[f]): with A == U the training score
                           // of a fold is n_f * sum over cells of max(0, trA - trU), which the search kernels pre-filter on
    float ratio;           // (float)A / (float)U, mdr.c:52
    int bw;                // words per plane per block (4 or 8)
    int w7;                // bw == 8 and only the first 7 words of a block hold samples (7 words compress to 3 POPC, not 4)
    int single;            // 1: every segment is exactly one block (block b <-> segment b, byte counters)
    int tri;               // single && every segment <= 100 samples && <= 24 blocks: three full words per plane and block plus a
                           // 4-bit tail per segment, the tails of eight segments sharing one word (see tri_* in epi_device.cuh):
                           // 3 words compress to 2 POPC, the tails are counted with nibble-wise adds on the ALU.  Logical
                           // bit positions stay those of bw = 4 (word 3 of a block = its tail, bits 0..3).
    int marg;              // rows carry, per group of four blocks, one 16-byte quad (N_0, N_1, N_2, missing): the SNP's own
                           // per-block genotype counts as packed byte counters and 0xFF in the byte of every block where the
                           // SNP has a sample in no plane (tri layout: always; single-block layouts: when the staged packer runs)
    int marg_off;          // word offset of the first quad inside a chunk row
    int marg_stride;       // words per group of four blocks in the marginal area: 4 (the quad), or 4 + kMissListWords when
    int mlist;             // the quad is followed by the group's list of the SNP's missing samples (single-block layouts
                           // other than tri): entries of 16 sample bits each, see kMissListWords in epi_kernels.cuh.  The
                           // search then derives genotype 2 of SNP i in EVERY block and takes the listed samples out of the
                           // derived cells one entry at a time; `missing` (word 3 of the quad) only marks the blocks of a
                           // group whose list overflowed, which are counted directly as before
    int nblocks;           // real blocks along the sample axis (single: nseg rounded up to a multiple of 4)
    int cb;                // blocks per chunk (single: multiple of 4)
    int nchunks;
    int row_words;         // words per chunk row
    int a_in[kMaxFolds];   // cases in fold f (its testing part)
    int u_in[kMaxFolds];
};

#if defined(__CUDACC__)
__host__ __device__
#endif
inline int hist_coarse_bins(int hist_bins) { return (hist_bins + 31) / 32; }

// Candidate / ranked entry as kept on the device (32 bytes, two 16-byte words).
struct __attribute__((aligned(16))) Cand {
    double ba;      // balanced accuracy; -inf encodes NaN (ranks last)
    int32_t i, j, k;
    uint32_t mask;
    int32_t tp, fp;
};
static_assert(sizeof(Cand) == 32, "Cand must be 32 bytes");

// 128-bit ranking key: the canonical order (accuracy descending, then SNP tuple ascending) as "larger key first".
struct __attribute__((aligned(16))) Key128 {
    unsigned long long lo, hi;  // hi = order-preserving image of the accuracy, lo = ~tuple
};

// development trace (SearchArgs::dbg): offered tuples are appended after the counters and the per-CTA clocks
constexpr unsigned long long kDbgTraceCount = 8 + 2 * 1024, kDbgTrace = kDbgTraceCount + 8, kDbgTraceCap = 1ULL << 20;
constexpr unsigned long long kDbgWords = kDbgTrace + 2 * kDbgTraceCap;

// Arguments of the order-2 / order-3 search kernels.
struct SearchArgs {
    const uint32_t *planes;     // [nchunks][snp_pad][row_words]
    const uint16_t *blk_desc;   // [nblocks] segment id | 0x8000 when last block of its segment
    const FoldLayout *fl;
    int64_t snp_pad;            // padded SNP rows per chunk
    int nv;                     // real SNP count
    int training;               // evaluate on the training (1) or testing (0) part
    int rank;                   // N
    int lists_in_smem;          // per-CTA top-N lists live in shared memory during the search
    uint64_t first, last;       // linear combination index range
    int edge_lo, edge_hi;       // first SNP of the tuples `first` and `last - 1`: tuples whose first SNP lies strictly between
                                // are inside the range, only the others are checked index by index
    // work list: order 2 -> unit = (i-tile, j-tile); order 3 -> unit = (i, j-tile); prefix[t] = units before row-group t
    const int64_t *unit_prefix;
    const int32_t *unit_jt0;    // first j-tile of each row group
    const int2 *unit_desc;      // order 2, when the list fits: tile origin (i0, j0) of every unit -- one load instead of a walk through
                                // the prefix table (whose dependent loads sat on the critical path of the producer warp)
    int it0, n_it;              // first row group, number of row groups
    int64_t num_units;
    // per-CTA candidate lists
    Cand *lists;                // [grid][F][rank]
    int *list_cnt;              // [grid][F]
    long long *gthr;            // [F] global score threshold (lower bound of the N-th best)
    unsigned long long *dbg;    // development counters (HPGV_DEBUG_COUNTERS=1, read back with hpgv_epi_debug_counters), else nullptr:
                                // [0] warps whose pre-filter passed, [1] lanes offered (score >= bound), [2] lanes left after the
                                // root snapshot, [3] list entries written, [4] lock spins, [5] hist look-ups, [8 + 2b], [9 + 2b]: CTA b's
                                // clock64 at its first step and at its end
    int *gfirst;                // CTAs whose first unit is in the histogram
    int first_wait;             // a CTA's second unit waits for gfirst == grid (HPGV_FIRST_WAIT=0: no wait, the look-up races)
    int fresh_bound;            // adopt the best bound any CTA has published once per unit (HPGV_FRESH_BOUND=0: round-1 behaviour)
    // score histogram (balanced TRAINING searches with equal folds, order 2): ghist[f][t] counts the pairs seen so far
    // whose pre-filter score of fold f is t; the N-th best score any CTA can derive from it bounds every list
    int *ghist;                 // [F][hist_bins] fine bins, then [F][hist_coarse_bins(hist_bins)] bins of 32 scores
    int *ghmax;                 // [F] largest score counted so far (-1: none)
    int hist_bins;              // A + 1
    int use_hist;
    // search3v2_kernel: rows of NI SNPs i per stage, per-SNP lists of missing samples (v2_mcap entries each)
    int v2_ni, v2_mcap;
    const uint32_t *v2_miss;    // [snp_pad][v2_mcap], see kMissEnd
    const uint32_t *v3_vmask;   // search3v3_kernel: [nblocks][slot words] the bit positions of every block that hold a sample
    int eval_fn;                // enum eval_function of model.h:84 (kEval* in epi_device.cuh); 1 = BA, the reference runner's choice (model.c:331)
    int prefilter;              // balanced classes, equal folds, TRAINING part and BA: the pre-filter of epilogue_balanced_t applies
    int tri_derive;             // derive genotype 2 of SNP i from SNP j's marginals in blocks where i has no missing sample (rows with marg)
    int list_scan;              // short per-CTA lists (N <= 64) are plain arrays, one lock per warp and fold (offer_batch_scan); HPGV_LIST_SCAN
    int nstages;                // shared-memory stages of the search kernel's ring (2 or 3)
    int stagger;                // 1: half of each sub-partition's warps starts half a unit late, 2: evenly spread phases, 0: off (HPGV_STAGGER)
};

}  // namespace hpgv
