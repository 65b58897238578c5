// epi_kernels.cuh -- CUDA kernels of the epistasis engine (sm_100a).
// (Included by epi_capi.cu and by one epi_k_*.cu per search-kernel family, which build.py compiles in parallel: everything
// that is not a template has internal linkage.)
//
//   pack_rows_kernel     bytes -> (fold, class)-segmented bit planes, staged through shared memory, with the per-block
//                        marginals and missing masks of the byte-counter layouts  (replaces set_genotypes_masks + get_k_folds_masks)
//   pack_planes_kernel   the same, one warp per word with gathered global loads (sample axes too long for shared memory)
//   search2_kernel       exhaustive order-2 MDR: counts, risk, BA, top-N      (replaces process_set_of_combinations)
//   search3_kernel       same for order 3
//   merge_kernel         per-CTA / per-rank top-N lists -> final top-N        (replaces the heap drain + MPI tree merge)
//   eval_kernel          per-combination dump for explicit combinations       (parity hook)
//
// Design of the search kernels (DESIGN.md has the long form).  One persistent CTA
// per SM walks a static round-robin share of the work units.  A unit is a tile of
// SNP tuples: warp <-> one row of the tile (i for order 2, j for order 3), lane <->
// the last SNP of the tuple, so the lane's own planes are one conflict-free LDS.128
// per plane and the other SNPs' planes are warp-uniform broadcast loads.  The sample
// axis is cut into chunks; for every (unit, chunk) step thread 0 issues two or three
// 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) that land the chunk rows of
// the tile in one of two shared-memory stages and complete on an mbarrier, one step
// ahead of the compute.  A thread ANDs the planes of its tuple cell by cell,
// compresses each block with LOP3 carry-save adders (ALU pipe), POPCs the compressed
// words (XU pipe) and accumulates the weighted counts with IMAD (FMA pipe) straight
// into packed per-segment counters (four byte counters or two 16-bit counters per
// word) that go to the thread's private slice of shared memory once per four blocks
// / once per segment.  After the last chunk the thread derives, per fold, the
// training table (total - in-fold), the exact high-risk mask, TP/FP and an integer
// score, and offers tuples that beat the running threshold to the CTA's top-N list.
//
// Three things keep the order-2 kernels off the obvious costs: (1) in a block where SNP i has no missing sample the cells
// of its genotype 2 follow from SNP j's own counts and the cells of genotypes 0 and 1 (derive_row2), so a third of the
// AND/POPC work is skipped in most blocks; (2) the tri layout packs segments of <= 100 samples as three words plus a 4-bit
// tail, 2 POPCs per block instead of 3; (3) a global histogram of pre-filter scores gives every CTA the bound of ALL pairs
// seen so far (hist_threshold), so the per-CTA lists see a few thousand offers per search instead of half a million.
#pragma once
#include <type_traits>
#include "epi_device.cuh"

namespace hpgv {

// ============================================================================
// Packing
// ============================================================================
// One warp per (snp, block, word): lane l reads the genotype byte of the sample
// mapped to bit l and three ballots produce the three plane words.
// perm[pos] = dataset column of the sample at bit position pos, or -1 (padding).
static __global__ void pack_planes_kernel(const uint8_t *__restrict__ raw, int64_t nv, int64_t nsamples,
                                   const int32_t *__restrict__ perm, const FoldLayout *__restrict__ flp, int64_t snp_pad,
                                   uint32_t *__restrict__ planes) {
    const FoldLayout &fl = *flp;
    const int bw = fl.bw;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t words_per_snp = (int64_t) fl.nblocks * bw;
    if (warp >= nv * words_per_snp) return;
    const int64_t snp = warp / words_per_snp;
    const int wb = (int) (warp % words_per_snp);
    const int b = wb / bw, w = wb % bw;
    const int32_t col = perm[(int64_t) wb * 32 + lane];
    const uint32_t g = col >= 0 ? raw[snp * nsamples + col] : 255u;
    const uint32_t m0 = __ballot_sync(0xffffffffu, g == 0);
    const uint32_t m1 = __ballot_sync(0xffffffffu, g == 1);
    const uint32_t m2 = __ballot_sync(0xffffffffu, g == 2);
    const uint32_t mine = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
    if (lane >= 3) return;
    if (fl.tri) {
        // the planes were zeroed: the tails of eight blocks are OR-ed into one word
        uint32_t *row = planes + snp * fl.row_words;
        if (w < 3) row[tri_word_off(b, lane, w)] = mine;
        else if (mine & 0xFu) atomicOr(row + tri_tail_off(fl.nblocks, b, lane), (mine & 0xFu) << tri_tail_shift(b));
    } else {
        planes[plane_word(fl, snp_pad, b, snp, lane, w)] = mine;
    }
}

// The same through shared memory, for sample axes whose permutation fits there (all but the largest cohorts): a CTA
// takes one SNP at a time, reads its bytes with coalesced 16-byte loads, gathers them through the permutation kept in
// shared memory, assembles the packed row(s) there and writes them out coalesced.  HBM-bound by design: nsamples bytes
// in, 3 x nsamples / 8 (+ padding) out per SNP.
struct PackSmem {
    size_t perm, ofs, row, out, total;
};
// list of a SNP's missing samples in one group of four blocks (FoldLayout::mlist), kMissListWords entries, 0 = end:
//   [15:0] the missing samples among 16 consecutive bit positions   [20:16] left shift that puts them back in their word (0 or 16)
//   [27:21] word offset of plane 0 of that word inside the group: block-in-group * 3 * slot words + word   [29:28] block in group
constexpr int kMissListWords = 8;
constexpr int kPackRows = 8;      // SNPs a CTA packs per iteration (their rows are contiguous on both sides)
__host__ __device__ inline PackSmem pack_smem_map(int64_t npos, int64_t nsamples, const FoldLayout &fl) {
    PackSmem m;
    m.perm = 0;
    m.ofs = (size_t) npos * 4;                                           // two ints per logical word (npos / 32 of them)
    m.row = m.ofs + (((size_t) (npos / 32) * 8 + 15) / 16) * 16;
    m.out = m.row + (((size_t) nsamples * kPackRows + 15) / 16 + 1) * 16;   // + 16: staged at the global misalignment
    m.total = m.out + (size_t) kPackRows * fl.nchunks * fl.row_words * 4;
    return m;
}
// TRI: the tri layout (c2 and its weak-scaled versions) with the loops over the four blocks of a group and their four logical
// words unrolled and every offset a compile-time expression: the generic path spends 69 warp instructions per 32-sample word
// (tables of offsets, tail / marginal / list bookkeeping on every lane), which made the packer issue-bound.
template <bool TRI>
static __global__ void __launch_bounds__(256) pack_rows_kernel(const uint8_t *__restrict__ raw, int64_t nv, int64_t nsamples,
                                                        const int32_t *__restrict__ perm, const FoldLayout *__restrict__ flp,
                                                        int64_t snp_pad, int64_t npos, uint32_t *__restrict__ planes) {
    extern __shared__ __align__(16) uint8_t psm[];
    const PackSmem m = pack_smem_map(npos, nsamples, *flp);
    int32_t *perm_s = reinterpret_cast<int32_t *>(psm + m.perm);
    int2 *ofs_s = reinterpret_cast<int2 *>(psm + m.ofs);
    uint8_t *row_s = psm + m.row;
    uint32_t *out_s = reinterpret_cast<uint32_t *>(psm + m.out);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int bw = flp->bw, lbw = bw == 8 ? 3 : 2, nwords = flp->nblocks * bw, tri = flp->tri, cb = flp->cb, nblocks = flp->nblocks;
    const int row_words = flp->row_words, nchunks = flp->nchunks, marg = flp->marg, marg_off = flp->marg_off, mstride = flp->marg_stride, mlist = flp->mlist;
    const int out_words = nchunks * row_words;                           // per SNP, chunk-major
    const int gstride = tri ? 12 : bw;                                   // words between the planes of one block
    const int bpt = flp->single ? 4 : 1;                                 // blocks per task (nblocks is a multiple of 4 then)
    const int tasks_per_row = nblocks / bpt;
    for (int64_t x = tid; x < npos; x += blockDim.x) perm_s[x] = perm[x];
    // x: where logical word wb of plane 0 goes inside the staged output of one SNP; tri tails: -(offset + 1) | shift << 24
    // y: the marginal quad of its block | byte shift of the block's counter << 24 | 1 << 30 when the word is a tri tail
    for (int wb = tid; wb < (TRI ? 0 : nwords); wb += blockDim.x) {
        const int b = wb >> lbw, w = wb & (bw - 1);
        int o;
        if (tri) o = w < 3 ? tri_word_off(b, 0, w) : -((tri_tail_off(nblocks, b, 0) + 1) | (tri_tail_shift(b) << 24));
        else o = (b / cb) * row_words + ((b % cb) * 3) * bw + w;
        const int mq = tri ? marg_off + (b >> 2) * 4 : (b / cb) * row_words + marg_off + ((b % cb) >> 2) * mstride;
        ofs_s[wb] = make_int2(o, mq | ((int) group_shift(b & 3) << 24) | ((tri && w == 3) ? (1 << 30) : 0));
    }
    for (int64_t snp0 = (int64_t) blockIdx.x * kPackRows; snp0 < nv; snp0 += (int64_t) gridDim.x * kPackRows) {
        const int nr = (int) min((int64_t) kPackRows, nv - snp0);
        // stage the rows: 16-byte loads from the aligned address at or below the first byte
        const uint8_t *src = raw + snp0 * nsamples;
        const int64_t nbytes = (int64_t) nr * nsamples;
        const int mis = (int) (reinterpret_cast<uintptr_t>(src) & 15);
        // a 16-byte group may reach up to 15 bytes outside the batch at either end: it must still lie inside the matrix
        // (whose ends are only byte aligned: a caller-owned device pointer), else the rows are read bytewise
        const int64_t before = snp0 * nsamples, after = (nv - snp0 - nr) * nsamples;
        if ((after >= 15 || ((nbytes + mis) & 15) == 0) && before >= mis) {
            const uint4 *src16 = reinterpret_cast<const uint4 *>(src - mis);
            const int n16 = (int) ((nbytes + mis + 15) / 16);
            for (int x = tid; x < n16; x += blockDim.x) reinterpret_cast<uint4 *>(row_s)[x] = __ldg(src16 + x);
        } else {
            for (int64_t x = tid; x < nbytes; x += blockDim.x) row_s[mis + x] = src[x];
        }
        for (int x = tid; x < nr * out_words; x += blockDim.x) out_s[x] = 0;
        __syncthreads();
        // one warp per (SNP, counter group of four blocks) when the rows carry marginals, else per (SNP, block): the
        // group's own genotype counts, its missing mask and the tri tails are assembled in registers (lane g = plane g)
        for (int task = warp; task < nr * tasks_per_row; task += nwarps) {
            const int r = task / tasks_per_row, tg = task - r * tasks_per_row;
            uint32_t *dst = out_s + r * out_words;
            const uint8_t *row = row_s + mis + (size_t) r * nsamples;
            if constexpr (TRI) {
                // group tg = blocks 4 tg .. 4 tg + 3; logical word (q, w) sits at bit positions ((4 tg + q) * 4 + w) * 32 ..
                const int32_t *pp = perm_s + (size_t) tg * 512 + lane;
                uint32_t *grp = dst + tg * 36;
                uint32_t nacc = 0, missacc = 0, tailacc = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t n = 0;
                    bool miss_here = false;
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        const int32_t col = pp[(q * 4 + w) * 32];
                        const uint32_t g = col >= 0 ? row[col] : 255u;
                        const uint32_t m0 = __ballot_sync(0xffffffffu, g == 0);
                        const uint32_t m1 = __ballot_sync(0xffffffffu, g == 1);
                        const uint32_t m2 = __ballot_sync(0xffffffffu, g == 2);
                        miss_here |= (col >= 0 && g > 2u);
                        uint32_t mine = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
                        if (w < 3) {
                            if (lane < 3) grp[lane * 12 + q * 3 + w] = mine;
                        } else {                                                   // the block's 4-bit tail
                            mine &= 0xFu;
                            tailacc |= mine << (group_shift(q) + 4 * (tg & 1));    // = tri_tail_shift(4 tg + q)
                        }
                        n += (uint32_t) __popc(mine);
                    }
                    nacc += n << group_shift(q);
                    if (__any_sync(0xffffffffu, miss_here)) missacc |= 0xFFu << group_shift(q);
                }
                const int mq = marg_off + tg * 4;
                if (lane < 3) {
                    if (tailacc) atomicOr(dst + (nblocks >> 2) * 36 + (tg >> 1) * 4 + lane, tailacc);   // shared with the neighbouring group
                    dst[mq + lane] = nacc;
                } else if (lane == 3) {
                    dst[mq + 3] = missacc;
                }
                continue;
            }
            uint32_t nacc = 0, missacc = 0, tailacc = 0, my_ent = 0;     // my_ent: lane e keeps entry e of the group's missing list
            int tail_o = 0, mq = 0, nent = 0;
            for (int q = 0; q < bpt; q++) {
                const int b = tg * bpt + q;
                uint32_t n = 0, inno_any = 0;
                for (int w = 0; w < bw; w++) {
                    const int wb = b * bw + w;
                    const int32_t col = perm_s[wb * 32 + lane];
                    const uint32_t g = col >= 0 ? row[col] : 255u;
                    const uint32_t m0 = __ballot_sync(0xffffffffu, g == 0);
                    const uint32_t m1 = __ballot_sync(0xffffffffu, g == 1);
                    const uint32_t m2 = __ballot_sync(0xffffffffu, g == 2);
                    if (marg) {
                        const uint32_t inno = __ballot_sync(0xffffffffu, col >= 0 && g > 2u);      // samples that are in no plane
                        inno_any |= inno;
                        if (mlist && inno) {
#pragma unroll
                            for (int half = 0; half < 2; half++) {
                                const uint32_t m16 = (inno >> (16 * half)) & 0xffffu;
                                if (!m16) continue;
                                if (lane == nent) my_ent = m16 | ((uint32_t) (16 * half) << 16) | ((uint32_t) (q * 3 * bw + w) << 21) | ((uint32_t) q << 28);
                                nent++;
                            }
                        }
                    }
                    uint32_t mine = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
                    const int2 om = ofs_s[wb];
                    mq = om.y & 0xFFFFFF;
                    if (om.x >= 0) {
                        if (lane < 3) dst[om.x + lane * gstride] = mine;
                    } else {                                                                     // tri tail: four bits of this block
                        mine &= 0xFu;
                        tailacc |= mine << ((-om.x) >> 24);
                        tail_o = ((-om.x) & 0xFFFFFF) - 1;
                    }
                    n += (uint32_t) __popc(mine);
                }
                nacc += n << group_shift(q);                        // bpt == 4 whenever marg is set: q = b & 3
                if (inno_any) missacc |= 0xFFu << group_shift(q);
            }
            // with a list that holds every missing sample of the group no block needs its genotype-2 cells counted directly
            const bool listed = mlist && nent <= kMissListWords;
            if (lane < 3) {
                if (tailacc) atomicOr(dst + tail_o + lane, tailacc);     // the tail word is shared with the neighbouring group
                if (marg) dst[mq + lane] = nacc;
            } else if (lane == 3 && marg) {
                dst[mq + 3] = listed ? 0u : missacc;
            }
            if (mlist && lane < kMissListWords) dst[mq + 4 + lane] = (listed && lane < nent) ? my_ent : 0u;
        }
        __syncthreads();
        // chunk ch of the nr SNPs is one contiguous run of nr * row_words words in global memory
        for (int ch = 0; ch < nchunks; ch++) {
            uint4 *dst = reinterpret_cast<uint4 *>(planes + ((int64_t) ch * snp_pad + snp0) * row_words);
            for (int x = tid; x < nr * (row_words / 4); x += blockDim.x) {
                const int r = x / (row_words / 4), q = x - r * (row_words / 4);
                dst[x] = reinterpret_cast<const uint4 *>(out_s + r * out_words + ch * row_words)[q];
            }
        }
        __syncthreads();
    }
}

// The samples every SNP is missing (in no plane although the bit position holds a sample), as a list of at most mcap - 1
// entries per SNP followed by kMissEnd (see search3v2_kernel for the entry format); rows past the last SNP get empty lists.
// One warp per SNP.  miss == nullptr: only the longest list's length is computed (max_cnt).
static __global__ void miss_list_kernel(const uint8_t *__restrict__ raw, int64_t nv, int64_t nrows, int64_t nsamples, const int32_t *__restrict__ perm,
                                 const FoldLayout *__restrict__ flp, const uint16_t *__restrict__ blk_desc, int64_t npos, int mcap,
                                 uint32_t *__restrict__ miss, int *__restrict__ max_cnt) {
    const FoldLayout &fl = *flp;
    const int lane = threadIdx.x & 31;
    const int64_t snp = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (snp >= nrows) return;
    const int bw = fl.bw, cb = fl.cb;
    int cnt = 0;
    if (snp < nv) {
        for (int64_t p0 = 0; p0 < npos; p0 += 32) {
            const int64_t pos = p0 + lane;
            const int32_t col = pos < npos ? perm[pos] : -1;
            const bool is_missing = col >= 0 && raw[snp * nsamples + col] > 2;
            const unsigned m = __ballot_sync(0xffffffffu, is_missing);
            if (is_missing && miss) {
                const int idx = cnt + __popc(m & ((1u << lane) - 1u));
                if (idx < mcap - 1) {
                    const int b = (int) (pos / (32 * bw)), w = (int) ((pos / 32) % bw), bit = (int) (pos & 31);
                    const int ch = b / cb, off = (b % cb) * 3 * bw + w;
                    uint32_t code;
                    if (fl.single) code = (uint32_t) (b >> 2) | ((uint32_t) (b & 3) << 10);
                    else { const int seg = blk_desc[b] & 0x7fff; code = (uint32_t) (seg >> 1) | ((uint32_t) (seg & 1) << 5); }
                    miss[snp * mcap + idx] = ((uint32_t) bit << 27) | ((uint32_t) off << 16) | ((uint32_t) ch << 12) | code;
                }
            }
            cnt += __popc(m);
        }
    }
    if (miss) {
        for (int x = min(cnt, mcap - 1) + lane; x < mcap; x += 32) miss[snp * mcap + x] = 0xFFFFFFFFu;
    } else if (lane == 0) {
        atomicMax(max_cnt, cnt);
    }
}

// Inverse of the packer for one SNP: byte masks in the reference's layout
// [genotype][S_pad] (model.c:28-74).  One thread per bit position.
static __global__ void unpack_masks_kernel(const uint32_t *__restrict__ planes, int64_t snp, int64_t snp_pad,
                                    const int32_t *__restrict__ perm, const FoldLayout *__restrict__ flp, int64_t npos,
                                    int A, int a_pad, int s_pad, uint8_t *__restrict__ out) {
    const FoldLayout &fl = *flp;
    const int64_t pos = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const int32_t col = perm[pos];
    if (col < 0) return;
    const int b = (int) (pos / (32 * fl.bw)), w = (int) ((pos / 32) % fl.bw), bit = (int) (pos & 31);
    const int dst = col < A ? col : a_pad + (col - A);
    for (int g = 0; g < 3; g++) {
        const uint32_t word = logical_word(planes, fl, snp_pad, b, snp, g, w);
        out[(int64_t) g * s_pad + dst] = ((word >> bit) & 1u) ? 0xFF : 0x00;
    }
}

// ============================================================================
// Shared pieces of the search kernels
// ============================================================================
// ---- 128-bit ranking keys ------------------------------------------------------
// The canonical order (accuracy descending, then SNP tuple ascending) as "larger key first"; tuples are unique within a
// fold, so keys are unique.  Used by the merge and by the global root below.
// (struct Key128: epi_types.h)
__device__ __forceinline__ bool key_ge(const Key128 &a, const Key128 &b) { return a.hi > b.hi || (a.hi == b.hi && a.lo >= b.lo); }
__device__ __forceinline__ bool key_gt(const Key128 &a, const Key128 &b) { return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo); }
__device__ __forceinline__ Key128 make_key(double ba, int i, int j, int k) {      // k < 0: order 2
    Key128 key;
    unsigned long long u = (unsigned long long) __double_as_longlong(ba);
    key.hi = (u >> 63) ? ~u : (u | 0x8000000000000000ULL);           // larger accuracy -> larger key
    unsigned long long t = k < 0 ? (((unsigned long long) (uint32_t) i << 32) | (uint32_t) j)
                                 : (((unsigned long long) (uint32_t) i << 42) | ((unsigned long long) (uint32_t) j << 21) | (uint32_t) k);
    key.lo = ~t;                                                      // smaller tuple -> larger key
    return key;
}
__device__ __forceinline__ Key128 cand_key(const Cand &c, int order) { return make_key(c.ba, c.i, c.j, order == 2 ? -1 : c.k); }

// step sequencing of search2_kernel (used by the producer lane only)
struct ProducerState {
    long long u, pre_u;            // current unit; unit whose descriptor `pre` was loaded one unit ahead
    int2 pre;
    int grp, chunk, cur_i0, cur_j0;
    int redo;                      // 0: main pass, 1: re-running the first unit, 2: no more work
    int started;
};

struct __align__(16) SearchCtl {
    uint64_t full[3];              // per stage (two or three are in use): the chunk rows have landed
    uint64_t empty[3];             // per stage: every warp is done reading it
    int4 meta[4];                  // step descriptors (stages + 1 slots in use): x = chunk (-1: no more work), y/z/w = tile origins
    long long thr[kMaxFolds];      // score a candidate must reach to be offered to the list
    int tq[kMaxFolds];             // the same bound on sum_c max(0, trA - trU) (balanced pre-filter): ceil(thr / n_f)
    int lock[kMaxFolds];
    int cnt[kMaxFolds];
    // the entry that ranks last in a FULL list, published under a sequence lock (odd = being written, 0 = list not full yet):
    // offers that rank at or after it are dropped without taking the fold's lock (floods of candidates that tie with the bound)
    int worst[kMaxFolds];          // array lists (offer_batch_scan): position of the entry that ranks last, once the list is full
    int root_seq[kMaxFolds];
    double root_ba[kMaxFolds];
    int root_t[kMaxFolds][3];
    int first_best[kMaxFolds];     // best pre-filter score of the CTA's first unit, per fold (-1: none); see hist_count_tuple
    int first_done;                // warps that are through with the CTA's first unit
    FoldLayout fl;
    ProducerState ps;              // the tri kernel keeps its step sequencing here instead of in registers (HPGV_PS_IN_SMEM); last
                                   // member on purpose: the other kernels' code (offsets of everything above) stays as it was
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// words between the counter slices of two consecutive threads: cells x counter words, made odd (bank-conflict free)
__host__ __device__ inline int counter_stride(int ncells, int nwc) { return (ncells * nwc) | 1; }

// dynamic shared memory map of the search kernels (the host uses the same function to size the launch)
struct SmemMap {
    size_t stage0, stage_bytes, counters, desc, lists, total;
};
__host__ __device__ inline SmemMap search_smem_map(const FoldLayout &fl, int rows, int ncells, int nthreads, int rank, bool lists_in_smem,
                                                   int nstages = 2) {
    SmemMap m;
    m.stage0 = align_up(sizeof(SearchCtl), 128);
    m.stage_bytes = align_up((size_t) rows * fl.row_words * 4, 128);
    m.counters = m.stage0 + (size_t) nstages * m.stage_bytes;
    const int nwc = fl.single ? fl.nblocks / 4 : fl.F;
    m.desc = m.counters + (size_t) counter_stride(ncells, nwc) * nthreads * 4;
    m.lists = align_up(m.desc + (fl.single ? 0 : (size_t) fl.nblocks * 2), 16);
    m.total = m.lists + (lists_in_smem ? (size_t) fl.F * rank * sizeof(Cand) : 0);
    return m;
}

// smallest t with t * n >= thr: the pre-filter's bound for a fold whose training part holds n cases and n controls
__device__ __forceinline__ int thr_quotient(long long thr, int n) {
    if (thr == LLONG_MIN || n <= 0) return INT_MIN;
    long long q = thr / n;
    if (thr % n > 0) q++;
    return q > INT_MAX ? INT_MAX : (q < INT_MIN ? INT_MIN : (int) q);
}

// ---- top-N list maintenance (one list per CTA and fold) ------------------------
// Replaces add_to_model_ranking (model.c:481-521) with a deterministic order.  A
// list that has reached N entries is kept as a binary heap whose root is the entry
// that ranks LAST in the canonical order, so an offer is one comparison with the
// root and, when it wins, one sift-down.  Lane 0 does the work under the fold's
// lock; the candidate is warp-uniform.
__device__ __forceinline__ void heap_sift_down(Cand *list, int n, int i, const Cand &c) {
    // place c at or below position i: children that rank after c move up
    for (;;) {
        int w = 2 * i + 1;
        if (w >= n) break;
        Cand cw = cand_load(list + w);
        if (w + 1 < n) {
            const Cand cr = cand_load(list + w + 1);
            if (cand_before(cw, cr)) { cw = cr; w = w + 1; }      // the right child ranks after the left one
        }
        if (!cand_before(c, cw)) break;                           // c ranks after both children: it stays here
        cand_store(list + i, cw);
        i = w;
    }
    cand_store(list + i, c);
}

// publish the entry that ranks last in a full list: sequence-locked copy for the lock-free pre-check, score bounds
__device__ __forceinline__ void publish_root(SearchCtl *ctl, const SearchArgs &a, int f, double ba, int i, int j, int k, int tp, int fp) {
    volatile int *seq = &ctl->root_seq[f];
    const int s0 = *seq;
    *seq = s0 + 1;                                   // odd: readers keep out
    __threadfence_block();
    *reinterpret_cast<volatile double *>(&ctl->root_ba[f]) = ba;
    *reinterpret_cast<volatile int *>(&ctl->root_t[f][0]) = i;
    *reinterpret_cast<volatile int *>(&ctl->root_t[f][1]) = j;
    *reinterpret_cast<volatile int *>(&ctl->root_t[f][2]) = k;
    __threadfence_block();
    *seq = s0 + 2;
    const FoldLayout &fl = ctl->fl;
    const int npos = a.training ? fl.A - fl.a_in[f] : fl.a_in[f];
    const int nneg = a.training ? fl.U - fl.u_in[f] : fl.u_in[f];
    const long long sc = (npos == 0 || nneg == 0) ? LLONG_MIN
                         : (a.eval_fn == kEvalBA ? ba_score(tp, fp, npos, nneg) : value_score(evaluate_fn(a.eval_fn, tp, npos - tp, fp, nneg - fp)));
    atomicMax(&ctl->thr[f], sc);
    atomicMax(&ctl->tq[f], thr_quotient(sc, npos));
    atomicMax(a.gthr + f, sc);
}

__device__ __forceinline__ void warp_offer(SearchCtl *ctl, const SearchArgs &a, Cand *lists, int f, const Cand &c, int lane) {
    if (lane == 0) {
        while (atomicCAS(&ctl->lock[f], 0, 1) != 0) __nanosleep(20);
        __threadfence_block();
        Cand *list = lists + (size_t) f * a.rank;
        const int cnt = *reinterpret_cast<volatile int *>(&ctl->cnt[f]);
        bool new_root = false;
        if (cnt < a.rank) {
            cand_store(list + cnt, c);
            *reinterpret_cast<volatile int *>(&ctl->cnt[f]) = cnt + 1;
            if (cnt + 1 == a.rank) {
                for (int i = a.rank / 2 - 1; i >= 0; i--) {       // Floyd heap construction
                    const Cand x = cand_load(list + i);
                    heap_sift_down(list, a.rank, i, x);
                }
                new_root = true;
            }
        } else {
            const Cand root = cand_load(list);
            if (cand_before(c, root)) {
                heap_sift_down(list, a.rank, 0, c);
                new_root = true;
            }
        }
        if (new_root) {
            // the root ranks last: its score is the threshold an offer must reach from now on
            const Cand root = cand_load(list);
            publish_root(ctl, a, f, root.ba, root.i, root.j, root.k, root.tp, root.fp);
        }
        __threadfence_block();
        atomicExch(&ctl->lock[f], 0);
    }
    __syncwarp();
}

// ---- short lists (N <= 64) as plain arrays ----------------------------------------
// A strong single SNP puts thousands of pairs on exactly the same score; the bound cannot rise above it, so all of them
// are offered, a whole warp of them at a time when the unit holds the SNP's own row.  For that case the warp takes the
// fold's lock ONCE for all its candidates and keeps the list as a plain array: an accepted candidate overwrites the entry
// that ranks last and all 32 lanes find the new last entry (two entries per lane, five shuffle rounds) instead of lane 0
// sifting a heap with dependent loads.
struct RootReg {                 // the entry that ranks last, in registers of every lane (warp-uniform)
    double ba;
    int i, j, k, tp, fp, idx;
};
__device__ __forceinline__ bool before4(double ba_a, int ia, int ja, int ka, double ba_b, int ib, int jb, int kb) {
    return cand_before(ba_a, ia, ja, ka, ba_b, ib, jb, kb);
}
__device__ __forceinline__ RootReg scan_last(const Cand *list, int n, int lane) {
    RootReg r;
    r.ba = 0.0; r.i = r.j = r.k = r.tp = r.fp = 0; r.idx = -1;
    for (int e = lane; e < n; e += 32) {
        const Cand x = cand_load(list + e);
        if (r.idx < 0 || before4(r.ba, r.i, r.j, r.k, x.ba, x.i, x.j, x.k)) {
            r.ba = x.ba; r.i = x.i; r.j = x.j; r.k = x.k; r.tp = x.tp; r.fp = x.fp; r.idx = e;
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const double oba = __shfl_xor_sync(0xffffffffu, r.ba, d);
        const int oi = __shfl_xor_sync(0xffffffffu, r.i, d), oj = __shfl_xor_sync(0xffffffffu, r.j, d), ok = __shfl_xor_sync(0xffffffffu, r.k, d);
        const int otp = __shfl_xor_sync(0xffffffffu, r.tp, d), ofp = __shfl_xor_sync(0xffffffffu, r.fp, d), oidx = __shfl_xor_sync(0xffffffffu, r.idx, d);
        // tuples are unique within a fold: the order is strict, both partners keep the same entry
        if (oidx >= 0 && (r.idx < 0 || before4(r.ba, r.i, r.j, r.k, oba, oi, oj, ok))) {
            r.ba = oba; r.i = oi; r.j = oj; r.k = ok; r.tp = otp; r.fp = ofp; r.idx = oidx;
        }
    }
    return r;
}
__device__ __forceinline__ void dbg_add(const SearchArgs &a, int slot, unsigned long long v) {
    if (a.dbg) atomicAdd(a.dbg + slot, v);
}
// development trace of the offered tuples: (i, j), (fold, score, bound); out of line so that it costs the kernels no registers
static __device__ __noinline__ void dbg_trace(unsigned long long *dbg, int si, int sj, int f, long long score, long long thr) {
    const unsigned long long pos = atomicAdd(dbg + kDbgTraceCount, 1ULL);
    if (pos >= kDbgTraceCap) return;
    dbg[kDbgTrace + 2 * pos] = ((unsigned long long) (uint32_t) si << 32) | (uint32_t) sj;
    dbg[kDbgTrace + 2 * pos + 1] = ((unsigned long long) f << 56) | ((unsigned long long) (score & 0xfffffff) << 28) | (unsigned long long) (thr & 0xfffffff);
}
__device__ __forceinline__ void offer_batch_scan(SearchCtl *ctl, const SearchArgs &a, Cand *lists, int f, unsigned want, double my_ba,
                                                 int si, int sj, int sk, uint32_t mask, int tp, int fp, int lane) {
    Cand *list = lists + (size_t) f * a.rank;
    if (lane == 0) {
        unsigned spins = 0;
        while (atomicCAS(&ctl->lock[f], 0, 1) != 0) { __nanosleep(20); spins++; }
        if (spins) dbg_add(a, 4, spins);
    }
    __syncwarp();
    __threadfence_block();
    int cnt = *reinterpret_cast<volatile int *>(&ctl->cnt[f]);
    RootReg root;
    root.ba = 0.0; root.i = root.j = root.k = root.tp = root.fp = 0; root.idx = -1;
    if (cnt >= a.rank) {
        root.idx = *reinterpret_cast<volatile int *>(&ctl->worst[f]);
        const Cand x = cand_load(list + root.idx);
        root.ba = x.ba; root.i = x.i; root.j = x.j; root.k = x.k; root.tp = x.tp; root.fp = x.fp;
    }
    bool changed = false;
    while (want) {                                   // warp-uniform
        const int src = __ffs(want) - 1;
        want &= want - 1;
        Cand c;
        c.ba = __shfl_sync(0xffffffffu, my_ba, src);
        c.i = __shfl_sync(0xffffffffu, si, src);
        c.j = __shfl_sync(0xffffffffu, sj, src);
        c.k = __shfl_sync(0xffffffffu, sk, src);
        c.mask = __shfl_sync(0xffffffffu, mask, src);
        c.tp = __shfl_sync(0xffffffffu, tp, src);
        c.fp = __shfl_sync(0xffffffffu, fp, src);
        int slot = -1;
        if (cnt < a.rank) slot = cnt++;
        else if (before4(c.ba, c.i, c.j, c.k, root.ba, root.i, root.j, root.k)) slot = root.idx;
        if (slot < 0) continue;
        if (lane == 0) { cand_store(list + slot, c); dbg_add(a, 3, 1); }
        if (cnt >= a.rank) {                         // the list is full and has changed: find the entry that ranks last now
            __threadfence_block();
            __syncwarp();
            root = scan_last(list, a.rank, lane);
            __syncwarp();                            // every lane has read the list before lane 0 stores again
            changed = true;
        }
    }
    if (lane == 0) {
        *reinterpret_cast<volatile int *>(&ctl->cnt[f]) = cnt;
        if (changed) {
            *reinterpret_cast<volatile int *>(&ctl->worst[f]) = root.idx;
            publish_root(ctl, a, f, root.ba, root.i, root.j, root.k, root.tp, root.fp);
        }
        __threadfence_block();
        atomicExch(&ctl->lock[f], 0);
    }
    __syncwarp();
}

// offer the tuples of the lanes whose score reaches the fold's running threshold
__device__ __forceinline__ void offer_fold(SearchCtl *ctl, const SearchArgs &a, Cand *lists, int f, bool valid, long long score,
                                           bool degenerate, int npos, int nneg, int si, int sj, int sk, uint32_t mask, int tp, int fp,
                                           int lane) {
    const long long thr = *reinterpret_cast<volatile long long *>(&ctl->thr[f]);
    unsigned want = __ballot_sync(0xffffffffu, valid && score >= thr);
    if (!want) return;
    if (lane == 0) dbg_add(a, 1, __popc(want));
#ifdef HPGV_DEV_TRACE                                  // development builds only (nvcc -DHPGV_DEV_TRACE): costs the kernels registers
    if (a.dbg && ((want >> lane) & 1u)) dbg_trace(a.dbg, si, sj, f, score, thr);
#endif
    double my_ba = 0.0;
    if ((want >> lane) & 1u) {
        if (degenerate) my_ba = -INFINITY;
        else if (a.eval_fn == kEvalBA) my_ba = balanced_accuracy(tp, fp, npos, nneg);
        else {
            my_ba = evaluate_fn(a.eval_fn, tp, npos - tp, fp, nneg - fp);
            my_ba = isnan(my_ba) ? -INFINITY : __dadd_rn(my_ba, 0.0);        // NaN ranks last, like a degenerate fold
        }
    }
    {
        // Lanes whose tuple ranks at or after the last entry of the CTA's full list cannot enter it: they drop out here, in
        // parallel and without the lock (a strong single SNP puts thousands of pairs on exactly the same score).  The
        // snapshot is only used when the sequence number is even and unchanged around the reads; a stale root ranks at or
        // after the current one, so it only lets more lanes through.
        const volatile int *seq = &ctl->root_seq[f];
        const int s0 = *seq;
        if (s0 > 0 && !(s0 & 1)) {
            __threadfence_block();
            const double rb = *reinterpret_cast<const volatile double *>(&ctl->root_ba[f]);
            const int ri = *reinterpret_cast<const volatile int *>(&ctl->root_t[f][0]);
            const int rj = *reinterpret_cast<const volatile int *>(&ctl->root_t[f][1]);
            const int rk = *reinterpret_cast<const volatile int *>(&ctl->root_t[f][2]);
            __threadfence_block();
            if (*seq == s0) {
                const bool in = (want >> lane) & 1u;
                want = __ballot_sync(0xffffffffu, in && cand_before(my_ba, si, sj, sk, rb, ri, rj, rk));
            }
        }
    }
    if (lane == 0 && want) dbg_add(a, 2, __popc(want));
    if (a.list_scan && a.rank <= 64 && a.lists_in_smem) {     // (lists in global memory keep the heap: fewer, dependent accesses)
        if (want) offer_batch_scan(ctl, a, lists, f, want, my_ba, si, sj, sk, mask, tp, fp, lane);
        return;
    }
    while (want) {
        // the threshold rises while the warp works through its lanes: drop the lanes that no longer reach it
        const long long now = *reinterpret_cast<volatile long long *>(&ctl->thr[f]);
        want &= __ballot_sync(0xffffffffu, score >= now);
        if (!want) break;
        const int src = __ffs(want) - 1;
        want &= want - 1;
        Cand c;
        c.i = __shfl_sync(0xffffffffu, si, src);
        c.j = __shfl_sync(0xffffffffu, sj, src);
        c.k = __shfl_sync(0xffffffffu, sk, src);
        c.mask = __shfl_sync(0xffffffffu, mask, src);
        c.tp = __shfl_sync(0xffffffffu, tp, src);
        c.fp = __shfl_sync(0xffffffffu, fp, src);
        c.ba = __shfl_sync(0xffffffffu, my_ba, src);
        warp_offer(ctl, a, lists, f, c, lane);
    }
}

// ---- per-tuple epilogue ---------------------------------------------------------
// cnts: this thread's private counters, word (c, k) at cnts[k * NCELLS + c].
// U8  (single-block segments): word k of cell c = bytes (A_2k, A_2k+1, U_2k, U_2k+1) of folds 2k and 2k+1
// !U8 (multi-block segments) : word f of cell c = A_f | U_f << 16
//
// General version: any class sizes, the exact high-risk rule with its float32 replay.
template <int NCELLS, bool U8>
__device__ __forceinline__ void epilogue_general(SearchCtl *ctl, const SearchArgs &a, Cand *lists, const uint32_t *cnts, int nwc,
                                                 int nthreads, bool valid, int si, int sj, int sk, int lane) {
    const FoldLayout &fl = ctl->fl;
    const RiskParams rp = risk_params(fl);
    const int nfolds = fl.F;
    int totA[NCELLS], totU[NCELLS];
#pragma unroll
    for (int c = 0; c < NCELLS; c++) { totA[c] = 0; totU[c] = 0; }
    for (int k = 0; k < nwc; k++) {
#pragma unroll
        for (int c = 0; c < NCELLS; c++) {
            const uint32_t w = cnts[k * NCELLS + c];
            if constexpr (U8) {
                totA[c] = __dp4a(w, 0x00000101u, (uint32_t) totA[c]);
                totU[c] = __dp4a(w, 0x01010000u, (uint32_t) totU[c]);
            } else {
                totA[c] += (int) (w & 0xffffu);
                totU[c] += (int) (w >> 16);
            }
        }
    }
    for (int f = 0; f < nfolds; f++) {
        int tp = 0, fp = 0;
        uint32_t mask = 0;
        const int k = U8 ? (f >> 1) : f;
#pragma unroll
        for (int c = 0; c < NCELLS; c++) {
            const uint32_t w = cnts[k * NCELLS + c];
            int inA, inU;
            if constexpr (U8) { inA = (int) ((w >> ((f & 1) * 8)) & 0xffu); inU = (int) ((w >> (16 + (f & 1) * 8)) & 0xffu); }
            else { inA = (int) (w & 0xffffu); inU = (int) (w >> 16); }
            const int trA = totA[c] - inA, trU = totU[c] - inU;
            const bool r = high_risk(trA, trU, rp);          // always on the TRAINING table (epistasis.c:34)
            const int ea = a.training ? trA : inA, eu = a.training ? trU : inU;
            tp += r ? ea : 0;
            fp += r ? eu : 0;
            mask |= (r ? 1u : 0u) << c;
        }
        const int npos = a.training ? fl.A - fl.a_in[f] : fl.a_in[f];
        const int nneg = a.training ? fl.U - fl.u_in[f] : fl.u_in[f];
        const bool degenerate = (npos == 0 || nneg == 0);   // BA = 0/0 = NaN in the reference (model.c:473)
        const long long score = degenerate ? LLONG_MIN
                                : (a.eval_fn == kEvalBA ? ba_score(tp, fp, npos, nneg) : value_score(evaluate_fn(a.eval_fn, tp, npos - tp, fp, nneg - fp)));
        offer_fold(ctl, a, lists, f, valid, score, degenerate, npos, nneg, si, sj, sk, mask, tp, fp, lane);
    }
}

// compile-time loop: f(std::integral_constant<int, 0>{}) ... f(std::integral_constant<int, N - 1>{})
template <int N, int I = 0, typename Fn>
__device__ __forceinline__ void static_for(Fn &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
}

// d = c + a.u16[0] * b.s8[0] + a.u16[1] * b.s8[1]
__device__ __forceinline__ int dp2a_lo_us(uint32_t a16x2, uint32_t b8, int c) {
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a16x2), "r"(b8), "r"(c));
    return d;
}

// risky = (trA >= trU) && (pair != 0); when risky: tpfp += ev, mask |= bit  (four instructions, two of them predicated)
template <uint32_t BIT>
__device__ __forceinline__ void risk_accumulate(int d, uint32_t tr, uint32_t ev, uint32_t &tpfp, uint32_t &mask) {
    asm("{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %2, 0;\n"
        "setp.ne.and.u32 p, %3, 0, p;\n"
        "@p add.u32 %0, %0, %4;\n"
        "@p or.b32 %1, %1, %5;\n"
        "}\n"
        : "+r"(tpfp), "+r"(mask)
        : "r"(d), "r"(tr), "r"(ev), "n"(BIT));
}

// one fold of the balanced epilogue; in_of(c) returns the in-fold pair (cases | controls << 16) of cell c
template <int NCELLS, bool TRAINING, typename InOf>
__device__ __forceinline__ void balanced_fold(SearchCtl *ctl, const SearchArgs &a, Cand *lists, const uint32_t (&tot)[NCELLS], int f,
                                              InOf in_of, bool valid, int si, int sj, int sk, int lane) {
    const FoldLayout &fl = ctl->fl;
    uint32_t tpfp = 0, mask = 0;
    auto cell = [&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const uint32_t in = in_of(c);
        const uint32_t tr = tot[c] - in;                     // no borrow: every half of tot >= the half of in
        const int d = dp2a_lo_us(tr, 0x0000FF01u, 0);        // trA - trU
        // d >= 0 and trA == 0 imply trU == 0, so "tr != 0" is the reference's "cell not empty" (0/0 = NaN is low risk)
        risk_accumulate<(1u << c)>(d, tr, TRAINING ? tr : in, tpfp, mask);
    };
    static_for<NCELLS>(cell);
    const int tp = (int) (tpfp & 0xffffu), fp = (int) (tpfp >> 16);
    const int npos = TRAINING ? fl.A - fl.a_in[f] : fl.a_in[f];
    const int nneg = TRAINING ? fl.U - fl.u_in[f] : fl.u_in[f];
    const bool degenerate = (npos == 0 || nneg == 0);
    const long long score = degenerate ? LLONG_MIN
                            : (a.eval_fn == kEvalBA ? ba_score(tp, fp, npos, nneg) : value_score(evaluate_fn(a.eval_fn, tp, npos - tp, fp, nneg - fp)));
    offer_fold(ctl, a, lists, f, valid, score, degenerate, npos, nneg, si, sj, sk, mask, tp, fp, lane);
}

// d = c + sum over the four bytes of a (unsigned) times the bytes of b (signed)
__device__ __forceinline__ int dp4a_us(uint32_t a8x4, uint32_t b8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a8x4), "r"(b8x4), "r"(c));
    return d;
}

// Pre-filter of the balanced TRAINING epilogue when every fold has as many cases as controls (the layout of
// get_k_folds on a balanced data set): the training part of fold f holds n_f cases and n_f controls, a cell is high
// risk iff trA >= trU (and not empty), so
//     score_f = TP * n_f - FP * n_f = n_f * sum_c max(0, trA_c - trU_c) = n_f * sum_c max(0, D_c - d_cf)
// with D_c = totA_c - totU_c and d_cf = inA_cf - inU_cf: one IDP (dot product with +-1 bytes) and one VIMNMX.RELU per
// cell and fold, no unpacking, no risk mask.  Returns whether any fold of this thread's tuple reaches the fold's bound
// ctl->tq; only then is the exact epilogue (masks, TP/FP, BA, list offer) run.  A stale (lower) bound only costs time.
template <int NCELLS, bool U8>
__device__ __forceinline__ bool balanced_prefilter(const SearchCtl *ctl, const uint32_t *cnts, int nwc) {
    int D[NCELLS];
#pragma unroll
    for (int c = 0; c < NCELLS; c++) D[c] = 0;
    for (int k = 0; k < nwc; k++) {
#pragma unroll
        for (int c = 0; c < NCELLS; c++) {
            const uint32_t w = cnts[k * NCELLS + c];
            D[c] = U8 ? dp4a_us(w, 0xFFFF0101u, D[c]) : dp2a_lo_us(w, 0x0000FF01u, D[c]);
        }
    }
    const volatile int *tq = ctl->tq;
    bool pass = false;
    for (int k = 0; k < nwc; k++) {
        if constexpr (U8) {                                   // w = bytes (A_2k, A_2k+1, U_2k, U_2k+1)
            int t0 = 0, t1 = 0;
#pragma unroll
            for (int c = 0; c < NCELLS; c++) {
                const uint32_t w = cnts[k * NCELLS + c];
                t0 += max(dp4a_us(w, 0x000100FFu, D[c]), 0);
                t1 += max(dp4a_us(w, 0x0100FF00u, D[c]), 0);
            }
            pass |= (t0 >= tq[2 * k]) | (t1 >= tq[2 * k + 1]);
        } else {                                              // w = A_k | U_k << 16
            int t = 0;
#pragma unroll
            for (int c = 0; c < NCELLS; c++) {
                const uint32_t w = cnts[k * NCELLS + c];
                t += max(dp2a_lo_us(w, 0x000001FFu, D[c]), 0);
            }
            pass |= t >= tq[k];
        }
    }
    return pass;
}

// mode of a step (from the producer): bit 0 = count scores into the global histogram -- those that reach the bound,
// or every score when bit 1 is clear; bit 1 = candidates are offered to the lists.  See hist_threshold().
constexpr int kModeCount = 1, kModeOffer = 2;

// The pre-filter scores of this thread's tuple once more, this time counted into the global histogram.  Runs for the
// whole first unit of a CTA (all = true) and afterwards only for warps that hold a candidate: kept out of line so that
// it costs the common path no registers.
template <int NCELLS, bool U8>
__device__ __noinline__ void hist_count_tuple(SearchCtl *ctl, int *ghist, int *ghmax, int hist_bins, const uint32_t *cnts, int nwc,
                                              bool valid, bool all) {
    int D[NCELLS];
#pragma unroll
    for (int c = 0; c < NCELLS; c++) D[c] = 0;
    for (int k = 0; k < nwc; k++) {
#pragma unroll
        for (int c = 0; c < NCELLS; c++) {
            const uint32_t w = cnts[k * NCELLS + c];
            D[c] = U8 ? dp4a_us(w, 0xFFFF0101u, D[c]) : dp2a_lo_us(w, 0x0000FF01u, D[c]);
        }
    }
    const volatile int *tq = ctl->tq;
    const int F = ctl->fl.F;
    for (int f = 0; f < F; f++) {
        const int k = U8 ? (f >> 1) : f;
        const uint32_t sel = U8 ? ((f & 1) ? 0x0100FF00u : 0x000100FFu) : 0x000001FFu;
        int t = 0;
#pragma unroll
        for (int c = 0; c < NCELLS; c++) {
            const uint32_t w = cnts[k * NCELLS + c];
            t += max(U8 ? dp4a_us(w, sel, D[c]) : dp2a_lo_us(w, sel, D[c]), 0);
        }
        const bool on = valid && (all || t >= tq[f]);
        const int m = __reduce_max_sync(0xffffffffu, on ? t : -1);
        if (m < 0) continue;
        if (all) {
            // first unit: the CTA keeps its best pair per fold in shared memory; the last warp through counts it into the
            // global histogram (first_unit_done).  The N best of the CTAs' best pairs bound the N-th best of all their pairs
            // well enough for the second units, at 148 atomics per hot address instead of one per warp -- which is what
            // the second units' look-up has to wait for.
            if ((threadIdx.x & 31) == 0) atomicMax(&ctl->first_best[f], m);
            continue;
        }
        if (on) {
            atomicAdd(ghist + (size_t) f * hist_bins + t, 1);
            atomicAdd(ghist + (size_t) F * hist_bins + (size_t) f * hist_coarse_bins(hist_bins) + (t >> 5), 1);
            if (t == m) atomicMax(ghmax + f, m);     // fire and forget; several lanes at most when scores tie
        }
    }
}

// a warp is through with its share of the CTA's first unit (counted only): the last one publishes the CTA's best pairs
static __device__ __noinline__ void first_unit_done(SearchCtl *ctl, int *ghist, int *ghmax, int *gfirst, int hist_bins, int nwarps, int lane) {
    __syncwarp();
    if (lane != 0) return;
    __threadfence_block();
    if (atomicAdd(&ctl->first_done, 1) != nwarps - 1) return;
    __threadfence_block();
    const int F = ctl->fl.F;
    for (int f = 0; f < F; f++) {
        const int t = *reinterpret_cast<volatile int *>(&ctl->first_best[f]);
        if (t < 0) continue;
        atomicAdd(ghist + (size_t) f * hist_bins + t, 1);
        atomicAdd(ghist + (size_t) F * hist_bins + (size_t) f * hist_coarse_bins(hist_bins) + (t >> 5), 1);
        atomicMax(ghmax + f, t);
    }
    __threadfence();
    atomicAdd(gfirst, 1);
}

// Fast version for balanced data sets (A == U <= 65535): every count pair travels as
// one register (cases | controls << 16); r = A/U = 1 makes the float32 rule exact,
// risky <=> trA >= trU and trA > 0 (see high_risk()).
template <int NCELLS, bool U8, bool TRAINING>
__device__ __forceinline__ void epilogue_balanced_t(SearchCtl *ctl, const SearchArgs &a, Cand *lists, const uint32_t *cnts, int nwc,
                                                    int nthreads, bool valid, int si, int sj, int sk, int lane, int mode) {
    const int nfolds = ctl->fl.F;
    if constexpr (TRAINING) {
        if (a.prefilter) {
            if (mode == kModeCount) {                          // the CTA's first unit: counted now, offered at the end
                hist_count_tuple<NCELLS, U8>(ctl, a.ghist, a.ghmax, a.hist_bins, cnts, nwc, valid, true);
                return;
            }
            const bool pass = balanced_prefilter<NCELLS, U8>(ctl, cnts, nwc);
            if (!__any_sync(0xffffffffu, pass && valid)) return;
            if ((threadIdx.x & 31) == 0) dbg_add(a, 0, 1);
            if (mode & kModeCount) hist_count_tuple<NCELLS, U8>(ctl, a.ghist, a.ghmax, a.hist_bins, cnts, nwc, valid, false);
        }
    }
    uint32_t tot[NCELLS];                     // total cases | total controls << 16
    if constexpr (U8) {
        uint32_t tA[NCELLS], tU[NCELLS];
#pragma unroll
        for (int c = 0; c < NCELLS; c++) { tA[c] = 0; tU[c] = 0; }
        for (int k = 0; k < nwc; k++) {
#pragma unroll
            for (int c = 0; c < NCELLS; c++) {
                const uint32_t w = cnts[k * NCELLS + c];
                tA[c] = __dp4a(w, 0x00000101u, tA[c]);
                tU[c] = __dp4a(w, 0x01010000u, tU[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < NCELLS; c++) tot[c] = tA[c] | (tU[c] << 16);
        for (int k = 0; 2 * k < nfolds; k++) {
            uint32_t w[NCELLS];
#pragma unroll
            for (int c = 0; c < NCELLS; c++) w[c] = cnts[k * NCELLS + c];
            balanced_fold<NCELLS, TRAINING>(ctl, a, lists, tot, 2 * k, [&](int c) { return __byte_perm(w[c], 0u, 0x4240); }, valid, si, sj, sk, lane);
            if (2 * k + 1 < nfolds)
                balanced_fold<NCELLS, TRAINING>(ctl, a, lists, tot, 2 * k + 1, [&](int c) { return __byte_perm(w[c], 0u, 0x4341); }, valid, si, sj, sk, lane);
        }
    } else {
#pragma unroll
        for (int c = 0; c < NCELLS; c++) tot[c] = 0;
        for (int k = 0; k < nwc; k++) {
#pragma unroll
            for (int c = 0; c < NCELLS; c++) tot[c] += cnts[k * NCELLS + c];
        }
        for (int f = 0; f < nfolds; f++)
            balanced_fold<NCELLS, TRAINING>(ctl, a, lists, tot, f, [&](int c) { return cnts[f * NCELLS + c]; }, valid, si, sj, sk, lane);
    }
}
template <int NCELLS, bool U8>
__device__ __forceinline__ void epilogue_balanced(SearchCtl *ctl, const SearchArgs &a, Cand *lists, const uint32_t *cnts, int nwc,
                                                  int nthreads, bool valid, int si, int sj, int sk, int lane, int mode) {
    if (a.training) epilogue_balanced_t<NCELLS, U8, true>(ctl, a, lists, cnts, nwc, nthreads, valid, si, sj, sk, lane, mode);
    else epilogue_balanced_t<NCELLS, U8, false>(ctl, a, lists, cnts, nwc, nthreads, valid, si, sj, sk, lane, mode);
}

// group_shift(q) (epi_device.cuh): shift of the byte counter of block q (0..3) of a four-block group: segments
// (2k A, 2k U, 2k+1 A, 2k+1 U) land in bytes (0, 2, 1, 3), i.e. the word reads (A_2k, A_2k+1, U_2k, U_2k+1)

// block Q (0..3) of a four-block group, single-block segments: the nine (27) cell counts go to byte group_shift(Q) of pk[]
// imiss: SNP i's missing mask of the four-block group (warp-uniform); genotype 2 of SNP i is only counted in the blocks it
// marks, the others get it from SNP j's marginals afterwards (derive_row2)
template <int BW, int Q>
__device__ __forceinline__ void single_block2(const uint32_t *irow, const uint32_t *jrow, uint32_t imiss, uint32_t (&pk)[9]) {
    constexpr int SW = slot_words(BW), off = Q * 3 * SW;
    uint32_t pj[3][BW];
#pragma unroll
    for (int g = 0; g < 3; g++) load_plane<BW>(jrow + off + g * SW, pj[g]);
#pragma unroll
    for (int ga = 0; ga < 3; ga++) {
        if (ga == 2 && !(imiss & (0xFFu << group_shift(Q)))) break;
        uint32_t pi[BW];
        load_plane<BW>(irow + off + ga * SW, pi);
#pragma unroll
        for (int gb = 0; gb < 3; gb++) pk[ga * 3 + gb] = cell_count2_acc<BW, (1u << group_shift(Q))>(pi, pj[gb], pk[ga * 3 + gb]);
    }
}
// In a block where SNP i has no missing sample every sample has one of i's three genotypes:
// n(2, gb) = N_gb(j) - n(0, gb) - n(1, gb), byte by byte (the bytes never borrow: N_gb >= n(0, gb) + n(1, gb) in every block)
// corr[gb]: the samples SNP i is missing that have genotype gb of SNP j (they are in no cell), from the group's list
__device__ __forceinline__ void derive_row2(uint32_t imiss, const uint4 nj, uint32_t (&pk)[9], uint32_t c0 = 0, uint32_t c1 = 0, uint32_t c2 = 0) {
    const uint32_t njv[3] = {nj.x, nj.y, nj.z}, corr[3] = {c0, c1, c2};
#pragma unroll
    for (int gb = 0; gb < 3; gb++) {
        const uint32_t derived = njv[gb] - pk[gb] - pk[3 + gb] - corr[gb];
        pk[6 + gb] = (derived & ~imiss) | (pk[6 + gb] & imiss);
    }
}
template <int BW, int Q>
__device__ __forceinline__ void single_block3(const uint32_t *irow, const uint32_t *jrow, const uint32_t *krow, uint32_t (&pk)[27]) {
    constexpr int SW = slot_words(BW), off = Q * 3 * SW;
    uint32_t pl[3][BW];
#pragma unroll
    for (int g = 0; g < 3; g++) load_plane<BW>(krow + off + g * SW, pl[g]);
#pragma unroll
    for (int ga = 0; ga < 3; ga++) {
        uint32_t pi[BW];
        load_plane<BW>(irow + off + ga * SW, pi);
#pragma unroll
        for (int gb = 0; gb < 3; gb++) {
            uint32_t pj[BW];
            load_plane<BW>(jrow + off + gb * SW, pj);
#pragma unroll
            for (int gc = 0; gc < 3; gc++) {
                const int c = ga * 9 + gb * 3 + gc;
                pk[c] = cell_count3_acc<BW, (1u << group_shift(Q))>(pi, pj, pl[gc], pk[c]);
            }
        }
    }
}

// one group (four blocks) of the tri layout: 36 words of the j row stay in registers, the i row is read plane by plane.
// imiss = SNP i's missing mask of the group (0xFF in the byte of a block where i has a sample in no plane, warp-uniform),
// nj = SNP j's own counts (N_0, N_1, N_2) of the group.  Genotype 2 of SNP i is only counted in the blocks imiss marks;
// elsewhere every sample of the block has one of the three genotypes of i, so n(2, gb) = N_gb(j) - n(0, gb) - n(1, gb).
// On entry pk holds the tail counts of the group; the bytes never borrow (N_gb >= n(0, gb) + n(1, gb) in every block).
__device__ __forceinline__ void tri_group2(const uint32_t *ig, const uint32_t *jg, uint32_t imiss, const uint4 nj, uint32_t (&pk)[9]) {
    uint32_t pj[3][12];
#pragma unroll
    for (int g = 0; g < 3; g++) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const uint4 t = *reinterpret_cast<const uint4 *>(jg + g * 12 + 4 * q);
            pj[g][4 * q] = t.x; pj[g][4 * q + 1] = t.y; pj[g][4 * q + 2] = t.z; pj[g][4 * q + 3] = t.w;
        }
    }
#pragma unroll
    for (int ga = 0; ga < 2; ga++) {
        uint32_t pi[12];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const uint4 t = *reinterpret_cast<const uint4 *>(ig + ga * 12 + 4 * q);
            pi[4 * q] = t.x; pi[4 * q + 1] = t.y; pi[4 * q + 2] = t.z; pi[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int gb = 0; gb < 3; gb++) {
            uint32_t acc = pk[ga * 3 + gb];
            acc = tri_count2_acc<(1u << group_shift(0))>(pi + 0, pj[gb] + 0, acc);
            acc = tri_count2_acc<(1u << group_shift(1))>(pi + 3, pj[gb] + 3, acc);
            acc = tri_count2_acc<(1u << group_shift(2))>(pi + 6, pj[gb] + 6, acc);
            acc = tri_count2_acc<(1u << group_shift(3))>(pi + 9, pj[gb] + 9, acc);
            pk[ga * 3 + gb] = acc;
        }
    }
    if (imiss) {                                    // warp-uniform: some block of the group needs genotype 2 counted
        uint32_t pi[12];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const uint4 t = *reinterpret_cast<const uint4 *>(ig + 24 + 4 * q);
            pi[4 * q] = t.x; pi[4 * q + 1] = t.y; pi[4 * q + 2] = t.z; pi[4 * q + 3] = t.w;
        }
        auto block = [&](auto qq) {
            constexpr int q = decltype(qq)::value;
            if (imiss & (0xFFu << group_shift(q))) {
#pragma unroll
                for (int gb = 0; gb < 3; gb++) pk[6 + gb] = tri_count2_acc<(1u << group_shift(q))>(pi + 3 * q, pj[gb] + 3 * q, pk[6 + gb]);
            }
        };
        static_for<4>(block);
    }
    derive_row2(imiss, nj, pk);
}

// The counting phase of a tuple is bound by the XU pipe (POPC), its epilogue by the ALU.  Warps of one SM sub-partition
// that run in lockstep leave each pipe idle half of the time; delaying the upper half of the warps once, by about half
// the counting time of a unit, lets one half count while the other half evaluates.  (The stage hand-off tolerates a
// drift of one step, so the offset persists.)
__device__ __forceinline__ void stagger_late_warps(int warp, int nwarps, int unit_cycles_per_warp, int mode) {
    // warps w, w + 4, w + 8, ... share a sub-partition; k = the warp's index among them, per = how many there are
    const int k = warp >> 2, per = (nwarps + 3) >> 2;
    if (per < 2) return;
    const long long unit = (long long) unit_cycles_per_warp * per;       // counting time of a unit when the XU is saturated
    // mode 1: the upper half of each sub-partition's warps starts half a unit late; mode 2: evenly spread phases
    const long long wait = mode == 2 ? unit * k / per : (k >= (per + 1) / 2 ? unit / 2 : 0);
    const long long t0 = clock64();
    while (clock64() - t0 < wait) {}
}

// once per unit: adopt the best bound any CTA has published for fold f
__device__ __forceinline__ void refresh_threshold(SearchCtl *ctl, const SearchArgs &a, int f) {
    const long long g = __ldcg(a.gthr + f);
    if (g > *reinterpret_cast<volatile long long *>(&ctl->thr[f])) {
        atomicMax(&ctl->thr[f], g);
        if (a.prefilter) atomicMax(&ctl->tq[f], thr_quotient(g, ctl->fl.A - ctl->fl.a_in[f]));
    }
}

// One warp, one fold: the largest score T such that at least N pairs counted in the global histogram reach it.
// Those pairs are distinct members of this search's range, so no pair below T can be among the N best of the fold: T
// becomes the bound of the CTA's pre-filter and list.  Counts are only ever too low (stale reads, pairs below some
// CTA's bound are not counted), which can only lower T.  Two levels: coarse bins of 32 scores are scanned from the
// top, 32 at a time, then the 32 fine bins of the coarse bin that holds the N-th pair.
__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}
static __device__ __noinline__ void hist_threshold(SearchCtl *ctl, const int *ghist, const int *ghmax, long long *gthr, int hist_bins, int rank, int f, int lane) {
    const int hmax = __ldcg(ghmax + f);
    const int cur = *reinterpret_cast<volatile int *>(&ctl->tq[f]);
    if (hmax < 0 || hmax < cur) return;
    const int F = ctl->fl.F, cbins = hist_coarse_bins(hist_bins);
    const int *h = ghist + (size_t) f * hist_bins;
    const int *hc = ghist + (size_t) F * hist_bins + (size_t) f * cbins;
    const int cfloor = max(cur, 0) >> 5;             // below this coarse bin no better bound can come out
    int cum = 0, T = INT_MIN;
    for (int cbase = hmax >> 5; cbase >= cfloor; cbase -= 32) {
        const int cbin = cbase - lane;               // lane 0 reads the highest bin of the window
        const int v = cbin >= 0 ? __ldcg(hc + cbin) : 0;
        const int incl = warp_inclusive_scan(v, lane);
        const unsigned hit = __ballot_sync(0xffffffffu, cum + incl >= rank);
        if (hit) {
            const int l = __ffs(hit) - 1;
            const int cstar = cbase - l;
            const int above = cum + __shfl_sync(0xffffffffu, incl, l) - __shfl_sync(0xffffffffu, v, l);
            const int t = cstar * 32 + 31 - lane;
            const int fv = t < hist_bins ? __ldcg(h + t) : 0;
            const int fincl = warp_inclusive_scan(fv, lane);
            const unsigned fhit = __ballot_sync(0xffffffffu, above + fincl >= rank);
            // the fine counts may lag the coarse one: the bin's lowest score is then the (valid) answer
            T = fhit ? cstar * 32 + 31 - (__ffs(fhit) - 1) : cstar * 32;
            break;
        }
        cum += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0 && T != INT_MIN && T > cur) {
        const long long sc = (long long) T * (ctl->fl.A - ctl->fl.a_in[f]);
        atomicMax(&ctl->tq[f], T);
        atomicMax(&ctl->thr[f], sc);
        atomicMax(gthr + f, sc);                     // every other CTA adopts it at its next unit (refresh_threshold)
    }
}

// common prologue: barriers, control block, counters, block descriptors
template <bool SINGLE>
__device__ __forceinline__ void search_init(SearchCtl *ctl, const SearchArgs &a, uint32_t *cnt_base, size_t cnt_words, uint16_t *desc) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int st = 0; st < 3; st++) {
            mbar_init(&ctl->full[st], 1);
            mbar_init(&ctl->empty[st], blockDim.x >> 5);
        }
        mbar_fence_init();
        ctl->fl = *a.fl;
    }
    if (tid < kMaxFolds) {
        ctl->thr[tid] = LLONG_MIN; ctl->lock[tid] = 0; ctl->cnt[tid] = 0; ctl->root_seq[tid] = 0; ctl->first_best[tid] = -1;
        if (tid == 0) ctl->first_done = 0;
        ctl->tq[tid] = tid < a.fl->F ? INT_MIN : INT_MAX;      // folds past F (odd F, byte-counter pairs) never pass
    }
    for (size_t x = tid; x < cnt_words; x += blockDim.x) cnt_base[x] = 0;   // halves that are never written must read 0
    if constexpr (!SINGLE) {
        const int nb = a.fl->nblocks;
        for (int x = tid; x < nb; x += blockDim.x) desc[x] = a.blk_desc[x];
    }
    __syncthreads();
}

// common tail of the kernel: publish the CTA's lists
__device__ __forceinline__ void search_publish(SearchCtl *ctl, const SearchArgs &a, Cand *lists) {
    __syncthreads();
    const int F = ctl->fl.F;
    if (a.lists_in_smem) {
        Cand *dst = a.lists + (size_t) blockIdx.x * F * a.rank;
        const int4 *s = reinterpret_cast<const int4 *>(lists);
        int4 *d = reinterpret_cast<int4 *>(dst);
        for (int x = threadIdx.x; x < F * a.rank * 2; x += blockDim.x) d[x] = s[x];
    }
    if (threadIdx.x < F) a.list_cnt[(size_t) blockIdx.x * F + threadIdx.x] = ctl->cnt[threadIdx.x];
}

// ============================================================================
// Order 2
// ============================================================================
// unit = (i-tile of TI = nwarps rows, j-tile of 32 rows); warp w <-> i = i0 + w; lane <-> j = j0 + lane
template <int BW, bool SINGLE, bool BALANCED>
__global__ void __launch_bounds__((SINGLE ? kTriWarps : kMaxWarps) * 32, 1) search2_kernel(const SearchArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SearchCtl *ctl = reinterpret_cast<SearchCtl *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x, TI = nthreads >> 5;
    const int NS = a.nstages;                       // stages of the ring: 3 when shared memory allows, else 2
    const SmemMap sm = search_smem_map(*a.fl, TI + kTileJ, 9, nthreads, a.rank, a.lists_in_smem != 0, NS);
    uint32_t *cnt_base = reinterpret_cast<uint32_t *>(smem_raw + sm.counters);
    uint16_t *desc = reinterpret_cast<uint16_t *>(smem_raw + sm.desc);
    Cand *lists = a.lists_in_smem ? reinterpret_cast<Cand *>(smem_raw + sm.lists) : a.lists + (size_t) blockIdx.x * a.fl->F * a.rank;
    search_init<SINGLE>(ctl, a, cnt_base, (sm.desc - sm.counters) / 4, desc);

    const int nblocks = ctl->fl.nblocks, cb = ctl->fl.cb, nchunks = ctl->fl.nchunks, roww = ctl->fl.row_words;
    constexpr int SW = slot_words(BW);
    const int nwc = SINGLE ? nblocks / 4 : ctl->fl.F;
    const uint32_t row_bytes = (uint32_t) roww * 4;

#ifdef HPGV_PS_IN_SMEM
    // ---- step sequencing (producer lane), state in shared memory (SearchCtl::ps) ----
    // This translation unit is compiled with HPGV_PS_IN_SMEM (epi_k_search2_tri.cu: the tri kernel, 96 registers, 20 warps):
    // as locals the state cost every thread a dozen registers in the counting loop, spills included (c2: 5.23 -> 5.05 ms).
    // The other variants have the registers and are 2 % faster with plain locals (c3, measured): they take the #else branch.
    ProducerState &ps = ctl->ps;
    if (tid == nthreads - 32) {
        ps.u = blockIdx.x; ps.pre_u = -1; ps.pre = make_int2(0, 0);
        ps.grp = 0; ps.chunk = 0; ps.cur_i0 = 0; ps.cur_j0 = 0; ps.redo = 0; ps.started = 0;
    }
    auto decode = [&]() {
        if (a.unit_desc) {
            const int2 d = ps.pre_u == ps.u ? ps.pre : __ldg(a.unit_desc + ps.u);
            ps.cur_i0 = d.x; ps.cur_j0 = d.y;
            ps.pre_u = ps.u + gridDim.x;
            if (ps.pre_u < a.num_units) ps.pre = __ldg(a.unit_desc + ps.pre_u);     // lands while the CTA counts
            return;
        }
        while (ps.grp + 1 < a.n_it && a.unit_prefix[ps.grp + 1] <= ps.u) ps.grp++;
        ps.cur_i0 = (a.it0 + ps.grp) * TI;
        ps.cur_j0 = (a.unit_jt0[ps.grp] + (int) (ps.u - a.unit_prefix[ps.grp])) * kTileJ;
    };
    auto mode_of = [&]() { return !a.use_hist || ps.redo ? kModeOffer : (ps.u == (long long) blockIdx.x ? kModeCount : (kModeCount | kModeOffer)); };
    auto next_desc = [&]() {                        // descriptors in step order, one per call; x = -1 once the work is done
        if (!ps.started) {
            ps.started = 1;
            if (ps.u < a.num_units) { decode(); return make_int4(0, ps.cur_i0, ps.cur_j0, mode_of()); }
            ps.redo = 2;
        }
        if (ps.redo == 2) return make_int4(-1, 0, 0, 0);
        if (++ps.chunk == nchunks) {
            ps.chunk = 0;
            if (ps.redo == 1) { ps.redo = 2; return make_int4(-1, 0, 0, 0); }
            ps.u += gridDim.x;
            if (ps.u >= a.num_units) {
                if (!a.use_hist) { ps.redo = 2; return make_int4(-1, 0, 0, 0); }
                ps.redo = 1; ps.u = blockIdx.x; ps.grp = 0;
            }
            decode();
        }
        return make_int4(ps.chunk, ps.cur_i0, ps.cur_j0, mode_of());
    };
#else
    // ---- step sequencing (thread 0): units blockIdx.x, blockIdx.x + gridDim.x, ... each cut into nchunks steps ----
    long long u = blockIdx.x;
    int grp = 0, chunk = 0, cur_i0 = 0, cur_j0 = 0;
    int2 pre = make_int2(0, 0);                     // descriptor of unit pre_u, loaded one unit ahead
    long long pre_u = -1;
    auto decode = [&]() {
        if (a.unit_desc) {
            const int2 d = pre_u == u ? pre : __ldg(a.unit_desc + u);
            cur_i0 = d.x; cur_j0 = d.y;
            pre_u = u + gridDim.x;
            if (pre_u < a.num_units) pre = __ldg(a.unit_desc + pre_u);     // lands while the CTA counts
            return;
        }
        while (grp + 1 < a.n_it && a.unit_prefix[grp + 1] <= u) grp++;
        cur_i0 = (a.it0 + grp) * TI;
        cur_j0 = (a.unit_jt0[grp] + (int) (u - a.unit_prefix[grp])) * kTileJ;
    };
    // With the score histogram the CTA's first unit is only counted (no list has a bound yet: every pair would be
    // offered) and run again, offers only, after the last unit -- by then the bound keeps nearly all of it out.
    int redo = 0;                                   // 0: main pass, 1: re-running the first unit, 2: no more work
    auto mode_of = [&]() { return !a.use_hist || redo ? kModeOffer : (u == (long long) blockIdx.x ? kModeCount : (kModeCount | kModeOffer)); };
    bool started = false;
    auto next_desc = [&]() {                        // descriptors in step order, one per call; x = -1 once the work is done
        if (!started) {
            started = true;
            if (u < a.num_units) { decode(); return make_int4(0, cur_i0, cur_j0, mode_of()); }
            redo = 2;
        }
        if (redo == 2) return make_int4(-1, 0, 0, 0);
        if (++chunk == nchunks) {
            chunk = 0;
            if (redo == 1) { redo = 2; return make_int4(-1, 0, 0, 0); }
            u += gridDim.x;
            if (u >= a.num_units) {
                if (!a.use_hist) { redo = 2; return make_int4(-1, 0, 0, 0); }
                redo = 1; u = blockIdx.x; grp = 0;
            }
            decode();
        }
        return make_int4(chunk, cur_i0, cur_j0, mode_of());
    };
#endif
    auto issue = [&](int st, int ch, int i0, int j0) {
        uint8_t *dst = smem_raw + sm.stage0 + (size_t) st * sm.stage_bytes;
        const char *src = reinterpret_cast<const char *>(a.planes) + (int64_t) ch * a.snp_pad * row_bytes;
        mbar_arrive_expect_tx(&ctl->full[st], (uint32_t) (TI + kTileJ) * row_bytes);
        bulk_g2s(dst, src + (int64_t) i0 * row_bytes, (uint32_t) TI * row_bytes, &ctl->full[st]);
        bulk_g2s(dst + (size_t) TI * row_bytes, src + (int64_t) j0 * row_bytes, (uint32_t) kTileJ * row_bytes, &ctl->full[st]);
    };
    // the producer is lane 0 of the LAST warp: at the top of its own step s it issues the copies of step s + NS - 1 into
    // the stage step s - 1 has used, once every warp has released it; a warp may thus run up to NS - 1 steps ahead of the
    // slowest one before it has to wait
    const bool producer = (tid == nthreads - 32);
    if (producer) {
        for (int t = 0; t < NS - 1; t++) {
            const int4 d = next_desc();
            ctl->meta[t] = d;
            if (d.x >= 0) issue(t, d.x, d.y, d.z);
            else mbar_arrive(&ctl->full[t]);                                   // end marker: a phase without data
        }
    }

    if (a.dbg && tid == 0) a.dbg[8 + 2 * blockIdx.x] = (unsigned long long) clock64();
    constexpr int NC = 9;
    uint32_t *cnts = cnt_base + (size_t) tid * counter_stride(NC, nwc);          // + k * NC + c
    uint32_t acc[9];
#pragma unroll
    for (int c = 0; c < 9; c++) acc[c] = 0;

    int st = 0, slot = 0;                           // s % NS, s % (NS + 1)
    uint32_t ph = 0;                                // (s / NS) & 1
    for (uint32_t s = 0;; s++) {
        if (producer) {
            const int4 next = next_desc();          // step s + NS - 1
            int tslot = slot + NS - 1;
            if (tslot > NS) tslot -= NS + 1;
            ctl->meta[tslot] = next;
            const int tst = st == 0 ? NS - 1 : st - 1;                         // its stage: the one of step s - 1
            if (s >= 1) mbar_wait(&ctl->empty[tst], st == 0 ? ph ^ 1u : ph);   // every warp has read step s - 1
            if (next.x >= 0) issue(tst, next.x, next.y, next.z);
            else mbar_arrive(&ctl->full[tst]);
        }
        __syncwarp();
        mbar_wait(&ctl->full[st], ph);
        const int4 meta = ctl->meta[slot];
        if (meta.x < 0) break;
        const int ch = meta.x;
        if (s == 0 && a.stagger) stagger_late_warps(warp, TI, nblocks * 9 * (BW == 3 ? 18 : (BW == 4 ? 24 : 34)), a.stagger);
        if (ch == 0) {
            // The best bound ANY CTA has derived (hist_threshold publishes its T, the lists publish their last entry): one L2
            // read per fold and unit, by a different warp every unit so that no warp is always the one that pays the round
            // trip.  Without it a CTA only learns about a strong SNP's pairs from its own share of them and from its own
            // histogram look-ups (every 8th / 32nd unit), and its lists take in thousands of candidates that a fresher
            // bound would have dropped (the slow rank of the 4- and 8-GPU runs of round 1).
            if ((a.fresh_bound || !a.use_hist) && lane < ctl->fl.F && warp == (int) ((s / (uint32_t) nchunks) % (uint32_t) TI))
                refresh_threshold(ctl, a, lane);
            if (a.use_hist) {
                // every unit at first, then ever more rarely: the bound rises with the logarithm of the pairs seen, and
                // the warps that pay the L2 round trips are late at the next hand-off of a stage
                // (staggered over the CTAs: somebody looks, and publishes, every unit)
                const uint32_t n = s / (uint32_t) nchunks, m = n + blockIdx.x;
                // The CTA's second unit is the first one whose pairs are offered, bounded by what ALL CTAs counted in their
                // first units.  The CTAs run in lockstep, so without a wait this look-up races with the other CTAs' counting:
                // when it finds fewer than N pairs the whole unit (640 pairs x F folds) is offered without a bound, and the
                // lists take a six-figure number of insertions before the next look-up (the slow rank of round 1's 4- and
                // 8-GPU runs).  All CTAs are co-resident (grid <= SMs, one CTA per SM); the wait is bounded all the same.
                if (n == 1 && a.first_wait) {         // (every warp: the ones that do not look up must not run ahead to their epilogue)
                    for (int spin = 0; spin < 20000 && *reinterpret_cast<volatile int *>(a.gfirst) < (int) gridDim.x; spin++) __nanosleep(100);
                    __threadfence();
                }
                if (n >= 1 && (n < 8 || (n < 64 ? (m & 7) == 0 : (m & 31) == 0)))
                    for (int f = warp; f < ctl->fl.F; f += TI) hist_threshold(ctl, a.ghist, a.ghmax, a.gthr, a.hist_bins, a.rank, f, lane);
            }
        }

        const uint32_t *sbase = reinterpret_cast<const uint32_t *>(smem_raw + sm.stage0 + (size_t) st * sm.stage_bytes);
        const uint32_t *irow = sbase + (size_t) warp * roww;
        const uint32_t *jrow = sbase + (size_t) (TI + lane) * roww;
        const int b_lo = ch * cb, b_hi = min(nblocks, b_lo + cb);

        if constexpr (BW == 3) {
            // tri layout (one chunk): two groups of four blocks share a tail word
            const int ngroups = nblocks >> 2, ntail = (nblocks + 7) >> 3;
            const uint32_t *itail = irow + ngroups * 36, *jtail = jrow + ngroups * 36;
            for (int m = 0; 2 * m < ngroups; m++) {
                uint32_t pk0[9], pk1[9];
#pragma unroll
                for (int c = 0; c < 9; c++) { pk0[c] = 0; pk1[c] = 0; }
                {
                    const uint4 tj = *reinterpret_cast<const uint4 *>(jtail + 4 * m);
                    const uint4 ti = *reinterpret_cast<const uint4 *>(itail + 4 * m);
                    const uint32_t tjv[3] = {tj.x, tj.y, tj.z}, tiv[3] = {ti.x, ti.y, ti.z};
#pragma unroll
                    for (int ga = 0; ga < 3; ga++)
#pragma unroll
                        for (int gb = 0; gb < 3; gb++) tail_count_acc(tiv[ga] & tjv[gb], pk0[ga * 3 + gb], pk1[ga * 3 + gb]);
                }
                const uint32_t *imarg = irow + ngroups * 36 + ntail * 4, *jmarg = jrow + ngroups * 36 + ntail * 4;
                const uint32_t derive_off = a.tri_derive ? 0u : 0xFFFFFFFFu;     // A/B switch: count every block directly
                tri_group2(irow + (2 * m) * 36, jrow + (2 * m) * 36, imarg[4 * (2 * m) + 3] | derive_off,
                           *reinterpret_cast<const uint4 *>(jmarg + 4 * (2 * m)), pk0);
#pragma unroll
                for (int c = 0; c < 9; c++) cnts[(2 * m) * 9 + c] = pk0[c];
                if (2 * m + 1 < ngroups) {
                    tri_group2(irow + (2 * m + 1) * 36, jrow + (2 * m + 1) * 36, imarg[4 * (2 * m + 1) + 3] | derive_off,
                               *reinterpret_cast<const uint4 *>(jmarg + 4 * (2 * m + 1)), pk1);
#pragma unroll
                    for (int c = 0; c < 9; c++) cnts[(2 * m + 1) * 9 + c] = pk1[c];
                }
            }
        } else if constexpr (SINGLE) {
            for (int b4 = b_lo; b4 < b_hi; b4 += 4) {
                uint32_t pk[9];
#pragma unroll
                for (int c = 0; c < 9; c++) pk[c] = 0;
                const int off = (b4 - b_lo) * 3 * SW;
                uint32_t imiss = 0xFFFFFFFFu;                                   // without marginals every block is counted
                uint4 nj = make_uint4(0u, 0u, 0u, 0u);
                uint32_t c0 = 0, c1 = 0, c2 = 0;
                if (ctl->fl.marg && a.tri_derive) {
                    const int mq = ctl->fl.marg_off + ((b4 - b_lo) >> 2) * ctl->fl.marg_stride;   // one quad (+ list) per four blocks
                    imiss = irow[mq + 3];
                    nj = *reinterpret_cast<const uint4 *>(jrow + mq);
                    if (ctl->fl.mlist) {
                        // SNP i's missing samples of this group, 16 bit positions per entry (warp-uniform): which genotype of SNP j
                        // do they have?  Those samples are in no cell, and the derivation below would put them into row 2.
                        const uint32_t *il = irow + mq + 4;
#pragma unroll 1
                        for (int e = 0; e < kMissListWords; e++) {
                            const uint32_t ent = il[e];
                            if (ent == 0) break;
                            const uint32_t m = (ent & 0xffffu) << ((ent >> 16) & 31u);
                            const uint32_t *jp = jrow + off + ((ent >> 21) & 127u);
                            const uint32_t unit = 1u << group_shift_rt(ent >> 28);
                            c0 += (uint32_t) __popc(jp[0] & m) * unit;
                            c1 += (uint32_t) __popc(jp[SW] & m) * unit;
                            c2 += (uint32_t) __popc(jp[2 * SW] & m) * unit;
                        }
                    }
                }
                single_block2<BW, 0>(irow + off, jrow + off, imiss, pk);
                single_block2<BW, 1>(irow + off, jrow + off, imiss, pk);
                single_block2<BW, 2>(irow + off, jrow + off, imiss, pk);
                single_block2<BW, 3>(irow + off, jrow + off, imiss, pk);
                derive_row2(imiss, nj, pk, c0, c1, c2);
                const int k = b4 >> 2;
#pragma unroll
                for (int c = 0; c < 9; c++) cnts[k * 9 + c] = pk[c];
            }
        } else {
            for (int b = b_lo; b < b_hi; b++) {
                const int off = (b - b_lo) * 3 * SW;
                uint32_t pj[3][BW];
#pragma unroll
                for (int g = 0; g < 3; g++) load_plane<BW>(jrow + off + g * SW, pj[g]);
#pragma unroll
                for (int ga = 0; ga < 3; ga++) {
                    uint32_t pi[BW];
                    load_plane<BW>(irow + off + ga * SW, pi);
#pragma unroll
                    for (int gb = 0; gb < 3; gb++) acc[ga * 3 + gb] = cell_count2_acc<BW, 1u>(pi, pj[gb], acc[ga * 3 + gb]);
                }
                const unsigned d = desc[b];
                if (d & 0x8000u) {
                    const int seg = d & 0x7fff;
#pragma unroll
                    for (int c = 0; c < 9; c++) {
                        reinterpret_cast<uint16_t *>(cnts + (seg >> 1) * NC + c)[seg & 1] = (uint16_t) acc[c];
                        acc[c] = 0;
                    }
                }
            }
        }

        __syncwarp();
        // the step descriptor is read again here (its slot is rewritten two steps later, which needs this warp's
        // arrival below) so that the tile origins hold no registers during the counting phase
        const int4 meta2 = ld_volatile_shared_int4(&ctl->meta[slot]);
        if (lane == 0) mbar_arrive(&ctl->empty[st]);        // this warp is done with the stage

        if (meta2.x == nchunks - 1) {
            const int i0 = meta2.y, j0 = meta2.z, mode = meta2.w;
            const int i = i0 + warp, j = j0 + lane;
            bool valid = (i < j) && (j < a.nv);
            if (valid && (i <= a.edge_lo || i >= a.edge_hi)) {
                const uint64_t idx = pair_index((uint64_t) a.nv, (uint64_t) i, (uint64_t) j);
                valid = idx >= a.first && idx < a.last;
            }
            if (__any_sync(0xffffffffu, valid)) {
                if constexpr (BALANCED) epilogue_balanced<9, SINGLE>(ctl, a, lists, cnts, nwc, nthreads, valid, i, j, -1, lane, mode);
                else epilogue_general<9, SINGLE>(ctl, a, lists, cnts, nwc, nthreads, valid, i, j, -1, lane);
            }
            if (mode == kModeCount) first_unit_done(ctl, a.ghist, a.ghmax, a.gfirst, a.hist_bins, TI, lane);
        }
        if (++st == NS) { st = 0; ph ^= 1u; }
        if (++slot > NS) slot = 0;
    }
    if (a.dbg && tid == 0) a.dbg[9 + 2 * blockIdx.x] = (unsigned long long) clock64();
    search_publish(ctl, a, lists);
}

// ============================================================================
// Order 3
// ============================================================================
// unit = (i, j-tile of TJ = nwarps rows); the CTA walks the k-tiles of the unit itself.
// warp w <-> j = j0 + w; lane <-> k = k0 + lane; one triple per thread and (unit, k-tile).
template <int BW, bool SINGLE, bool BALANCED>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) search3_kernel(const SearchArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SearchCtl *ctl = reinterpret_cast<SearchCtl *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x, TJ = nthreads >> 5;
    const SmemMap sm = search_smem_map(*a.fl, 1 + TJ + kTileJ, 27, nthreads, a.rank, a.lists_in_smem != 0);
    uint32_t *cnt_base = reinterpret_cast<uint32_t *>(smem_raw + sm.counters);
    uint16_t *desc = reinterpret_cast<uint16_t *>(smem_raw + sm.desc);
    Cand *lists = a.lists_in_smem ? reinterpret_cast<Cand *>(smem_raw + sm.lists) : a.lists + (size_t) blockIdx.x * a.fl->F * a.rank;
    search_init<SINGLE>(ctl, a, cnt_base, (sm.desc - sm.counters) / 4, desc);

    const int nblocks = ctl->fl.nblocks, cb = ctl->fl.cb, nchunks = ctl->fl.nchunks, roww = ctl->fl.row_words;
    constexpr int SW = slot_words(BW);
    const int nwc = SINGLE ? nblocks / 4 : ctl->fl.F;
    const uint32_t row_bytes = (uint32_t) roww * 4;
    const int nkt = (a.nv + kTileJ - 1) / kTileJ;

    // ---- step sequencing (thread 0): unit -> k-tiles -> chunks ----
    long long u = blockIdx.x;
    int grp = 0, chunk = 0, cur_i = 0, cur_j0 = 0, cur_kt = 0;
    auto decode = [&]() {
        while (grp + 1 < a.n_it && a.unit_prefix[grp + 1] <= u) grp++;
        cur_i = a.it0 + grp;
        cur_j0 = (a.unit_jt0[grp] + (int) (u - a.unit_prefix[grp])) * TJ;
        cur_kt = (cur_j0 + 1) / kTileJ;        // first k-tile that can hold k > j0
    };
    auto issue = [&](int st, int ch, int i, int j0, int k0) {
        uint8_t *dst = smem_raw + sm.stage0 + (size_t) st * sm.stage_bytes;
        const char *src = reinterpret_cast<const char *>(a.planes) + (int64_t) ch * a.snp_pad * row_bytes;
        mbar_arrive_expect_tx(&ctl->full[st], (uint32_t) (1 + TJ + kTileJ) * row_bytes);
        bulk_g2s(dst, src + (int64_t) i * row_bytes, row_bytes, &ctl->full[st]);
        bulk_g2s(dst + row_bytes, src + (int64_t) j0 * row_bytes, (uint32_t) TJ * row_bytes, &ctl->full[st]);
        bulk_g2s(dst + (size_t) (1 + TJ) * row_bytes, src + (int64_t) k0 * row_bytes, (uint32_t) kTileJ * row_bytes, &ctl->full[st]);
    };
    // meta: x = chunk, y = i, z = j0, w = k0
    const bool producer = (tid == nthreads - 32);
    if (producer) {
        if (u < a.num_units) {
            decode();
            ctl->meta[0] = make_int4(0, cur_i, cur_j0, cur_kt * kTileJ);
            issue(0, 0, cur_i, cur_j0, cur_kt * kTileJ);
        } else {
            ctl->meta[0] = make_int4(-1, 0, 0, 0);
            mbar_arrive(&ctl->full[0]);
        }
    }

    constexpr int NC = 27;
    uint32_t *cnts = cnt_base + (size_t) tid * counter_stride(NC, nwc);
    uint32_t acc[27];
#pragma unroll
    for (int c = 0; c < 27; c++) acc[c] = 0;

    for (uint32_t s = 0;; s++) {
        const int st = s & 1;
        if (producer) {
            int4 next = make_int4(-1, 0, 0, 0);
            if (u < a.num_units) {
                if (++chunk == nchunks) {
                    chunk = 0;
                    if (++cur_kt == nkt) {
                        u += gridDim.x;
                        if (u < a.num_units) decode();
                    }
                }
                if (u < a.num_units) next = make_int4(chunk, cur_i, cur_j0, cur_kt * kTileJ);
            }
            ctl->meta[(s + 1) % 3] = next;
            if (s >= 1) mbar_wait(&ctl->empty[st ^ 1], ((s - 1) >> 1) & 1);
            if (next.x >= 0) issue(st ^ 1, next.x, next.y, next.z, next.w);
            else mbar_arrive(&ctl->full[st ^ 1]);
        }
        __syncwarp();
        mbar_wait(&ctl->full[st], (s >> 1) & 1);
        const int4 meta = ctl->meta[s % 3];
        if (meta.x < 0) break;
        const int ch = meta.x;
        if (s == 0 && a.stagger) stagger_late_warps(warp, TJ, nblocks * 27 * (BW == 4 ? 24 : 34), a.stagger);
        if (ch == 0 && warp == 0 && lane < ctl->fl.F) refresh_threshold(ctl, a, lane);

        const uint32_t *sbase = reinterpret_cast<const uint32_t *>(smem_raw + sm.stage0 + (size_t) st * sm.stage_bytes);
        const uint32_t *irow = sbase;
        const uint32_t *jrow = sbase + (size_t) (1 + warp) * roww;
        const uint32_t *krow = sbase + (size_t) (1 + TJ + lane) * roww;
        const int b_lo = ch * cb, b_hi = min(nblocks, b_lo + cb);

        if constexpr (SINGLE) {
            for (int b4 = b_lo; b4 < b_hi; b4 += 4) {
                uint32_t pk[27];
#pragma unroll
                for (int c = 0; c < 27; c++) pk[c] = 0;
                const int off = (b4 - b_lo) * 3 * SW;
                single_block3<BW, 0>(irow + off, jrow + off, krow + off, pk);
                single_block3<BW, 1>(irow + off, jrow + off, krow + off, pk);
                single_block3<BW, 2>(irow + off, jrow + off, krow + off, pk);
                single_block3<BW, 3>(irow + off, jrow + off, krow + off, pk);
                const int k = b4 >> 2;
#pragma unroll
                for (int c = 0; c < 27; c++) cnts[k * 27 + c] = pk[c];
            }
        } else {
            for (int b = b_lo; b < b_hi; b++) {
                const int off = (b - b_lo) * 3 * SW;
                uint32_t pl[3][BW];
#pragma unroll
                for (int g = 0; g < 3; g++) load_plane<BW>(krow + off + g * SW, pl[g]);
#pragma unroll
                for (int ga = 0; ga < 3; ga++) {
                    uint32_t pi[BW];
                    load_plane<BW>(irow + off + ga * SW, pi);
#pragma unroll
                    for (int gb = 0; gb < 3; gb++) {
                        uint32_t pj[BW];
                        load_plane<BW>(jrow + off + gb * SW, pj);
#pragma unroll
                        for (int gc = 0; gc < 3; gc++) {
                            const int c = ga * 9 + gb * 3 + gc;
                            acc[c] = cell_count3_acc<BW, 1u>(pi, pj, pl[gc], acc[c]);
                        }
                    }
                }
                const unsigned d = desc[b];
                if (d & 0x8000u) {
                    const int seg = d & 0x7fff;
#pragma unroll
                    for (int c = 0; c < 27; c++) {
                        reinterpret_cast<uint16_t *>(cnts + (seg >> 1) * NC + c)[seg & 1] = (uint16_t) acc[c];
                        acc[c] = 0;
                    }
                }
            }
        }

        __syncwarp();
        const int4 meta2 = ld_volatile_shared_int4(&ctl->meta[s % 3]);      // see search2_kernel
        if (lane == 0) mbar_arrive(&ctl->empty[st]);

        if (meta2.x == nchunks - 1) {
            const int i = meta2.y, j0 = meta2.z, k0 = meta2.w;
            const int j = j0 + warp, k = k0 + lane;
            bool valid = (i < j) && (j < k) && (k < a.nv);
            if (valid && (i <= a.edge_lo || i >= a.edge_hi)) {
                const uint64_t idx = triple_index((uint64_t) a.nv, (uint64_t) i, (uint64_t) j, (uint64_t) k);
                valid = idx >= a.first && idx < a.last;
            }
            if (__any_sync(0xffffffffu, valid)) {
                if constexpr (BALANCED) epilogue_balanced<27, SINGLE>(ctl, a, lists, cnts, nwc, nthreads, valid, i, j, k, lane, kModeOffer);
                else epilogue_general<27, SINGLE>(ctl, a, lists, cnts, nwc, nthreads, valid, i, j, k, lane);
            }
        }
    }
    search_publish(ctl, a, lists);
}

// ============================================================================
// Order 3, resident (j, k) tiles and streamed i rows
// ============================================================================
// unit = (tile of TJ j rows, tile of 32 k rows); the rows of the tile -- every chunk of them -- stay in shared memory while
// the i rows (i < j) stream through a ring of stages, NI of them per step.  Two warps share a j: warp (ga, jw), ga in {0, 1},
// lane <-> k, so a thread counts the NINE cells (ga, gb, gc) of one triple and the pair of threads (0, jw, lane), (1, jw, lane)
// covers the 18 cells of i's genotypes 0 and 1.  The cells of genotype 2 of SNP i are not counted:
//     n(2, gb, gc) = n_jk(gb, gc) - n(0, gb, gc) - n(1, gb, gc) - n(i missing, gb, gc)
// with n_jk the pair table of (j, k), counted once per unit (it is the same for every i).  The last term is sparse: the
// packer lists the handful of samples every SNP is missing (miss_list_kernel); before the pair of threads evaluates a triple
// the ga = 0 thread takes SNP i's missing samples out of the pair table, one shared-memory atomic each, and the ga = 1 thread
// puts them back afterwards.  A third of the AND / carry-save / POPC work of the plain kernel is gone, 27 x F counter words
// per triple become 9 x F per thread (twice the warps fit an SM), and the j and k planes are read from HBM/L2 once per unit
// instead of once per i.  After the counts the two threads meet at a named barrier (64 threads), split the folds between
// them and evaluate them from the three tables; a second barrier keeps the next i's counts off tables that are still read.
struct Smem3Map {
    size_t tile, ring0, stage_bytes, stage_rows_bytes, counters, njk, desc, lists, total;
};
__host__ __device__ inline Smem3Map search3v2_smem_map(const FoldLayout &fl, int tj, int ni, int nstages, int mcap, int rank, bool lists_in_smem) {
    Smem3Map m;
    const size_t snp_bytes = (size_t) fl.nchunks * fl.row_words * 4;              // every chunk of one SNP
    m.tile = align_up(sizeof(SearchCtl), 128);
    m.ring0 = m.tile + align_up((size_t) (tj + kTileJ) * snp_bytes, 128);
    m.stage_rows_bytes = (size_t) ni * snp_bytes;
    m.stage_bytes = align_up(m.stage_rows_bytes + (size_t) ni * mcap * 4, 128);
    m.counters = m.ring0 + (size_t) nstages * m.stage_bytes;
    const int nwc = fl.single ? fl.nblocks / 4 : fl.F;
    const size_t slice = (size_t) counter_stride(9, nwc) * 4;
    m.njk = m.counters + slice * (size_t) (2 * tj * 32);
    m.desc = m.njk + slice * (size_t) (tj * 32);
    m.lists = align_up(m.desc + (fl.single ? 0 : (size_t) fl.nblocks * 2), 16);
    m.total = m.lists + (lists_in_smem ? (size_t) fl.F * rank * sizeof(Cand) : 0);
    return m;
}

// entry of a SNP's list of missing samples (made by miss_list_kernel): where the sample sits in a chunk row and which
// counter it belongs to.  0xFFFFFFFF ends the list.
//   [31:27] bit   [26:16] word offset inside the chunk row: (block in chunk * 3) * slot words + word   [15:12] chunk
//   [11:0]  counter: 16-bit counters: fold | class << 5;  byte counters: word (block / 4) | (block & 3) << 10
constexpr uint32_t kMissEnd = 0xFFFFFFFFu;

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// in-fold count pair (cases | controls << 16) of fold f from a counter word
template <bool U8>
__device__ __forceinline__ uint32_t fold_pair(uint32_t w, int f) {
    if constexpr (U8) return __byte_perm(w, 0u, (f & 1) ? 0x4341 : 0x4240);
    else return w;
}

template <int BW, bool SINGLE, bool BALANCED>
__global__ void __launch_bounds__(kMaxWarps * 32, 1) search3v2_kernel(const SearchArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SearchCtl *ctl = reinterpret_cast<SearchCtl *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5, TJ = nwarps >> 1;
    const int ga = warp / TJ, jw = warp - ga * TJ;               // warps [0, TJ): genotype 0 of SNP i, [TJ, 2 TJ): genotype 1'
    const int NI = a.v2_ni, NS = a.nstages;
    const Smem3Map sm = search3v2_smem_map(*a.fl, TJ, NI, NS, a.v2_mcap, a.rank, a.lists_in_smem != 0);
    uint32_t *cnt_base = reinterpret_cast<uint32_t *>(smem_raw + sm.counters);
    uint16_t *desc = reinterpret_cast<uint16_t *>(smem_raw + sm.desc);
    Cand *lists = a.lists_in_smem ? reinterpret_cast<Cand *>(smem_raw + sm.lists) : a.lists + (size_t) blockIdx.x * a.fl->F * a.rank;
    search_init<SINGLE>(ctl, a, cnt_base, (sm.desc - sm.counters) / 4, desc);      // (barriers: full/empty[0..1] = the ring, full[2] = the tile)
    constexpr int SW = slot_words(BW);
    const int nblocks = ctl->fl.nblocks, cb = ctl->fl.cb, nchunks = ctl->fl.nchunks, roww = ctl->fl.row_words, F = ctl->fl.F;
    const int nwc = SINGLE ? nblocks / 4 : F;
    const int stride = counter_stride(9, nwc);
    const uint32_t row_bytes = (uint32_t) roww * 4;
    const int tile_rows = TJ + kTileJ;

    // ---- step sequencing (lane 0 of the last warp): unit -> groups of NI i rows ----
    // meta: x = first i of the group (-1: no more work), y = j0, z = k0, w = number of i rows of this group | new unit << 16
    long long u = blockIdx.x;
    int cur_j0 = 0, cur_k0 = 0, cur_i = 0, cur_iend = 0;
    bool fresh = false;
    auto decode = [&]() {
        const int2 d = __ldg(a.unit_desc + u);
        cur_j0 = d.x; cur_k0 = d.y;
        cur_i = a.edge_lo;                                         // first SNP of the range's first triple
        cur_iend = min(a.edge_hi, min(cur_j0 + TJ - 2, a.nv - 3)) + 1;   // i < j <= j0 + TJ - 1, and inside the range
        fresh = true;
    };
    auto next_desc = [&]() {
        for (;;) {
            if (u >= a.num_units) return make_int4(-1, 0, 0, 0);
            if (cur_i < cur_iend) {
                const int n = min(NI, cur_iend - cur_i);
                const int4 d = make_int4(cur_i, cur_j0, cur_k0, n | (fresh ? 1 << 16 : 0));
                cur_i += n; fresh = false;
                return d;
            }
            u += gridDim.x;
            if (u < a.num_units) decode();
        }
    };
    auto issue_tile = [&](int j0, int k0) {
        uint8_t *dst = smem_raw + sm.tile;
        mbar_arrive_expect_tx(&ctl->full[2], (uint32_t) nchunks * tile_rows * row_bytes);
        for (int ch = 0; ch < nchunks; ch++) {
            const char *src = reinterpret_cast<const char *>(a.planes) + (int64_t) ch * a.snp_pad * row_bytes;
            bulk_g2s(dst + (size_t) ch * tile_rows * row_bytes, src + (int64_t) j0 * row_bytes, (uint32_t) TJ * row_bytes, &ctl->full[2]);
            bulk_g2s(dst + ((size_t) ch * tile_rows + TJ) * row_bytes, src + (int64_t) k0 * row_bytes, (uint32_t) kTileJ * row_bytes, &ctl->full[2]);
        }
    };
    auto issue_rows = [&](int st, int i0) {                         // always NI rows: the planes are padded past the last SNP
        uint8_t *dst = smem_raw + sm.ring0 + (size_t) st * sm.stage_bytes;
        mbar_arrive_expect_tx(&ctl->full[st], (uint32_t) (nchunks * NI) * row_bytes + (uint32_t) (NI * a.v2_mcap * 4));
        for (int ch = 0; ch < nchunks; ch++) {
            const char *src = reinterpret_cast<const char *>(a.planes) + (int64_t) ch * a.snp_pad * row_bytes;
            bulk_g2s(dst + (size_t) ch * NI * row_bytes, src + (int64_t) i0 * row_bytes, (uint32_t) NI * row_bytes, &ctl->full[st]);
        }
        bulk_g2s(dst + sm.stage_rows_bytes, a.v2_miss + (int64_t) i0 * a.v2_mcap, (uint32_t) (NI * a.v2_mcap * 4), &ctl->full[st]);
    };
    const bool producer = (tid == blockDim.x - 32);
    if (producer) {
        if (u < a.num_units) decode();
        const int4 d = next_desc();
        ctl->meta[0] = d;
        if (d.x >= 0) { issue_tile(d.y, d.z); issue_rows(0, d.x); }
        else mbar_arrive(&ctl->full[0]);
    }

    uint32_t *mine = cnt_base + (size_t) ((ga * TJ + jw) * 32 + lane) * stride;          // this thread's nine cells
    const uint32_t *tab0 = cnt_base + (size_t) ((0 * TJ + jw) * 32 + lane) * stride;     // n(0, ., .)  of the triple
    const uint32_t *tab1 = cnt_base + (size_t) ((1 * TJ + jw) * 32 + lane) * stride;     // n(1, ., .)
    uint32_t *tabjk = reinterpret_cast<uint32_t *>(smem_raw + sm.njk) + (size_t) (jw * 32 + lane) * stride;   // n_jk(., .)
    const uint32_t *tile = reinterpret_cast<const uint32_t *>(smem_raw + sm.tile);
    const int fsplit = (F + 1) >> 1;                                  // folds [0, fsplit): the ga = 0 thread, the rest: its partner
    const int f_lo = ga == 0 ? 0 : fsplit, f_hi = ga == 0 ? fsplit : F;
    const RiskParams rp = risk_params(ctl->fl);
    uint32_t tile_phase = 0;

    // counts the nine cells pi x (gb, gc) of every block into dst (pi = plane `plane` of SNP i; PAIR: no pi, the pair table).
    // PAIR is a compile-time switch: a run-time test inside the cell loops would cut them into basic blocks of one cell
    // each and keep the scheduler from interleaving the cells' dependent LOP3 chains.
    auto count_tables = [&](auto pair_tag, const uint32_t *irows, int istride, int plane, uint32_t *dst) {
        constexpr bool PAIR = decltype(pair_tag)::value;
        uint32_t acc[9];
#pragma unroll
        for (int c = 0; c < 9; c++) acc[c] = 0;
        for (int ch = 0; ch < nchunks; ch++) {
            const uint32_t *jr = tile + ((size_t) ch * tile_rows + jw) * roww;
            const uint32_t *kr = tile + ((size_t) ch * tile_rows + TJ + lane) * roww;
            const uint32_t *ir = irows + (size_t) ch * istride;
            const int b_lo = ch * cb, b_hi = min(nblocks, b_lo + cb);
            for (int b = b_lo; b < b_hi; b++) {
                const int off = (b - b_lo) * 3 * SW;
                uint32_t pl[3][BW], pi[BW];
#pragma unroll
                for (int g = 0; g < 3; g++) load_plane<BW>(kr + off + g * SW, pl[g]);
                if constexpr (!PAIR) load_plane<BW>(ir + off + plane * SW, pi);
#pragma unroll
                for (int gb = 0; gb < 3; gb++) {
                    uint32_t pj[BW];
                    load_plane<BW>(jr + off + gb * SW, pj);
#pragma unroll
                    for (int gc = 0; gc < 3; gc++) {
                        if constexpr (SINGLE) {
                            // byte counters, four blocks to a word: the weight of the block is its byte's unit
                            const uint32_t kq = 1u << group_shift(b & 3);
                            uint32_t n = 0;
                            if constexpr (PAIR) n = cell_count2_acc<BW, 1u>(pj, pl[gc], 0u);
                            else n = cell_count3_acc<BW, 1u>(pi, pj, pl[gc], 0u);
                            acc[gb * 3 + gc] += n * kq;
                        } else {
                            if constexpr (PAIR) acc[gb * 3 + gc] = cell_count2_acc<BW, 1u>(pj, pl[gc], acc[gb * 3 + gc]);
                            else acc[gb * 3 + gc] = cell_count3_acc<BW, 1u>(pi, pj, pl[gc], acc[gb * 3 + gc]);
                        }
                    }
                }
                if constexpr (SINGLE) {
                    if ((b & 3) == 3) {
#pragma unroll
                        for (int c = 0; c < 9; c++) { dst[(b >> 2) * 9 + c] = acc[c]; acc[c] = 0; }
                    }
                } else {
                    const unsigned d = desc[b];
                    if (d & 0x8000u) {
                        const int seg = d & 0x7fff;
#pragma unroll
                        for (int c = 0; c < 9; c++) {
                            reinterpret_cast<uint16_t *>(dst + (seg >> 1) * 9 + c)[seg & 1] = (uint16_t) acc[c];
                            acc[c] = 0;
                        }
                    }
                }
            }
        }
    };

    int st = 0;
    uint32_t ph = 0;
    const uint32_t *miss = nullptr;                                 // the lists of the current stage's SNPs
    // sign = -1: the samples SNP i (row ii of the stage) is missing leave the cells (j, k) puts them in; +1: they return
    auto missing_fixup = [&](int ii, int sign) {            // each thread of the pair takes every other entry, both ways
        const uint32_t *ml = miss + (size_t) ii * a.v2_mcap;
        for (int m = ga; m < a.v2_mcap; m += 2) {
            const uint32_t e = ml[m];
            if (e == kMissEnd) break;                       // (the end marker fills the list's tail: either parity meets it)
            const int bit = e >> 27, off = (e >> 16) & 0x7ff, ch = (e >> 12) & 0xf, code = e & 0xfff;
            const uint32_t *jr = tile + ((size_t) ch * tile_rows + jw) * roww + off;
            const uint32_t *kr = tile + ((size_t) ch * tile_rows + TJ + lane) * roww + off;
            const uint32_t j1 = (jr[SW] >> bit) & 1u, j2 = (jr[2 * SW] >> bit) & 1u, jv = ((jr[0] >> bit) & 1u) | j1 | j2;
            const uint32_t k1 = (kr[SW] >> bit) & 1u, k2 = (kr[2 * SW] >> bit) & 1u, kv = ((kr[0] >> bit) & 1u) | k1 | k2;
            if (jv & kv) {
                const int c = (int) (j1 + 2 * j2) * 3 + (int) (k1 + 2 * k2);
                uint32_t delta, *word;
                if constexpr (SINGLE) { word = tabjk + (code & 0x3ff) * 9 + c; delta = 1u << group_shift(code >> 10); }
                else { word = tabjk + (code & 0x1f) * 9 + c; delta = 1u << (16 * (code >> 5)); }
                atomicAdd(word, sign < 0 ? 0u - delta : delta);     // (the partner's fix-up of the previous SNP may still be under way)
            }
        }
    };
    for (uint32_t s = 0;; s++) {
        const int nst = st + 1 == NS ? 0 : st + 1;
        int4 next = make_int4(-1, 0, 0, 0);
        if (producer) {
            next = next_desc();                                     // step s + 1
            ctl->meta[(s + 1) & 3] = next;
            // its stage was last read by step s + 1 - NS; a new unit's tile may only land once every warp is through step s
            if (s + 1 >= (uint32_t) NS && !(next.x >= 0 && (next.w >> 16))) mbar_wait(&ctl->empty[nst], ((s + 1 - NS) / NS) & 1);
            if (next.x >= 0 && !(next.w >> 16)) issue_rows(nst, next.x);
            else if (next.x < 0) mbar_arrive(&ctl->full[nst]);
        }
        __syncwarp();
        mbar_wait(&ctl->full[st], ph);
        const int4 meta = ctl->meta[s & 3];
        if (meta.x < 0) break;
        const int i0 = meta.x, j0 = meta.y, k0 = meta.z, ni = meta.w & 0xffff;
        if (meta.w >> 16) {
            // a new unit: its tile has been requested when the last step of the previous unit was done; count the pair table
            mbar_wait(&ctl->full[2], tile_phase);
            tile_phase ^= 1u;
            if (ga == 0) count_tables(std::true_type{}, tile, 0, 0, tabjk);
            pair_barrier(1 + jw);
        }
        // the best bound any CTA has published, every 32 steps, by a different warp each time
        if ((s & 31) == 0 && lane < F && warp == (int) ((s >> 5) % (uint32_t) nwarps)) refresh_threshold(ctl, a, lane);
        const uint32_t *stage = reinterpret_cast<const uint32_t *>(smem_raw + sm.ring0 + (size_t) st * sm.stage_bytes);
        miss = reinterpret_cast<const uint32_t *>(smem_raw + sm.ring0 + (size_t) st * sm.stage_bytes + sm.stage_rows_bytes);
        const int j = j0 + jw, k = k0 + lane;
        for (int ii = 0; ii < ni; ii++) {
            const int i = i0 + ii;
            count_tables(std::false_type{}, stage + (size_t) ii * roww, NI * roww, ga, mine);
            missing_fixup(ii, -1);                                  // SNP i's missing samples leave the pair table ...
            pair_barrier(1 + jw);                                   // both tables of the triple are complete

            bool valid = (i < j) && (j < k) && (k < a.nv);
            if (valid && (i <= a.edge_lo || i >= a.edge_hi)) {
                const uint64_t idx = triple_index((uint64_t) a.nv, (uint64_t) i, (uint64_t) j, (uint64_t) k);
                valid = idx >= a.first && idx < a.last;
            }
            if (__any_sync(0xffffffffu, valid)) {
                // totals over the folds: cells of i's genotypes 0, 1 and (derived) 2
                uint32_t tot[27];
#pragma unroll
                for (int c = 0; c < 27; c++) tot[c] = 0;
                for (int kk = 0; kk < nwc; kk++) {
#pragma unroll
                    for (int c = 0; c < 9; c++) {
                        const uint32_t w0 = tab0[kk * 9 + c], w1 = tab1[kk * 9 + c], w2 = tabjk[kk * 9 + c] - w0 - w1;
                        if constexpr (SINGLE) {
                            tot[c] += __dp4a(w0, 0x00000101u, 0u) | (__dp4a(w0, 0x01010000u, 0u) << 16);
                            tot[9 + c] += __dp4a(w1, 0x00000101u, 0u) | (__dp4a(w1, 0x01010000u, 0u) << 16);
                            tot[18 + c] += __dp4a(w2, 0x00000101u, 0u) | (__dp4a(w2, 0x01010000u, 0u) << 16);
                        } else {
                            tot[c] += w0; tot[9 + c] += w1; tot[18 + c] += w2;
                        }
                    }
                }
                for (int f = f_lo; f < f_hi; f++) {
                    const int kk = SINGLE ? (f >> 1) : f;
                    auto in_of = [&](int c) -> uint32_t {
                        const int g = c / 9, cc = c - g * 9;
                        const uint32_t w0 = tab0[kk * 9 + cc], w1 = tab1[kk * 9 + cc];
                        const uint32_t w = g == 0 ? w0 : (g == 1 ? w1 : tabjk[kk * 9 + cc] - w0 - w1);
                        return fold_pair<SINGLE>(w, f);
                    };
                    const int npos = a.training ? ctl->fl.A - ctl->fl.a_in[f] : ctl->fl.a_in[f];
                    const int nneg = a.training ? ctl->fl.U - ctl->fl.u_in[f] : ctl->fl.u_in[f];
                    const bool degenerate = (npos == 0 || nneg == 0);
                    if constexpr (BALANCED) {
                        if (a.prefilter) {
                            // score / n_f = sum over the cells of max(0, trA - trU); below the fold's bound nothing can be offered
                            int t = 0;
#pragma unroll
                            for (int c = 0; c < 27; c++) t += max(dp2a_lo_us(in_of(c), 0x000001FFu, dp2a_lo_us(tot[c], 0x0000FF01u, 0)), 0);
                            if (!__any_sync(0xffffffffu, valid && t >= *reinterpret_cast<volatile int *>(&ctl->tq[f]))) continue;
                        }
                        if (a.training) balanced_fold<27, true>(ctl, a, lists, tot, f, in_of, valid, i, j, k, lane);
                        else balanced_fold<27, false>(ctl, a, lists, tot, f, in_of, valid, i, j, k, lane);
                    } else {
                        int tp = 0, fp = 0;
                        uint32_t mask = 0;
#pragma unroll
                        for (int c = 0; c < 27; c++) {
                            const uint32_t in = in_of(c);
                            const int inA = (int) (in & 0xffffu), inU = (int) (in >> 16);
                            const int trA = (int) (tot[c] & 0xffffu) - inA, trU = (int) (tot[c] >> 16) - inU;
                            const bool r = high_risk(trA, trU, rp);
                            tp += r ? (a.training ? trA : inA) : 0;
                            fp += r ? (a.training ? trU : inU) : 0;
                            mask |= (r ? 1u : 0u) << c;
                        }
                        const long long score = degenerate ? LLONG_MIN
                                                : (a.eval_fn == kEvalBA ? ba_score(tp, fp, npos, nneg) : value_score(evaluate_fn(a.eval_fn, tp, npos - tp, fp, nneg - fp)));
                        offer_fold(ctl, a, lists, f, valid, score, degenerate, npos, nneg, i, j, k, mask, tp, fp, lane);
                    }
                }
            }
            pair_barrier(1 + jw);                                   // the partner is done reading this thread's table
            missing_fixup(ii, +1);                                  // ... and come back
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->empty[st]);
        if (producer && next.x >= 0 && (next.w >> 16)) {
            // the next step starts a new unit: wait until every warp has left this one, then fetch its tile and its first rows
            mbar_wait(&ctl->empty[st], (s / NS) & 1);
            // (the stages of the steps before this one have been released earlier: all of the ring is free)
            issue_tile(next.y, next.z);
            issue_rows(nst, next.x);
        }
        st = nst;
        if (st == 0) ph ^= 1u;
    }
    search_publish(ctl, a, lists);
}

// ============================================================================
// Order 3, balanced cohorts: one thread per triple, pair table in registers, two-pass epilogue
// ============================================================================
// Same tiling and staging as search3v2_kernel (unit = (TJ j rows, 32 k rows) resident, i rows streamed), but ONE thread
// counts both planes (0 and 1') of SNP i for its triple -- 18 cells, 18 x F counter words in its private slice -- and
// keeps the pair table n_jk (9 x F words, constant over the whole i loop of a unit) in REGISTERS: with eight warps per CTA
// a thread may use 255 of them.  Plane 1' = "genotype 1 or missing" of SNP i (valid & ~plane 0 & ~plane 2, one LOP3 per
// word; `valid` = the bit positions of a block that hold samples), so
//     n(2, gb, gc) = n_jk(gb, gc) - n(0, gb, gc) - n(1', gb, gc)
// is exact whatever is missing where, and what remains is to take the samples SNP i is missing out of n(1', ., .).  The
// epilogue does that in two passes over the cells, which needs no second table, no atomics and no undo:
//   pass 1  cells of genotypes 0 and 2 of SNP i (from n0, n1', n_jk)  ->  partial pre-filter sums t_f
//   fix-up  the thread's own table: n1' -> n1, one plain read-modify-write per missing sample of SNP i
//   pass 2  cells of genotype 1 (from n1)                             ->  t_f complete, compared with the fold's bound
// Only folds that pass (a handful per million triples) get the exact epilogue (risky cells, TP / FP, accuracy, list).
struct Smem3bMap {
    size_t tile, ring0, stage_bytes, stage_rows_bytes, vmask, counters, desc, lists, total;
};
__host__ __device__ inline Smem3bMap search3v3_smem_map(const FoldLayout &fl, int tj, int ni, int mcap, int rank, bool lists_in_smem) {
    Smem3bMap m;
    const size_t snp_bytes = (size_t) fl.nchunks * fl.row_words * 4;
    m.tile = align_up(sizeof(SearchCtl), 128);
    m.ring0 = m.tile + align_up((size_t) (tj + kTileJ) * snp_bytes, 128);
    m.stage_rows_bytes = (size_t) ni * snp_bytes;
    m.stage_bytes = align_up(m.stage_rows_bytes + (size_t) ni * mcap * 4, 128);
    m.vmask = m.ring0 + 2 * m.stage_bytes;
    m.counters = align_up(m.vmask + (size_t) fl.nblocks * slot_words(fl.w7 ? 7 : fl.bw) * 4, 16);
    const int nwc = fl.single ? fl.nblocks / 4 : fl.F;
    m.desc = m.counters + (size_t) counter_stride(18, nwc) * 4 * (size_t) (tj * 32);
    m.lists = align_up(m.desc + (fl.single ? 0 : (size_t) fl.nblocks * 2), 16);
    m.total = m.lists + (lists_in_smem ? (size_t) fl.F * rank * sizeof(Cand) : 0);
    return m;
}

// risk flags, TP / FP and risky-cell bits of nine cells of one fold (balanced cohorts): tot / in_of as in balanced_fold
template <int BITBASE, bool TRAINING, typename InOf>
__device__ __forceinline__ void risk_cells9(const uint32_t (&tot)[9], InOf in_of, uint32_t &tpfp, uint32_t &mask) {
    auto cell = [&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const uint32_t in = in_of(c);
        const uint32_t tr = tot[c] - in;
        const int d = dp2a_lo_us(tr, 0x0000FF01u, 0);
        risk_accumulate<(1u << (BITBASE + c))>(d, tr, TRAINING ? tr : in, tpfp, mask);
    };
    static_for<9>(cell);
}

template <int BW, bool SINGLE, int NWCMAX>
__global__ void __launch_bounds__(256, 1) search3v3_kernel(const SearchArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SearchCtl *ctl = reinterpret_cast<SearchCtl *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TJ = blockDim.x >> 5, jw = warp;
    const int NI = a.v2_ni;
    constexpr int NS = 2;
    const Smem3bMap sm = search3v3_smem_map(*a.fl, TJ, NI, a.v2_mcap, a.rank, a.lists_in_smem != 0);
    uint32_t *cnt_base = reinterpret_cast<uint32_t *>(smem_raw + sm.counters);
    uint32_t *vmask = reinterpret_cast<uint32_t *>(smem_raw + sm.vmask);
    uint16_t *desc = reinterpret_cast<uint16_t *>(smem_raw + sm.desc);
    Cand *lists = a.lists_in_smem ? reinterpret_cast<Cand *>(smem_raw + sm.lists) : a.lists + (size_t) blockIdx.x * a.fl->F * a.rank;
    search_init<SINGLE>(ctl, a, cnt_base, (sm.desc - sm.counters) / 4, desc);      // (barriers: full/empty[0..1] = the ring, full[2] = the tile)
    constexpr int SW = slot_words(BW);
    const int nblocks = ctl->fl.nblocks, cb = ctl->fl.cb, nchunks = ctl->fl.nchunks, roww = ctl->fl.row_words, F = ctl->fl.F;
    for (int x = tid; x < nblocks * SW; x += blockDim.x) vmask[x] = a.v3_vmask[x];
    __syncthreads();
    const int nwc = SINGLE ? nblocks / 4 : F;
    const int stride = counter_stride(18, nwc);
    const uint32_t row_bytes = (uint32_t) roww * 4;
    const int tile_rows = TJ + kTileJ;

    // ---- step sequencing: as in search3v2_kernel ----
    long long u = blockIdx.x;
    int cur_j0 = 0, cur_k0 = 0, cur_i = 0, cur_iend = 0;
    bool fresh = false;
    auto decode = [&]() {
        const int2 d = __ldg(a.unit_desc + u);
        cur_j0 = d.x; cur_k0 = d.y;
        cur_i = a.edge_lo;
        cur_iend = min(a.edge_hi, min(cur_j0 + TJ - 2, a.nv - 3)) + 1;
        fresh = true;
    };
    auto next_desc = [&]() {
        for (;;) {
            if (u >= a.num_units) return make_int4(-1, 0, 0, 0);
            if (cur_i < cur_iend) {
                const int n = min(NI, cur_iend - cur_i);
                const int4 d = make_int4(cur_i, cur_j0, cur_k0, n | (fresh ? 1 << 16 : 0));
                cur_i += n; fresh = false;
                return d;
            }
            u += gridDim.x;
            if (u < a.num_units) decode();
        }
    };
    auto issue_tile = [&](int j0, int k0) {
        uint8_t *dst = smem_raw + sm.tile;
        mbar_arrive_expect_tx(&ctl->full[2], (uint32_t) nchunks * tile_rows * row_bytes);
        for (int ch = 0; ch < nchunks; ch++) {
            const char *src = reinterpret_cast<const char *>(a.planes) + (int64_t) ch * a.snp_pad * row_bytes;
            bulk_g2s(dst + (size_t) ch * tile_rows * row_bytes, src + (int64_t) j0 * row_bytes, (uint32_t) TJ * row_bytes, &ctl->full[2]);
            bulk_g2s(dst + ((size_t) ch * tile_rows + TJ) * row_bytes, src + (int64_t) k0 * row_bytes, (uint32_t) kTileJ * row_bytes, &ctl->full[2]);
        }
    };
    auto issue_rows = [&](int st, int i0) {
        uint8_t *dst = smem_raw + sm.ring0 + (size_t) st * sm.stage_bytes;
        mbar_arrive_expect_tx(&ctl->full[st], (uint32_t) (nchunks * NI) * row_bytes + (uint32_t) (NI * a.v2_mcap * 4));
        for (int ch = 0; ch < nchunks; ch++) {
            const char *src = reinterpret_cast<const char *>(a.planes) + (int64_t) ch * a.snp_pad * row_bytes;
            bulk_g2s(dst + (size_t) ch * NI * row_bytes, src + (int64_t) i0 * row_bytes, (uint32_t) NI * row_bytes, &ctl->full[st]);
        }
        bulk_g2s(dst + sm.stage_rows_bytes, a.v2_miss + (int64_t) i0 * a.v2_mcap, (uint32_t) (NI * a.v2_mcap * 4), &ctl->full[st]);
    };
    const bool producer = (tid == blockDim.x - 32);
    if (producer) {
        if (u < a.num_units) decode();
        const int4 d = next_desc();
        ctl->meta[0] = d;
        if (d.x >= 0) { issue_tile(d.y, d.z); issue_rows(0, d.x); }
        else mbar_arrive(&ctl->full[0]);
    }

    uint32_t *mine = cnt_base + (size_t) (jw * 32 + lane) * stride;      // word (kk, c): mine[kk * 18 + c], c < 9: n0, else n1
    const uint32_t *tile = reinterpret_cast<const uint32_t *>(smem_raw + sm.tile);
    uint32_t njk[NWCMAX][9], totjk[9];
#pragma unroll
    for (int kk = 0; kk < NWCMAX; kk++)
#pragma unroll
        for (int c = 0; c < 9; c++) njk[kk][c] = 0;
#pragma unroll
    for (int c = 0; c < 9; c++) totjk[c] = 0;
    uint32_t tile_phase = 0;

    // totals (cases | controls << 16) of a counter word over the folds it holds
    auto word_total = [&](uint32_t w) -> uint32_t {
        if constexpr (SINGLE) return __dp4a(w, 0x00000101u, 0u) | (__dp4a(w, 0x01010000u, 0u) << 16);
        else return w;
    };
    // counts (PAIR: the nine cells of the pair table; else the 18 cells of planes 0 and 1' of the staged row) into dst
    auto count_tables = [&](auto pair_tag, const uint32_t *irows, int istride, uint32_t *dst) {
        constexpr bool PAIR = decltype(pair_tag)::value;
        uint32_t acc0[9], acc1[9];
#pragma unroll
        for (int c = 0; c < 9; c++) { acc0[c] = 0; acc1[c] = 0; }
        for (int ch = 0; ch < nchunks; ch++) {
            const uint32_t *jr = tile + ((size_t) ch * tile_rows + jw) * roww;
            const uint32_t *kr = tile + ((size_t) ch * tile_rows + TJ + lane) * roww;
            const uint32_t *ir = irows + (size_t) ch * istride;
            const int b_lo = ch * cb, b_hi = min(nblocks, b_lo + cb);
            for (int b = b_lo; b < b_hi; b++) {
                const int off = (b - b_lo) * 3 * SW;
                uint32_t pl[3][BW], p0[BW], p1[BW];
#pragma unroll
                for (int g = 0; g < 3; g++) load_plane<BW>(kr + off + g * SW, pl[g]);
                if constexpr (!PAIR) {
                    uint32_t p2[BW], vm[BW];
                    load_plane<BW>(ir + off, p0);
                    load_plane<BW>(ir + off + 2 * SW, p2);
                    load_plane<BW>(vmask + b * SW, vm);
#pragma unroll
                    for (int w = 0; w < BW; w++) asm("lop3.b32 %0, %1, %2, %3, 0x10;" : "=r"(p1[w]) : "r"(vm[w]), "r"(p0[w]), "r"(p2[w]));   // valid & ~p0 & ~p2
                }
                const uint32_t kq = SINGLE ? (1u << group_shift_rt((uint32_t) b & 3u)) : 1u;
#pragma unroll
                for (int gb = 0; gb < 3; gb++) {
                    uint32_t pj[BW];
                    load_plane<BW>(jr + off + gb * SW, pj);
#pragma unroll
                    for (int gc = 0; gc < 3; gc++) {
                        const int c = gb * 3 + gc;
                        if constexpr (PAIR) {
                            if constexpr (SINGLE) acc0[c] += cell_count2_acc<BW, 1u>(pj, pl[gc], 0u) * kq;
                            else acc0[c] = cell_count2_acc<BW, 1u>(pj, pl[gc], acc0[c]);
                        } else if constexpr (SINGLE) {
                            acc0[c] += cell_count3_acc<BW, 1u>(p0, pj, pl[gc], 0u) * kq;
                            acc1[c] += cell_count3_acc<BW, 1u>(p1, pj, pl[gc], 0u) * kq;
                        } else {
                            acc0[c] = cell_count3_acc<BW, 1u>(p0, pj, pl[gc], acc0[c]);
                            acc1[c] = cell_count3_acc<BW, 1u>(p1, pj, pl[gc], acc1[c]);
                        }
                    }
                }
                if constexpr (SINGLE) {
                    if ((b & 3) == 3) {
#pragma unroll
                        for (int c = 0; c < 9; c++) {
                            dst[(b >> 2) * 18 + c] = acc0[c]; acc0[c] = 0;
                            if constexpr (!PAIR) { dst[(b >> 2) * 18 + 9 + c] = acc1[c]; acc1[c] = 0; }
                        }
                    }
                } else {
                    const unsigned d = desc[b];
                    if (d & 0x8000u) {
                        const int seg = d & 0x7fff;
#pragma unroll
                        for (int c = 0; c < 9; c++) {
                            reinterpret_cast<uint16_t *>(dst + (seg >> 1) * 18 + c)[seg & 1] = (uint16_t) acc0[c]; acc0[c] = 0;
                            if constexpr (!PAIR) { reinterpret_cast<uint16_t *>(dst + (seg >> 1) * 18 + 9 + c)[seg & 1] = (uint16_t) acc1[c]; acc1[c] = 0; }
                        }
                    }
                }
            }
        }
    };

    const uint32_t *miss = nullptr;
    // the samples SNP i (row ii of the stage) is missing: out of (sign -1) / back into (+1) the thread's own table of plane 1'
    auto missing_fixup = [&](int ii, int sign) {
        const uint32_t *ml = miss + (size_t) ii * a.v2_mcap;
        for (int m = 0; m < a.v2_mcap; m++) {
            const uint32_t e = ml[m];
            if (e == kMissEnd) break;
            const int bit = e >> 27, off = (e >> 16) & 0x7ff, ch = (e >> 12) & 0xf, code = e & 0xfff;
            const uint32_t *jr = tile + ((size_t) ch * tile_rows + jw) * roww + off;
            const uint32_t *kr = tile + ((size_t) ch * tile_rows + TJ + lane) * roww + off;
            const uint32_t j1 = (jr[SW] >> bit) & 1u, j2 = (jr[2 * SW] >> bit) & 1u, jv = ((jr[0] >> bit) & 1u) | j1 | j2;
            const uint32_t k1 = (kr[SW] >> bit) & 1u, k2 = (kr[2 * SW] >> bit) & 1u, kv = ((kr[0] >> bit) & 1u) | k1 | k2;
            if (jv & kv) {
                const int c = (int) (j1 + 2 * j2) * 3 + (int) (k1 + 2 * k2);
                uint32_t delta, *word;
                if constexpr (SINGLE) { word = mine + (code & 0x3ff) * 18 + 9 + c; delta = 1u << group_shift_rt((uint32_t) code >> 10); }
                else { word = mine + (code & 0x1f) * 18 + 9 + c; delta = 1u << (16 * (code >> 5)); }
                *word += sign < 0 ? 0u - delta : delta;
            }
        }
    };
    auto njk_at = [&](int kk, int c) -> uint32_t {          // run-time kk (the rare exact path): select chain over the registers
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < NWCMAX; q++)
#pragma unroll
            for (int cc = 0; cc < 9; cc++) if (q == kk && cc == c) v = njk[q][cc];
        return v;
    };

    constexpr int NF = SINGLE ? 2 * NWCMAX : NWCMAX;        // folds a thread may have to carry partial sums for
    int st = 0;
    uint32_t ph = 0;
    for (uint32_t s = 0;; s++) {
        const int nst = st ^ 1;
        int4 next = make_int4(-1, 0, 0, 0);
        if (producer) {
            next = next_desc();
            ctl->meta[(s + 1) & 3] = next;
            if (s + 1 >= (uint32_t) NS && !(next.x >= 0 && (next.w >> 16))) mbar_wait(&ctl->empty[nst], ((s + 1 - NS) / NS) & 1);
            if (next.x >= 0 && !(next.w >> 16)) issue_rows(nst, next.x);
            else if (next.x < 0) mbar_arrive(&ctl->full[nst]);
        }
        __syncwarp();
        mbar_wait(&ctl->full[st], ph);
        const int4 meta = ctl->meta[s & 3];
        if (meta.x < 0) break;
        const int i0 = meta.x, j0 = meta.y, k0 = meta.z, ni = meta.w & 0xffff;
        if (meta.w >> 16) {
            // a new unit: count its pair table into the thread's slice, then take it into registers for the whole i loop
            mbar_wait(&ctl->full[2], tile_phase);
            tile_phase ^= 1u;
            count_tables(std::true_type{}, tile, 0, mine);
#pragma unroll
            for (int c = 0; c < 9; c++) totjk[c] = 0;
#pragma unroll
            for (int kk = 0; kk < NWCMAX; kk++)
                if (kk < nwc) {
#pragma unroll
                    for (int c = 0; c < 9; c++) { njk[kk][c] = mine[kk * 18 + c]; totjk[c] += word_total(njk[kk][c]); }
                }
        }
        if ((s & 31) == 0 && lane < F && warp == (int) ((s >> 5) % (uint32_t) TJ)) refresh_threshold(ctl, a, lane);
        const uint32_t *stage = reinterpret_cast<const uint32_t *>(smem_raw + sm.ring0 + (size_t) st * sm.stage_bytes);
        miss = reinterpret_cast<const uint32_t *>(smem_raw + sm.ring0 + (size_t) st * sm.stage_bytes + sm.stage_rows_bytes);
        const int j = j0 + jw, k = k0 + lane;
        for (int ii = 0; ii < ni; ii++) {
            const int i = i0 + ii;
            count_tables(std::false_type{}, stage + (size_t) ii * roww, NI * roww, mine);
            bool valid = (i < j) && (j < k) && (k < a.nv);
            if (valid && (i <= a.edge_lo || i >= a.edge_hi)) {
                const uint64_t idx = triple_index((uint64_t) a.nv, (uint64_t) i, (uint64_t) j, (uint64_t) k);
                valid = idx >= a.first && idx < a.last;
            }
            if (!__any_sync(0xffffffffu, valid)) continue;

            // ---- pass 1: cells of genotypes 0 and 2 of SNP i ----
            uint32_t tot0[9], tot2[9];
#pragma unroll
            for (int c = 0; c < 9; c++) { tot0[c] = 0; tot2[c] = totjk[c]; }
#pragma unroll
            for (int kk = 0; kk < NWCMAX; kk++)
                if (kk < nwc) {
#pragma unroll
                    for (int c = 0; c < 9; c++) {
                        const uint32_t w0 = word_total(mine[kk * 18 + c]);
                        tot0[c] += w0;
                        tot2[c] -= w0 + word_total(mine[kk * 18 + 9 + c]);
                    }
                }
            int t[NF];
#pragma unroll
            for (int f = 0; f < NF; f++) t[f] = 0;
            if (a.prefilter) {
                int D0[9], D2[9];
#pragma unroll
                for (int c = 0; c < 9; c++) { D0[c] = dp2a_lo_us(tot0[c], 0x0000FF01u, 0); D2[c] = dp2a_lo_us(tot2[c], 0x0000FF01u, 0); }
#pragma unroll
                for (int kk = 0; kk < NWCMAX; kk++)
                    if (kk < nwc) {
#pragma unroll
                        for (int c = 0; c < 9; c++) {
                            const uint32_t w0 = mine[kk * 18 + c], w2 = njk[kk][c] - w0 - mine[kk * 18 + 9 + c];
                            if constexpr (SINGLE) {
                                t[2 * kk] += max(dp4a_us(w0, 0x000100FFu, D0[c]), 0) + max(dp4a_us(w2, 0x000100FFu, D2[c]), 0);
                                t[2 * kk + 1] += max(dp4a_us(w0, 0x0100FF00u, D0[c]), 0) + max(dp4a_us(w2, 0x0100FF00u, D2[c]), 0);
                            } else {
                                t[kk] += max(dp2a_lo_us(w0, 0x000001FFu, D0[c]), 0) + max(dp2a_lo_us(w2, 0x000001FFu, D2[c]), 0);
                            }
                        }
                    }
            }
            // ---- the thread's own table of plane 1': out go the samples SNP i is missing ----
            missing_fixup(ii, -1);
            // ---- pass 2: cells of genotype 1 ----
            uint32_t tot1[9];
#pragma unroll
            for (int c = 0; c < 9; c++) tot1[c] = 0;
#pragma unroll
            for (int kk = 0; kk < NWCMAX; kk++)
                if (kk < nwc) {
#pragma unroll
                    for (int c = 0; c < 9; c++) tot1[c] += word_total(mine[kk * 18 + 9 + c]);
                }
            unsigned pass_folds = 0xffffffffu;                   // without the pre-filter every fold gets the exact epilogue
            if (a.prefilter) {
                int D1[9];
#pragma unroll
                for (int c = 0; c < 9; c++) D1[c] = dp2a_lo_us(tot1[c], 0x0000FF01u, 0);
#pragma unroll
                for (int kk = 0; kk < NWCMAX; kk++)
                    if (kk < nwc) {
#pragma unroll
                        for (int c = 0; c < 9; c++) {
                            const uint32_t w1 = mine[kk * 18 + 9 + c];
                            if constexpr (SINGLE) {
                                t[2 * kk] += max(dp4a_us(w1, 0x000100FFu, D1[c]), 0);
                                t[2 * kk + 1] += max(dp4a_us(w1, 0x0100FF00u, D1[c]), 0);
                            } else {
                                t[kk] += max(dp2a_lo_us(w1, 0x000001FFu, D1[c]), 0);
                            }
                        }
                    }
                pass_folds = 0;
#pragma unroll
                for (int f = 0; f < NF; f++)
                    if (f < F && __any_sync(0xffffffffu, valid && t[f] >= *reinterpret_cast<volatile int *>(&ctl->tq[f]))) pass_folds |= 1u << f;
            }
            // ---- exact epilogue of the folds that passed (rare) ----
            for (int f = 0; f < F; f++) {
                if (!((pass_folds >> f) & 1u)) continue;
                const int kk = SINGLE ? (f >> 1) : f;
                uint32_t tpfp = 0, mask = 0;
                auto in1 = [&](int c) -> uint32_t { return fold_pair<SINGLE>(mine[kk * 18 + 9 + c], f); };
                if (a.training) risk_cells9<9, true>(tot1, in1, tpfp, mask); else risk_cells9<9, false>(tot1, in1, tpfp, mask);
                missing_fixup(ii, +1);                           // back to plane 1': the derived cells are exact against it
                auto in0 = [&](int c) -> uint32_t { return fold_pair<SINGLE>(mine[kk * 18 + c], f); };
                auto in2 = [&](int c) -> uint32_t { return fold_pair<SINGLE>(njk_at(kk, c) - mine[kk * 18 + c] - mine[kk * 18 + 9 + c], f); };
                if (a.training) { risk_cells9<0, true>(tot0, in0, tpfp, mask); risk_cells9<18, true>(tot2, in2, tpfp, mask); }
                else { risk_cells9<0, false>(tot0, in0, tpfp, mask); risk_cells9<18, false>(tot2, in2, tpfp, mask); }
                missing_fixup(ii, -1);
                const int tp = (int) (tpfp & 0xffffu), fp = (int) (tpfp >> 16);
                const int npos = a.training ? ctl->fl.A - ctl->fl.a_in[f] : ctl->fl.a_in[f];
                const int nneg = a.training ? ctl->fl.U - ctl->fl.u_in[f] : ctl->fl.u_in[f];
                const bool degenerate = (npos == 0 || nneg == 0);
                const long long score = degenerate ? LLONG_MIN
                                        : (a.eval_fn == kEvalBA ? ba_score(tp, fp, npos, nneg) : value_score(evaluate_fn(a.eval_fn, tp, npos - tp, fp, nneg - fp)));
                offer_fold(ctl, a, lists, f, valid, score, degenerate, npos, nneg, i, j, k, mask, tp, fp, lane);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->empty[st]);
        if (producer && next.x >= 0 && (next.w >> 16)) {
            mbar_wait(&ctl->empty[st], (s / NS) & 1);
            issue_tile(next.y, next.z);
            issue_rows(nst, next.x);
        }
        st = nst;
        if (st == 0) ph ^= 1u;
    }
    search_publish(ctl, a, lists);
}

// ============================================================================
// Merge: many partial top-N lists -> the final top-N of every fold
// ============================================================================
// One CTA per fold.  Every entry gets a 128-bit key that realises the canonical
// order (BA descending, then SNP tuple ascending) as "larger key first"; tuples are
// unique within a fold, so keys are unique.  An MSB-first radix select (16 passes of
// 8 bits, shared-memory histograms) finds the key of the N-th best entry; the
// entries at or above it are then ranked by counting and written straight to their
// final position.
struct MergeArgs {
    const Cand *lists;          // [nlists][F][rank_in]
    const int *list_cnt;        // [nlists][F]  (nullptr: every list has rank_in entries, empty ones marked by i < 0)
    int nlists, F, rank_in, rank_out, training;
    const FoldLayout *fl;
    void *out;                  // hpgv_epi_model_t [F][rank_out]
    int order;
};

struct ModelOut {               // == hpgv_epi_model_t
    double accuracy;
    int32_t snp[3];
    uint32_t risky_mask;
    uint32_t conf[4];
};

__device__ __forceinline__ unsigned key_byte(const Key128 &k, int pass) {   // pass 0 = most significant byte
    return pass < 8 ? (unsigned) (k.hi >> (56 - 8 * pass)) & 0xffu : (unsigned) (k.lo >> (56 - 8 * (pass - 8))) & 0xffu;
}
__device__ __forceinline__ bool key_prefix_eq(const Key128 &k, const Key128 &p, int pass) {   // the first `pass` bytes agree
    if (pass == 0) return true;
    if (pass < 8) return (k.hi >> (64 - 8 * pass)) == (p.hi >> (64 - 8 * pass));
    if (k.hi != p.hi) return false;
    if (pass == 8) return true;
    return (k.lo >> (64 - 8 * (pass - 8))) == (p.lo >> (64 - 8 * (pass - 8)));
}

constexpr int kMergeThreads = 1024;
constexpr int kMergeSort = 8192;         // up to this many valid entries per fold are sorted in shared memory (148 lists x 50 = 7400)
// the sorted (key, slot) arrays and the selection arrays of the general path share the same bytes: a fold takes one path or the other
__host__ __device__ inline size_t merge_smem_bytes(int rank_out) {
    const size_t sel = (size_t) rank_out * 20 + 16, srt = (size_t) kMergeSort * 20;
    return sel > srt ? sel : srt;
}

static __global__ void __launch_bounds__(kMergeThreads) merge_kernel(const MergeArgs m) {
    extern __shared__ __align__(16) uint8_t msmem[];
    Key128 *selkey = reinterpret_cast<Key128 *>(msmem);                       // [rank_out]
    int *selidx = reinterpret_cast<int *>(msmem + (size_t) m.rank_out * 16);   // [rank_out]
    __shared__ unsigned hist[256];
    __shared__ Key128 prefix;
    __shared__ int want_sh, nsel, nvalid_sh;

    const int f = blockIdx.x, tid = threadIdx.x;
    const FoldLayout &fl = *m.fl;
    const int npos = m.training ? fl.A - fl.a_in[f] : fl.a_in[f];
    const int nneg = m.training ? fl.U - fl.u_in[f] : fl.u_in[f];
    const bool degenerate = (npos == 0 || nneg == 0);
    const int total = m.nlists * m.rank_in;
    auto entry = [&](int e) { return m.lists + ((size_t) (e / m.rank_in) * m.F + f) * m.rank_in + (e % m.rank_in); };
    auto entry_valid = [&](int e, Cand &c) {
        const int l = e / m.rank_in, n = e % m.rank_in;
        const int cnt = m.list_cnt ? m.list_cnt[(size_t) l * m.F + f] : m.rank_in;
        if (n >= cnt) return false;
        c = *entry(e);
        return c.i >= 0;
    };

    // valid entries are counted and, while they fit, compacted (key, slot) into shared memory
    Key128 *ckey = reinterpret_cast<Key128 *>(msmem);           // [kMergeSort]
    int *cidx = reinterpret_cast<int *>(ckey + kMergeSort);     // [kMergeSort]
    if (tid == 0) { prefix.hi = 0; prefix.lo = 0; nsel = 0; nvalid_sh = 0; }
    __syncthreads();
    for (int e = tid; e < total; e += kMergeThreads) {
        Cand c;
        if (!entry_valid(e, c)) continue;
        const int pos = atomicAdd(&nvalid_sh, 1);
        if (pos < kMergeSort) { ckey[pos] = cand_key(c, m.order); cidx[pos] = e; }
    }
    __syncthreads();
    if (nvalid_sh <= kMergeSort) {
        // The usual case (148 per-CTA lists of 50, or the lists of 8 ranks): a bitonic sort of the compacted keys, descending.
        // Its cost depends on the number of entries only through the power of two above it -- a rank whose lists are full
        // takes as long as one whose lists the score histogram kept short (the counting / radix-select it replaces took
        // 90 us and 255 us for the two halves of a 2-GPU pair range; every other rank waits for the slowest in the all-gather).
        const int n = nvalid_sh;
        int P = 32;
        while (P < n) P <<= 1;
        for (int t = n + tid; t < P; t += kMergeThreads) { ckey[t].hi = 0; ckey[t].lo = 0; cidx[t] = -1; }   // below every real key
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < P / 2; t += kMergeThreads) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const Key128 a = ckey[i], b = ckey[l];
                    const bool desc = (i & k) == 0;
                    if (desc ? key_gt(b, a) : key_gt(a, b)) {
                        ckey[i] = b; ckey[l] = a;
                        const int x = cidx[i]; cidx[i] = cidx[l]; cidx[l] = x;
                    }
                }
                __syncthreads();
            }
        }
        ModelOut *out = reinterpret_cast<ModelOut *>(m.out) + (size_t) f * m.rank_out;
        for (int t = tid; t < m.rank_out; t += kMergeThreads) {
            ModelOut r;
            if (t < n) {
                const Cand c = *entry(cidx[t]);
                r.accuracy = (degenerate || c.ba == -INFINITY) ? nan("") : c.ba;
                r.snp[0] = c.i; r.snp[1] = c.j; r.snp[2] = m.order == 3 ? c.k : -1;
                r.risky_mask = c.mask;
                r.conf[0] = (uint32_t) c.tp; r.conf[1] = (uint32_t) (npos - c.tp);
                r.conf[2] = (uint32_t) c.fp; r.conf[3] = (uint32_t) (nneg - c.fp);
            } else {
                r.accuracy = nan("");
                r.snp[0] = r.snp[1] = r.snp[2] = -1;
                r.risky_mask = 0;
                r.conf[0] = r.conf[1] = r.conf[2] = r.conf[3] = 0;
            }
            out[t] = r;
        }
        return;
    }
    __syncthreads();                     // the selection arrays below reuse the bytes of the compacted keys
    const int want = min(m.rank_out, nvalid_sh);
    if (tid == 0) want_sh = want;
    __syncthreads();

    if (want > 0) {
        for (int pass = 0; pass < 16; pass++) {
            for (int x = tid; x < 256; x += kMergeThreads) hist[x] = 0;
            __syncthreads();
            const Key128 p = prefix;
            for (int e = tid; e < total; e += kMergeThreads) {
                Cand c;
                if (!entry_valid(e, c)) continue;
                const Key128 k = cand_key(c, m.order);
                if (key_prefix_eq(k, p, pass)) atomicAdd(&hist[key_byte(k, pass)], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                int need = want_sh;               // among the entries matching the prefix we still need the `need` largest
                int bin = 255;
                for (; bin > 0; bin--) {
                    if ((int) hist[bin] >= need) break;
                    need -= (int) hist[bin];
                }
                want_sh = need;
                if (pass < 8) prefix.hi |= (unsigned long long) bin << (56 - 8 * pass);
                else prefix.lo |= (unsigned long long) bin << (56 - 8 * (pass - 8));
            }
            __syncthreads();
        }
        // prefix is now the key of the want-th best entry: collect everything at or above it
        const Key128 kth = prefix;
        for (int e = tid; e < total; e += kMergeThreads) {
            Cand c;
            if (!entry_valid(e, c)) continue;
            const Key128 k = cand_key(c, m.order);
            if (key_ge(k, kth)) {
                const int pos = atomicAdd(&nsel, 1);
                if (pos < m.rank_out) { selkey[pos] = k; selidx[pos] = e; }
            }
        }
    }
    __syncthreads();
    const int ns = min(nsel, m.rank_out);
    ModelOut *out = reinterpret_cast<ModelOut *>(m.out) + (size_t) f * m.rank_out;
    for (int t = tid; t < ns; t += kMergeThreads) {
        const Key128 k = selkey[t];
        int rank = 0;
        for (int o = 0; o < ns; o++) rank += key_gt(selkey[o], k) ? 1 : 0;
        const Cand c = *entry(selidx[t]);
        ModelOut r;
        r.accuracy = (degenerate || c.ba == -INFINITY) ? nan("") : c.ba;
        r.snp[0] = c.i; r.snp[1] = c.j; r.snp[2] = m.order == 3 ? c.k : -1;
        r.risky_mask = c.mask;
        r.conf[0] = (uint32_t) c.tp; r.conf[1] = (uint32_t) (npos - c.tp);
        r.conf[2] = (uint32_t) c.fp; r.conf[3] = (uint32_t) (nneg - c.fp);
        out[rank] = r;
    }
    for (int t = ns + tid; t < m.rank_out; t += kMergeThreads) {
        ModelOut r;
        r.accuracy = nan("");
        r.snp[0] = r.snp[1] = r.snp[2] = -1;
        r.risky_mask = 0;
        r.conf[0] = r.conf[1] = r.conf[2] = r.conf[3] = 0;
        out[t] = r;
    }
}

// hpgv_epi_model_t lists (all-gathered from the ranks) -> Cand lists for merge_kernel
static __global__ void models_to_cands_kernel(const ModelOut *in, int64_t n, Cand *out) {
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const ModelOut r = in[t];
    Cand c;
    c.ba = isnan(r.accuracy) ? -INFINITY : r.accuracy;
    c.i = r.snp[0]; c.j = r.snp[1]; c.k = r.snp[2];
    c.mask = r.risky_mask; c.tp = (int) r.conf[0]; c.fp = (int) r.conf[2];
    out[t] = c;
}

// ============================================================================
// Parity hook: explicit combinations, one warp each (simple on purpose)
// ============================================================================
static __global__ void eval_kernel(const uint32_t *__restrict__ planes, const uint16_t *__restrict__ blk_desc,
                            const FoldLayout *__restrict__ flp, int64_t snp_pad, int order, int training, int eval_fn,
                            int64_t ncomb, const int32_t *__restrict__ combs, const uint32_t *__restrict__ risky_in,
                            int32_t *counts_aff, int32_t *counts_unaff, uint32_t *risky_mask, uint32_t *conf, double *acc) {
    extern __shared__ int segcnt_all[];                 // [warps][nseg][C]
    const FoldLayout &fl = *flp;
    const int C = order == 2 ? 9 : 27;
    const int bw = fl.bw;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t comb = (int64_t) blockIdx.x * (blockDim.x >> 5) + wib;
    int *segcnt = segcnt_all + (size_t) wib * fl.nseg * C;
    if (comb >= ncomb) return;
    for (int x = lane; x < fl.nseg * C; x += 32) segcnt[x] = 0;
    __syncwarp();
    int64_t s[3];
    for (int o = 0; o < order; o++) s[o] = combs[comb * order + o];
    for (int b = 0; b < fl.nblocks; b++) {
        const int seg = blk_desc[b] & 0x7fff;
        if (seg >= fl.nseg) continue;                   // padding block of the single-block layout
        for (int x = lane; x < C * bw; x += 32) {
            const int c = x / bw, w = x % bw;
            uint32_t v = 0xffffffffu;
            int rem = c;
            for (int o = order - 1; o >= 0; o--) {
                const int g = rem % 3;
                rem /= 3;
                v &= logical_word(planes, fl, snp_pad, b, s[o], g, w);
            }
            atomicAdd(&segcnt[seg * C + c], __popc(v));
        }
    }
    __syncwarp();
    const RiskParams rp = risk_params(fl);
    for (int f = lane; f < fl.F; f += 32) {
        int tp = 0, fp = 0;
        uint32_t mask = 0;
        for (int c = 0; c < C; c++) {
            int totA = 0, totU = 0;
            for (int g = 0; g < fl.F; g++) { totA += segcnt[(2 * g) * C + c]; totU += segcnt[(2 * g + 1) * C + c]; }
            const int inA = segcnt[(2 * f) * C + c], inU = segcnt[(2 * f + 1) * C + c];
            const int trA = totA - inA, trU = totU - inU;
            // risky_in: the caller's own risky cells (confusion_matrix of model.c:337-460 takes them as an argument)
            const bool r = risky_in ? ((risky_in[comb * fl.F + f] >> c) & 1u) != 0 : high_risk(trA, trU, rp);
            tp += r ? (training ? trA : inA) : 0;
            fp += r ? (training ? trU : inU) : 0;
            mask |= (r ? 1u : 0u) << c;
            if (counts_aff) counts_aff[(comb * fl.F + f) * C + c] = trA;
            if (counts_unaff) counts_unaff[(comb * fl.F + f) * C + c] = trU;
        }
        const int npos = training ? fl.A - fl.a_in[f] : fl.a_in[f];
        const int nneg = training ? fl.U - fl.u_in[f] : fl.u_in[f];
        if (risky_mask) risky_mask[comb * fl.F + f] = mask;
        if (conf) {
            uint32_t *m = conf + (comb * fl.F + f) * 4;
            m[0] = (uint32_t) tp; m[1] = (uint32_t) (npos - tp); m[2] = (uint32_t) fp; m[3] = (uint32_t) (nneg - fp);
        }
        if (acc) acc[comb * fl.F + f] = evaluate_fn(eval_fn, tp, npos - tp, fp, nneg - fp);      // 0/0 = NaN like the reference
    }
}

// the device high-risk rule on explicit count pairs, and evaluate_model on explicit confusion matrices (parity hooks)
static __global__ void high_risk_kernel(const int32_t *__restrict__ ca, const int32_t *__restrict__ cu, int64_t n, int A, int U, int32_t *__restrict__ flags) {
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    RiskParams rp;
    rp.balanced = (A == U); rp.A = A; rp.U = U; rp.ratio = (float) A / (float) U;          // mdr.c:52
    flags[t] = high_risk(ca[t], cu[t], rp) ? 1 : 0;
}
static __global__ void evaluate_kernel(int eval_fn, int64_t n, const uint32_t *__restrict__ conf, double *__restrict__ out) {
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    out[t] = evaluate_fn(eval_fn, (int) conf[4 * t], (int) conf[4 * t + 1], (int) conf[4 * t + 2], (int) conf[4 * t + 3]);
}

// ============================================================================
// Pipe micro-benchmark (roofline denominator): independent chains per thread
// ============================================================================
template <int KIND>
__global__ void pipe_peak_kernel(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t x[8], acc[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { x[q] = seed * (threadIdx.x + 1) + q * 0x9E3779B9u + blockIdx.x; acc[q] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (KIND == 0) { acc[q] += __popc(x[q] ^ acc[q]); }                          // 1 POPC (+1 LOP +1 IADD)
                else if (KIND == 1) { acc[q] = xor3(acc[q], x[q], x[(q + 1) & 7]); x[q] = maj3(x[q], acc[q], x[(q + 3) & 7]); }   // 2 LOP3
                else { uint32_t t = xor3(acc[q], x[q], x[(q + 1) & 7]); x[q] = maj3(x[q], t, x[(q + 5) & 7]); acc[q] = t + __popc(t); }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) r ^= acc[q] ^ x[q];
    if (r == 0x12345678u) sink[0] = r;
}

}  // namespace hpgv
