// epi_kernels.cuh -- CUDA kernels of the epistasis engine (sm_100a).
//
//   pack_planes_kernel   bytes -> (fold, class)-segmented bit planes          (replaces set_genotypes_masks + get_k_folds_masks)
//   search2_kernel       exhaustive order-2 MDR: counts, risk, BA, top-N      (replaces process_set_of_combinations)
//   search3_kernel       same for order 3
//   merge_kernel         per-CTA / per-rank top-N lists -> final top-N        (replaces the heap drain + MPI tree merge)
//   eval_kernel          per-combination dump for explicit combinations       (parity hook)
//
// Design (DESIGN.md has the long form): a persistent CTA = 1 producer warp +
// NWARPS consumer warps.  The producer pulls work units from a global counter
// and streams, block by block along the sample axis, the bit planes of the
// unit's SNP rows into a shared-memory ring with 1-D bulk async copies
// (cp.async.bulk -> SASS UBLKCP) completing on mbarriers.  A consumer thread
// owns PPT SNP combinations (lane <-> last SNP of the tuple, so that row loads
// are one conflict-free LDS.128 per lane; the other SNPs are warp-uniform
// broadcast loads), ANDs the planes, compresses with LOP3 carry-save adders
// and POPCs, and writes one count per (cell, segment) into its private slice of
// shared memory.  After the last block the thread derives the whole-sample
// table, the F per-fold training tables (total - in-fold), the exact high-risk
// masks, TP/FP and an integer score per fold, and offers candidates that beat
// the running threshold to the CTA's top-N list.
#pragma once
#include "epi_device.cuh"

namespace hpgv {

// ============================================================================
// Packing
// ============================================================================
// One warp per (snp, block, word): lane l reads the genotype byte of the sample
// mapped to bit l and three ballots produce the three plane words.
// perm[pos] = dataset column of the sample at bit position pos, or -1 (padding).
template <int BW>
__global__ void pack_planes_kernel(const uint8_t *__restrict__ raw, int64_t nv, int64_t nsamples,
                                   const int32_t *__restrict__ perm, int nblocks, int64_t snp_pad,
                                   uint32_t *__restrict__ planes) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t words_per_snp = (int64_t) nblocks * BW;
    if (warp >= nv * words_per_snp) return;
    const int64_t snp = warp / words_per_snp;
    const int wb = (int) (warp % words_per_snp);
    const int b = wb / BW, w = wb % BW;
    const int32_t col = perm[(int64_t) wb * 32 + lane];
    const uint32_t g = col >= 0 ? raw[snp * nsamples + col] : 255u;
    const uint32_t m0 = __ballot_sync(0xffffffffu, g == 0);
    const uint32_t m1 = __ballot_sync(0xffffffffu, g == 1);
    const uint32_t m2 = __ballot_sync(0xffffffffu, g == 2);
    if (lane < 3) {
        int pos = w;
        if (BW == 8) pos = (((w >> 2) ^ swizzle_of(snp)) << 2) | (w & 3);
        planes[(((int64_t) b * snp_pad + snp) * 3 + lane) * BW + pos] = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
    }
}

// Inverse of the packer for one SNP: byte masks in the reference's layout
// [genotype][S_pad] (model.c:28-74).  One thread per sample column.
template <int BW>
__global__ void unpack_masks_kernel(const uint32_t *__restrict__ planes, int64_t snp, int64_t snp_pad,
                                    const int32_t *__restrict__ perm, int64_t npos, int A, int a_pad, int s_pad,
                                    uint8_t *__restrict__ out) {
    const int64_t pos = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= npos) return;
    const int32_t col = perm[pos];
    if (col < 0) return;
    const int b = (int) (pos / (32 * BW)), w = (int) ((pos / 32) % BW), bit = (int) (pos & 31);
    int wp = w;
    if (BW == 8) wp = (((w >> 2) ^ swizzle_of(snp)) << 2) | (w & 3);
    const int dst = col < A ? col : a_pad + (col - A);
    for (int g = 0; g < 3; g++) {
        uint32_t word = planes[(((int64_t) b * snp_pad + snp) * 3 + g) * BW + wp];
        out[(int64_t) g * s_pad + dst] = ((word >> bit) & 1u) ? 0xFF : 0x00;
    }
}

// ============================================================================
// Shared pieces of the search kernels
// ============================================================================
constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kSearchThreads = kConsumers + 32;   // + producer warp

struct __align__(16) SearchCtl {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    int4 meta[kStages];            // x = block index (-1: no more work), y/z/w = tile origins
    long long thr[kMaxFolds];      // score a candidate must reach to be offered to the list
    int lock[kMaxFolds];
    int cnt[kMaxFolds];
    int min_idx[kMaxFolds];
    FoldLayout fl;
};

// ---- top-N list maintenance (one list per CTA and fold, in global memory) ----
// Called by a full warp with a warp-uniform candidate.  Replaces
// add_to_model_ranking (model.c:481-521) with a deterministic order.
__device__ __forceinline__ void warp_offer(SearchCtl *ctl, const SearchArgs &a, int f, const Cand &c, int lane) {
    if (lane == 0) {
        while (atomicCAS(&ctl->lock[f], 0, 1) != 0) __nanosleep(32);
    }
    __syncwarp();
    __threadfence_block();
    Cand *list = a.lists + ((size_t) blockIdx.x * ctl->fl.F + f) * a.rank;
    // every lane reads the list state BEFORE lane 0 modifies it, so the decisions below are warp-uniform
    const int cnt = *reinterpret_cast<volatile int *>(&ctl->cnt[f]);
    bool replace = false;
    int mi = 0;
    if (cnt >= a.rank) {
        mi = *reinterpret_cast<volatile int *>(&ctl->min_idx[f]);
        const Cand worst = cand_load(list + mi);
        replace = cand_before(c, worst);
    }
    __syncwarp();
    bool rescan;
    if (cnt < a.rank) {
        if (lane == 0) {
            cand_store(list + cnt, c);
            *reinterpret_cast<volatile int *>(&ctl->cnt[f]) = cnt + 1;
        }
        rescan = (cnt + 1 == a.rank);
    } else {
        if (replace && lane == 0) cand_store(list + mi, c);
        rescan = replace;
    }
    __syncwarp();
    if (rescan) {
        // list is full: find the entry that ranks last; its score is the new threshold
        __threadfence_block();
        // sentinel that ranks before every real entry, so lanes without an entry never win
        Cand w;
        w.ba = INFINITY; w.i = -1; w.j = -1; w.k = -1; w.mask = 0; w.tp = 0; w.fp = 0;
        int widx = -1;
        for (int e = lane; e < a.rank; e += 32) {
            Cand x = cand_load(list + e);
            if (cand_before(w, x)) { w = x; widx = e; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Cand o;
            o.ba = __shfl_xor_sync(0xffffffffu, w.ba, off);
            o.i = __shfl_xor_sync(0xffffffffu, w.i, off);
            o.j = __shfl_xor_sync(0xffffffffu, w.j, off);
            o.k = __shfl_xor_sync(0xffffffffu, w.k, off);
            o.tp = __shfl_xor_sync(0xffffffffu, w.tp, off);
            o.fp = __shfl_xor_sync(0xffffffffu, w.fp, off);
            o.mask = 0;
            const int oidx = __shfl_xor_sync(0xffffffffu, widx, off);
            if (cand_before(w, o)) { w = o; widx = oidx; }
        }
        if (lane == 0) {
            const FoldLayout &fl = ctl->fl;
            const int npos = a.training ? fl.A - fl.a_in[f] : fl.a_in[f];
            const int nneg = a.training ? fl.U - fl.u_in[f] : fl.u_in[f];
            long long s = (npos == 0 || nneg == 0) ? LLONG_MIN : ba_score(w.tp, w.fp, npos, nneg);
            *reinterpret_cast<volatile int *>(&ctl->min_idx[f]) = widx;
            atomicMax(&ctl->thr[f], s);
            atomicMax(a.gthr + f, s);
        }
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) atomicExch(&ctl->lock[f], 0);
    __syncwarp();
}

// ---- per-combination epilogue -------------------------------------------------
// cnts: this thread's private counters, word (c, k) at cnts[(c * nwc + k) * kConsumers].
// CBITS == 8 : word k of cell c = counts of segments 4k..4k+3 = (A_2k, U_2k, A_2k+1, U_2k+1)
// CBITS == 16: word k of cell c = counts of segments 2k, 2k+1  = (A_k, U_k)
template <int NCELLS, int CBITS>
__device__ __forceinline__ void combo_epilogue(SearchCtl *ctl, const SearchArgs &a, const uint32_t *cnts, int nwc,
                                               bool valid, int si, int sj, int sk, int lane) {
    const FoldLayout &fl = ctl->fl;
    const RiskParams rp = risk_params(fl);
    const int nfolds = fl.F;
    int totA[NCELLS], totU[NCELLS];
#pragma unroll
    for (int c = 0; c < NCELLS; c++) { totA[c] = 0; totU[c] = 0; }
    for (int k = 0; k < nwc; k++) {
#pragma unroll
        for (int c = 0; c < NCELLS; c++) {
            uint32_t w = cnts[(c * nwc + k) * kConsumers];
            if constexpr (CBITS == 8) {
                totA[c] = __dp4a(w, 0x00010001u, (uint32_t) totA[c]);
                totU[c] = __dp4a(w, 0x01000100u, (uint32_t) totU[c]);
            } else {
                totA[c] += (int) (w & 0xffffu);
                totU[c] += (int) (w >> 16);
            }
        }
    }
    for (int f = 0; f < nfolds; f++) {
        int tp = 0, fp = 0;
        uint32_t mask = 0;
        const int k = (CBITS == 8) ? (f >> 1) : f;
        const int sh = (CBITS == 8) ? ((f & 1) * 16) : 0;
#pragma unroll
        for (int c = 0; c < NCELLS; c++) {
            uint32_t w = cnts[(c * nwc + k) * kConsumers] >> sh;
            int inA, inU;
            if constexpr (CBITS == 8) { inA = (int) (w & 0xffu); inU = (int) ((w >> 8) & 0xffu); }
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
        const long long score = degenerate ? LLONG_MIN : ba_score(tp, fp, npos, nneg);
        const long long thr = *reinterpret_cast<volatile long long *>(&ctl->thr[f]);
        unsigned want = __ballot_sync(0xffffffffu, valid && score >= thr);
        while (want) {
            const int src = __ffs(want) - 1;
            want &= want - 1;
            Cand c;
            c.i = __shfl_sync(0xffffffffu, si, src);
            c.j = __shfl_sync(0xffffffffu, sj, src);
            c.k = __shfl_sync(0xffffffffu, sk, src);
            c.mask = __shfl_sync(0xffffffffu, mask, src);
            c.tp = __shfl_sync(0xffffffffu, tp, src);
            c.fp = __shfl_sync(0xffffffffu, fp, src);
            c.ba = degenerate ? -INFINITY : balanced_accuracy(c.tp, c.fp, npos, nneg);
            warp_offer(ctl, a, f, c, lane);
        }
    }
}

template <int CBITS>
__device__ __forceinline__ void store_count(uint32_t *cnts, int c, int nwc, int seg, uint32_t v) {
    if constexpr (CBITS == 8) {
        reinterpret_cast<uint8_t *>(cnts + (c * nwc + (seg >> 2)) * kConsumers)[seg & 3] = (uint8_t) v;
    } else {
        reinterpret_cast<uint16_t *>(cnts + (c * nwc + (seg >> 1)) * kConsumers)[seg & 1] = (uint16_t) v;
    }
}

__host__ __device__ inline int words_per_cell(int nseg, int cbits) { return cbits == 8 ? (nseg + 3) / 4 : (nseg + 1) / 2; }

// dynamic shared memory: [SearchCtl][stage ring][counters]
template <int BW>
__host__ __device__ inline size_t stage_words(int rows) { return (size_t) rows * 3 * BW; }

// ============================================================================
// Order 2
// ============================================================================
// unit = (i-tile of TI = 8*PPT rows, j-tile of 32 rows); warp w, slot p <-> i = i0 + w*PPT + p; lane <-> j = j0 + lane
template <int BW, int CBITS, int PPT, bool SINGLE>
__global__ void __launch_bounds__(kSearchThreads) search2_kernel(const SearchArgs a) {
    constexpr int TI = kConsumerWarps * PPT;
    constexpr int ROWW = 3 * BW;                      // words per staged row
    constexpr int STAGEW = (TI + kTileJ) * ROWW;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SearchCtl *ctl = reinterpret_cast<SearchCtl *>(smem_raw);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(SearchCtl) + 127) / 128) * 128);
    uint32_t *cnt_base = stage + kStages * STAGEW;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nseg = a.fl->nseg, nblocks = a.fl->nblocks;
    const int nwc = words_per_cell(nseg, CBITS);

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], kConsumerWarps); }
        mbar_fence_init();
        ctl->fl = *a.fl;
    }
    if (tid < kMaxFolds) { ctl->thr[tid] = LLONG_MIN; ctl->lock[tid] = 0; ctl->cnt[tid] = 0; ctl->min_idx[tid] = 0; }
    for (int x = tid; x < PPT * 9 * nwc * kConsumers; x += kSearchThreads) cnt_base[x] = 0;   // padding bytes must read 0
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ------------------------------ producer ------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            const char *planes = reinterpret_cast<const char *>(a.planes);
            for (;;) {
                const unsigned long long u = atomicAdd(a.unit_counter, 1ULL);
                if (u >= (unsigned long long) a.num_units) break;
                int lo = 0, hi = a.n_it - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if ((unsigned long long) a.unit_prefix[mid] <= u) lo = mid; else hi = mid - 1;
                }
                const int i0 = (a.it0 + lo) * TI;
                const int j0 = (a.unit_jt0[lo] + (int) (u - (unsigned long long) a.unit_prefix[lo])) * kTileJ;
                for (int b = 0; b < nblocks; b++, it++) {
                    const int st = it % kStages;
                    mbar_wait(&ctl->empty[st], ((it / kStages) & 1) ^ 1);
                    ctl->meta[st] = make_int4(b, i0, j0, 0);
                    mbar_arrive_expect_tx(&ctl->full[st], STAGEW * 4);
                    const char *src = planes + (int64_t) b * a.snp_pad * (ROWW * 4);
                    uint32_t *dst = stage + st * STAGEW;
                    bulk_g2s(dst, src + (int64_t) i0 * (ROWW * 4), TI * ROWW * 4, &ctl->full[st]);
                    bulk_g2s(dst + TI * ROWW, src + (int64_t) j0 * (ROWW * 4), kTileJ * ROWW * 4, &ctl->full[st]);
                }
            }
            const int st = it % kStages;
            mbar_wait(&ctl->empty[st], ((it / kStages) & 1) ^ 1);
            ctl->meta[st] = make_int4(-1, 0, 0, 0);
            mbar_arrive(&ctl->full[st]);
        }
    } else {
        // ------------------------------ consumers ------------------------------
        uint32_t *cnts = cnt_base + tid;          // + (p*9 + c) * nwc * kConsumers + k * kConsumers
        uint32_t acc[PPT][9];
#pragma unroll
        for (int p = 0; p < PPT; p++)
#pragma unroll
            for (int c = 0; c < 9; c++) acc[p][c] = 0;

        for (uint32_t it = 0;; it++) {
            const int st = it % kStages;
            mbar_wait(&ctl->full[st], (it / kStages) & 1);
            const int4 meta = ctl->meta[st];
            if (meta.x < 0) break;
            const int b = meta.x, i0 = meta.y, j0 = meta.z;
            const uint32_t *sbase = stage + st * STAGEW;
            const unsigned desc = a.blk_desc[b];
            const int seg = desc & 0x7fff;

            uint32_t pj[3][BW];
            {
                const uint32_t *jrow = sbase + (TI + lane) * ROWW;
                const int swz = swizzle_of(j0 + lane);
#pragma unroll
                for (int g = 0; g < 3; g++) load_plane<BW>(jrow, g, swz, pj[g]);
            }
#pragma unroll
            for (int p = 0; p < PPT; p++) {
                const int il = warp * PPT + p;
                const uint32_t *irow = sbase + il * ROWW;
                const int swz = swizzle_of(i0 + il);
#pragma unroll
                for (int ga = 0; ga < 3; ga++) {
                    uint32_t pi[BW];
                    load_plane<BW>(irow, ga, swz, pi);
#pragma unroll
                    for (int gb = 0; gb < 3; gb++) {
                        const uint32_t n = cell_count2<BW>(pi, pj[gb]);
                        if constexpr (SINGLE) store_count<CBITS>(cnts + (p * 9) * nwc * kConsumers, ga * 3 + gb, nwc, seg, n);
                        else acc[p][ga * 3 + gb] += n;
                    }
                }
            }
            if constexpr (!SINGLE) {
                if (desc & 0x8000u) {
#pragma unroll
                    for (int p = 0; p < PPT; p++)
#pragma unroll
                        for (int c = 0; c < 9; c++) {
                            store_count<CBITS>(cnts + (p * 9) * nwc * kConsumers, c, nwc, seg, acc[p][c]);
                            acc[p][c] = 0;
                        }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->empty[st]);

            if (b == 0 && warp == 0 && lane < ctl->fl.F) atomicMax(&ctl->thr[lane], __ldcg(a.gthr + lane));

            if (b == nblocks - 1) {
                const int j = j0 + lane;
#pragma unroll 1
                for (int p = 0; p < PPT; p++) {
                    const int i = i0 + warp * PPT + p;
                    bool valid = (i < j) && (j < a.nv);
                    if (valid) {
                        const uint64_t idx = pair_index((uint64_t) a.nv, (uint64_t) i, (uint64_t) j);
                        valid = idx >= a.first && idx < a.last;
                    }
                    combo_epilogue<9, CBITS>(ctl, a, cnts + (p * 9) * nwc * kConsumers, nwc, valid, i, j, -1, lane);
                }
            }
        }
        // all consumer warps are done inserting: publish list sizes
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
        if (tid < ctl->fl.F) a.list_cnt[(size_t) blockIdx.x * ctl->fl.F + tid] = ctl->cnt[tid];
    }
}

// ============================================================================
// Order 3
// ============================================================================
// unit = (i, j-tile of 8 rows, k-tile of 32 rows); warp w <-> j = j0 + w; lane <-> k = k0 + lane; one triple per thread.
// Work list: producer takes (i, j-tile) super-units from the global counter and walks the k-tiles itself.
template <int BW, int CBITS, bool SINGLE>
__global__ void __launch_bounds__(kSearchThreads) search3_kernel(const SearchArgs a) {
    constexpr int TJ = kConsumerWarps;                // 8 j rows
    constexpr int ROWW = 3 * BW;
    constexpr int ROWS = 1 + TJ + kTileJ;             // i row, j rows, k rows
    constexpr int STAGEW = ((ROWS * ROWW + 3) / 4) * 4;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SearchCtl *ctl = reinterpret_cast<SearchCtl *>(smem_raw);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(SearchCtl) + 127) / 128) * 128);
    uint32_t *cnt_base = stage + kStages * STAGEW;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nseg = a.fl->nseg, nblocks = a.fl->nblocks;
    const int nwc = words_per_cell(nseg, CBITS);

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], kConsumerWarps); }
        mbar_fence_init();
        ctl->fl = *a.fl;
    }
    if (tid < kMaxFolds) { ctl->thr[tid] = LLONG_MIN; ctl->lock[tid] = 0; ctl->cnt[tid] = 0; ctl->min_idx[tid] = 0; }
    for (int x = tid; x < 27 * nwc * kConsumers; x += kSearchThreads) cnt_base[x] = 0;
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            uint32_t it = 0;
            const char *planes = reinterpret_cast<const char *>(a.planes);
            const int nkt = (a.nv + kTileJ - 1) / kTileJ;
            for (;;) {
                const unsigned long long u = atomicAdd(a.unit_counter, 1ULL);
                if (u >= (unsigned long long) a.num_units) break;
                // super-unit u -> (i, j-tile): prefix over i rows (it0 = first row)
                int lo = 0, hi = a.n_it - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if ((unsigned long long) a.unit_prefix[mid] <= u) lo = mid; else hi = mid - 1;
                }
                const int i = a.it0 + lo;
                const int j0 = (a.unit_jt0[lo] + (int) (u - (unsigned long long) a.unit_prefix[lo])) * TJ;
                // k-tiles that can hold k > j0 + 1 ... (first k is j0 + 1 at the earliest)
                for (int kt = (j0 + 1) / kTileJ; kt < nkt; kt++) {
                    const int k0 = kt * kTileJ;
                    for (int b = 0; b < nblocks; b++, it++) {
                        const int st = it % kStages;
                        mbar_wait(&ctl->empty[st], ((it / kStages) & 1) ^ 1);
                        ctl->meta[st] = make_int4(b, i, j0, k0);
                        mbar_arrive_expect_tx(&ctl->full[st], ROWS * ROWW * 4);
                        const char *src = planes + (int64_t) b * a.snp_pad * (ROWW * 4);
                        uint32_t *dst = stage + st * STAGEW;
                        bulk_g2s(dst, src + (int64_t) i * (ROWW * 4), ROWW * 4, &ctl->full[st]);
                        bulk_g2s(dst + ROWW, src + (int64_t) j0 * (ROWW * 4), TJ * ROWW * 4, &ctl->full[st]);
                        bulk_g2s(dst + (1 + TJ) * ROWW, src + (int64_t) k0 * (ROWW * 4), kTileJ * ROWW * 4, &ctl->full[st]);
                    }
                }
            }
            const int st = it % kStages;
            mbar_wait(&ctl->empty[st], ((it / kStages) & 1) ^ 1);
            ctl->meta[st] = make_int4(-1, 0, 0, 0);
            mbar_arrive(&ctl->full[st]);
        }
    } else {
        uint32_t *cnts = cnt_base + tid;
        uint32_t acc[27];
#pragma unroll
        for (int c = 0; c < 27; c++) acc[c] = 0;

        for (uint32_t it = 0;; it++) {
            const int st = it % kStages;
            mbar_wait(&ctl->full[st], (it / kStages) & 1);
            const int4 meta = ctl->meta[st];
            if (meta.x < 0) break;
            const int b = meta.x, i = meta.y, j0 = meta.z, k0 = meta.w;
            const uint32_t *sbase = stage + st * STAGEW;
            const unsigned desc = a.blk_desc[b];
            const int seg = desc & 0x7fff;
            const int j = j0 + warp, k = k0 + lane;

            uint32_t pk[3][BW];
            {
                const uint32_t *krow = sbase + (1 + TJ + lane) * ROWW;
                const int swz = swizzle_of(k);
#pragma unroll
                for (int g = 0; g < 3; g++) load_plane<BW>(krow, g, swz, pk[g]);
            }
            const uint32_t *irow = sbase;
            const uint32_t *jrow = sbase + (1 + warp) * ROWW;
            const int swz_i = swizzle_of(i), swz_j = swizzle_of(j);
#pragma unroll
            for (int ga = 0; ga < 3; ga++) {
                uint32_t pi[BW];
                load_plane<BW>(irow, ga, swz_i, pi);
#pragma unroll
                for (int gb = 0; gb < 3; gb++) {
                    uint32_t pj[BW];
                    load_plane<BW>(jrow, gb, swz_j, pj);
#pragma unroll
                    for (int gc = 0; gc < 3; gc++) {
                        const uint32_t n = cell_count3<BW>(pi, pj, pk[gc]);
                        const int c = ga * 9 + gb * 3 + gc;
                        if constexpr (SINGLE) store_count<CBITS>(cnts, c, nwc, seg, n);
                        else acc[c] += n;
                    }
                }
            }
            if constexpr (!SINGLE) {
                if (desc & 0x8000u) {
#pragma unroll
                    for (int c = 0; c < 27; c++) { store_count<CBITS>(cnts, c, nwc, seg, acc[c]); acc[c] = 0; }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->empty[st]);

            if (b == 0 && warp == 0 && lane < ctl->fl.F) atomicMax(&ctl->thr[lane], __ldcg(a.gthr + lane));

            if (b == nblocks - 1) {
                bool valid = (i < j) && (j < k) && (k < a.nv);
                if (valid) {
                    const uint64_t idx = triple_index((uint64_t) a.nv, (uint64_t) i, (uint64_t) j, (uint64_t) k);
                    valid = idx >= a.first && idx < a.last;
                }
                combo_epilogue<27, CBITS>(ctl, a, cnts, nwc, valid, i, j, k, lane);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
        if (tid < ctl->fl.F) a.list_cnt[(size_t) blockIdx.x * ctl->fl.F + tid] = ctl->cnt[tid];
    }
}

// ============================================================================
// Merge: many partial top-N lists -> the final top-N of every fold
// ============================================================================
// One CTA per fold.  Entries below the global threshold cannot be in the top N;
// the survivors are ranked by counting (tuples are unique within a fold, so the
// canonical order is total) and written straight to their final position.
struct MergeArgs {
    const Cand *lists;          // [nlists][F][rank_in]
    const int *list_cnt;        // [nlists][F]  (nullptr: every list has rank_in entries, empty ones marked by i < 0)
    const long long *gthr;      // [F] or nullptr
    int nlists, F, rank_in, rank_out, training;
    const FoldLayout *fl;
    void *out;                  // hpgv_epi_model_t [F][rank_out]
    int order;
    int *sel;                   // scratch [F][nlists * rank_in]: indices of surviving entries
};

struct ModelOut {               // == hpgv_epi_model_t
    double accuracy;
    int32_t snp[3];
    uint32_t risky_mask;
    uint32_t conf[4];
};

__global__ void __launch_bounds__(1024) merge_kernel(const MergeArgs m) {
    __shared__ int nsel;
    const int f = blockIdx.x;
    int *sel = m.sel + (size_t) f * m.nlists * m.rank_in;
    const FoldLayout &fl = *m.fl;
    const int npos = m.training ? fl.A - fl.a_in[f] : fl.a_in[f];
    const int nneg = m.training ? fl.U - fl.u_in[f] : fl.u_in[f];
    const bool degenerate = (npos == 0 || nneg == 0);
    const long long thr = m.gthr ? m.gthr[f] : LLONG_MIN;
    if (threadIdx.x == 0) nsel = 0;
    __syncthreads();
    const int total = m.nlists * m.rank_in;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int l = e / m.rank_in, n = e % m.rank_in;
        const int cnt = m.list_cnt ? m.list_cnt[(size_t) l * m.F + f] : m.rank_in;
        if (n >= cnt) continue;
        const Cand c = cand_load(m.lists + ((size_t) l * m.F + f) * m.rank_in + n);
        if (c.i < 0) continue;
        const long long s = degenerate ? LLONG_MIN : ba_score(c.tp, c.fp, npos, nneg);
        if (s >= thr) sel[atomicAdd(&nsel, 1)] = e;
    }
    __syncthreads();
    const int ns = nsel;
    ModelOut *out = reinterpret_cast<ModelOut *>(m.out) + (size_t) f * m.rank_out;
    for (int t = threadIdx.x; t < ns; t += blockDim.x) {
        const int e = sel[t];
        const Cand c = cand_load(m.lists + ((size_t) (e / m.rank_in) * m.F + f) * m.rank_in + (e % m.rank_in));
        int rank = 0;
        for (int o = 0; o < ns; o++) {
            const int e2 = sel[o];
            const Cand d = cand_load(m.lists + ((size_t) (e2 / m.rank_in) * m.F + f) * m.rank_in + (e2 % m.rank_in));
            rank += cand_before(d, c) ? 1 : 0;
        }
        if (rank < m.rank_out) {
            ModelOut r;
            r.accuracy = degenerate ? nan("") : c.ba;
            r.snp[0] = c.i; r.snp[1] = c.j; r.snp[2] = m.order == 3 ? c.k : -1;
            r.risky_mask = c.mask;
            r.conf[0] = (uint32_t) c.tp; r.conf[1] = (uint32_t) (npos - c.tp);
            r.conf[2] = (uint32_t) c.fp; r.conf[3] = (uint32_t) (nneg - c.fp);
            out[rank] = r;
        }
    }
    for (int t = (ns < m.rank_out ? ns : m.rank_out) + threadIdx.x; t < m.rank_out; t += blockDim.x) {
        ModelOut r;
        r.accuracy = nan("");
        r.snp[0] = r.snp[1] = r.snp[2] = -1;
        r.risky_mask = 0;
        r.conf[0] = r.conf[1] = r.conf[2] = r.conf[3] = 0;
        out[t] = r;
    }
}

// hpgv_epi_model_t lists (all-gathered from the ranks) -> Cand lists for merge_kernel
__global__ void models_to_cands_kernel(const ModelOut *in, int64_t n, Cand *out) {
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
template <int BW>
__global__ void eval_kernel(const uint32_t *__restrict__ planes, const uint16_t *__restrict__ blk_desc,
                            const FoldLayout *__restrict__ flp, int64_t snp_pad, int order, int training,
                            int64_t ncomb, const int32_t *__restrict__ combs,
                            int32_t *counts_aff, int32_t *counts_unaff, uint32_t *risky_mask, uint32_t *conf, double *acc) {
    extern __shared__ int segcnt_all[];                 // [warps][nseg][C]
    const FoldLayout &fl = *flp;
    const int C = order == 2 ? 9 : 27;
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
        for (int x = lane; x < C * BW; x += 32) {
            const int c = x / BW, w = x % BW;
            uint32_t v = 0xffffffffu;
            int rem = c;
            for (int o = order - 1; o >= 0; o--) {
                const int g = rem % 3;
                rem /= 3;
                int wp = w;
                if (BW == 8) wp = (((w >> 2) ^ swizzle_of(s[o])) << 2) | (w & 3);
                v &= planes[(((int64_t) b * snp_pad + s[o]) * 3 + g) * BW + wp];
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
            const bool r = high_risk(trA, trU, rp);
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
        if (acc) acc[comb * fl.F + f] = (npos == 0 || nneg == 0) ? nan("") : balanced_accuracy(tp, fp, npos, nneg);
    }
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
