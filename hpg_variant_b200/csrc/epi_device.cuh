// epi_device.cuh -- device helpers: mbarrier / bulk-copy PTX, LOP3 carry-save
// counting, the exact high-risk rule, balanced accuracy, candidate ordering.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <limits.h>
#include "epi_types.h"

namespace hpgv {

// ----------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -- sm_90+/sm_100a
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// the waiting warp is suspended by the hardware (up to the hinted time) instead of spinning through issue slots
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HPGV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra HPGV_DONE;\n"
        "bra HPGV_WAIT;\n"
        "HPGV_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
}
// a 16-byte shared-memory read the compiler may neither hoist nor merge with an earlier read of the same address
__device__ __forceinline__ int4 ld_volatile_shared_int4(const int4 *p) {
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}
// global -> shared, completion signalled on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ----------------------------------------------------------------------------
// bit counting: AND the planes, compress with carry-save adders (LOP3 on the
// ALU pipe), POPC (XU pipe) the compressed words, weight and accumulate with
// IMAD (FMA pipe).  Per block of BW words: BW=4 -> 3 POPC, BW=8 -> 4 POPC,
// BW=7 (8-word slots, last word empty) -> 3 POPC.
// ----------------------------------------------------------------------------
// words a block occupies in a row (BW = 7 counts 7 words of an 8-word slot)
__host__ __device__ constexpr int slot_words(int BW) { return BW == 7 ? 8 : BW; }

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t and3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// acc + v * K with K a compile-time constant (kept as IMAD so that it issues on the FMA pipe, not the ALU)
template <uint32_t K>
__device__ __forceinline__ uint32_t mad_const(uint32_t v, uint32_t acc) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "n"(K), "r"(acc));
    return d;
}

// acc += K * (number of set bits in x[0..BW))
template <int BW, uint32_t K>
__device__ __forceinline__ uint32_t count_words_acc(const uint32_t (&x)[BW], uint32_t acc) {
    if constexpr (BW == 4) {
        const uint32_t ones = xor3(x[0], x[1], x[2]);
        const uint32_t twos = maj3(x[0], x[1], x[2]);
        acc = mad_const<K>(__popc(ones), acc);
        acc = mad_const<K>(__popc(x[3]), acc);
        acc = mad_const<2 * K>(__popc(twos), acc);
        return acc;
    } else if constexpr (BW == 7) {
        // 8-word slots whose last word is never populated (segments of <= 224 samples): 7 words -> 3 POPC
        const uint32_t s1 = xor3(x[0], x[1], x[2]), c1 = maj3(x[0], x[1], x[2]);
        const uint32_t s2 = xor3(x[3], x[4], x[5]), c2 = maj3(x[3], x[4], x[5]);
        const uint32_t s3 = xor3(s1, s2, x[6]),     c3 = maj3(s1, s2, x[6]);
        const uint32_t t  = xor3(c1, c2, c3),       f  = maj3(c1, c2, c3);
        acc = mad_const<K>(__popc(s3), acc);
        acc = mad_const<2 * K>(__popc(t), acc);
        acc = mad_const<4 * K>(__popc(f), acc);
        return acc;
    } else {
        static_assert(BW == 8, "block width must be 4, 7 (of 8) or 8 words");
        const uint32_t s1 = xor3(x[0], x[1], x[2]), c1 = maj3(x[0], x[1], x[2]);
        const uint32_t s2 = xor3(x[3], x[4], x[5]), c2 = maj3(x[3], x[4], x[5]);
        const uint32_t s3 = xor3(s1, s2, x[6]),     c3 = maj3(s1, s2, x[6]);
        const uint32_t t  = xor3(c1, c2, c3),       f  = maj3(c1, c2, c3);
        acc = mad_const<K>(__popc(s3), acc);
        acc = mad_const<K>(__popc(x[7]), acc);
        acc = mad_const<2 * K>(__popc(t), acc);
        acc = mad_const<4 * K>(__popc(f), acc);
        return acc;
    }
}

template <int BW, uint32_t K>
__device__ __forceinline__ uint32_t cell_count2_acc(const uint32_t (&a)[BW], const uint32_t (&b)[BW], uint32_t acc) {
    uint32_t x[BW];
#pragma unroll
    for (int w = 0; w < BW; w++) x[w] = a[w] & b[w];
    return count_words_acc<BW, K>(x, acc);
}
template <int BW, uint32_t K>
__device__ __forceinline__ uint32_t cell_count3_acc(const uint32_t (&a)[BW], const uint32_t (&b)[BW], const uint32_t (&c)[BW], uint32_t acc) {
    uint32_t x[BW];
#pragma unroll
    for (int w = 0; w < BW; w++) x[w] = and3(a[w], b[w], c[w]);
    return count_words_acc<BW, K>(x, acc);
}

// One plane (BW words) from a staged row in shared memory (16-byte aligned).
template <int BW>
__device__ __forceinline__ void load_plane(const uint32_t *p, uint32_t (&v)[BW]) {
#pragma unroll
    for (int q = 0; q < slot_words(BW) / 4; q++) {
        const uint4 t = *reinterpret_cast<const uint4 *>(p + 4 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z;
        if (4 * q + 3 < BW) v[4 * q + 3] = t.w;
    }
}

// word offset of (chunk, snp, block-in-chunk, plane, word) inside the packed planes
__host__ __device__ __forceinline__ int64_t plane_word(const FoldLayout &fl, int64_t snp_pad, int b, int64_t snp, int g, int w) {
    const int ch = b / fl.cb, bl = b % fl.cb;
    return ((int64_t) ch * snp_pad + snp) * fl.row_words + (bl * 3 + g) * fl.bw + w;
}

// ---- tri layout (FoldLayout::tri) ---------------------------------------------
// One chunk.  A row is ngroups = nblocks/4 groups of 36 words followed by the tails and the per-group marginals:
//   main word (b, g, w), w < 3 : (b/4)*36 + g*12 + (b%4)*3 + w      (a group = 3 planes x 4 blocks x 3 words = 9 LDS.128)
//   tail word (m, g), m = b/8  : ngroups*36 + m*4 + g               (one LDS.128 = the three planes of eight blocks' tails)
// The tail of block b = 4k + q sits in the nibble that the byte counters want: counter word k keeps block q in byte
// group_shift(q)/8; tail word m = k/2 keeps it in nibble 2*byte + (k&1), so that after a nibble-wise popcount n,
// n & 0x0F0F0F0F adds to counter word 2m and (n >> 4) & 0x0F0F0F0F to counter word 2m + 1.
__host__ __device__ constexpr uint32_t group_shift(int q) { return q == 0 ? 0u : (q == 1 ? 16u : (q == 2 ? 8u : 24u)); }
// the same for a run-time block index: q = 0, 1, 2, 3 -> 0, 16, 8, 24
__host__ __device__ __forceinline__ uint32_t group_shift_rt(uint32_t q) { return ((q & 1u) << 4) | ((q & 2u) << 2); }
__host__ __device__ __forceinline__ int tri_word_off(int b, int g, int w) { return (b >> 2) * 36 + g * 12 + (b & 3) * 3 + w; }
__host__ __device__ __forceinline__ int tri_tail_off(int nblocks, int b, int g) { return (nblocks >> 2) * 36 + (b >> 3) * 4 + g; }
__host__ __device__ __forceinline__ int tri_tail_shift(int b) { return (int) (group_shift(b & 3) / 8 * 2 + ((b >> 2) & 1)) * 4; }
// After the tails, one LDS.128 per group k: words (N_0, N_1, N_2, missing) -- N_g = the SNP's own per-block counts of
// genotype g, four byte counters in the layout of the cell counters; missing = 0xFF in the byte of every block in which
// the SNP has a sample that is in no plane.  In a block without missing samples of SNP i the cells of one genotype of i
// follow from the others: n(2, gb) = N_gb(j) - n(0, gb) - n(1, gb)  (see tri_group2).
__host__ __device__ __forceinline__ int tri_marg_off(int nblocks, int k) { return (nblocks >> 2) * 36 + ((nblocks + 7) >> 3) * 4 + k * 4; }
__host__ __device__ __forceinline__ int tri_row_words(int nblocks) {
    int rw = (nblocks >> 2) * 36 + ((nblocks + 7) >> 3) * 4 + (nblocks >> 2) * 4;
    if ((rw / 4) % 2 == 0) rw += 4;
    return rw;
}

// logical word w (of fl.bw) of plane g of block b of one SNP, whatever the physical layout
__device__ __forceinline__ uint32_t logical_word(const uint32_t *__restrict__ planes, const FoldLayout &fl, int64_t snp_pad, int b,
                                                 int64_t snp, int g, int w) {
    if (fl.tri) {
        const uint32_t *row = planes + snp * fl.row_words;
        if (w < 3) return row[tri_word_off(b, g, w)];
        return (row[tri_tail_off(fl.nblocks, b, g)] >> tri_tail_shift(b)) & 0xFu;
    }
    return planes[plane_word(fl, snp_pad, b, snp, g, w)];
}

// nibble-wise popcount of x added to two packed byte counters: even nibbles -> lo, odd nibbles -> hi
__device__ __forceinline__ void tail_count_acc(uint32_t x, uint32_t &lo, uint32_t &hi) {
    x = x - ((x >> 1) & 0x55555555u);
    x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
    const uint32_t e = x & 0x0F0F0F0Fu;
    lo += e;
    hi += (x - e) >> 4;
}

// one block of the tri layout: 3 words -> 2 POPC, weighted into byte SHIFT/8 of the packed counter
template <uint32_t K>
__device__ __forceinline__ uint32_t tri_count2_acc(const uint32_t *a, const uint32_t *b, uint32_t acc) {
    const uint32_t x0 = a[0] & b[0], x1 = a[1] & b[1], x2 = a[2] & b[2];
    acc = mad_const<K>(__popc(xor3(x0, x1, x2)), acc);
    acc = mad_const<2 * K>(__popc(maj3(x0, x1, x2)), acc);
    return acc;
}

// ----------------------------------------------------------------------------
// High-risk rule -- bit-exact with mdr_high_risk_combinations2 (mdr.c:45-75).
//
// The reference evaluates, in float32 with round-to-nearest and no fusing,
//     r = (float)A/(float)U; pu = cu*r; rr = (ca+cu)/(pu+ca); nu = pu*rr; na = (ca+cu)-nu; risky = na >= nu
// In exact arithmetic this is ca*U >= cu*A (and cell non-empty).  Each of the
// six operations has relative error <= 2^-24, which can only flip the
// comparison when |ca*U - cu*A| <= ~6.7e-7 * cu*A; outside a 2^-19 band
// (2.8x margin) the integer sign decides, inside it the six float operations
// are replayed with explicit _rn intrinsics (never contracted to FMA).
// When A == U, r == 1 and every step is exact: risky <=> ca >= cu and ca > 0.
// Exhaustively cross-checked against the float32 sequence in tests/test_risk_rule.py.
// ----------------------------------------------------------------------------
__device__ __forceinline__ bool high_risk_float32(int ca, int cu, float ratio) {
    float fa = __int2float_rn(ca), fu = __int2float_rn(cu);
    float total = __fadd_rn(fa, fu);
    float pu = __fmul_rn(fu, ratio);
    float rr = __fdiv_rn(total, __fadd_rn(pu, fa));
    float nu = __fmul_rn(pu, rr);
    float na = __fsub_rn(total, nu);
    return na >= nu;   // NaN (empty cell) -> false
}
struct RiskParams {
    int balanced, A, U;
    float ratio;
};
__device__ __forceinline__ RiskParams risk_params(const FoldLayout &fl) {
    RiskParams rp;
    rp.balanced = fl.balanced; rp.A = fl.A; rp.U = fl.U; rp.ratio = fl.ratio;
    return rp;
}
__device__ __forceinline__ bool high_risk(int ca, int cu, const RiskParams &rp) {
    if (rp.balanced) return (ca >= cu) && (ca > 0);
    long long m = (long long) cu * rp.A;
    long long d = (long long) ca * rp.U - m;
    long long band = (m >> 19) + 1;
    if (d > band) return true;
    if (d < -band) return false;
    return high_risk_float32(ca, cu, rp.ratio);
}

// Balanced accuracy exactly as evaluate_model(BA), model.c:473, from
// {TP, FN, FP, TN} with FN = P - TP, TN = N - FP (model.c:445-453).
__device__ __forceinline__ double balanced_accuracy(int tp, int fp, int npos, int nneg) {
    double TP = (double) tp, FN = (double) (npos - tp), FP = (double) fp, TN = (double) (nneg - fp);
    double a = __ddiv_rn(TP, __dadd_rn(TP, FN));
    double b = __ddiv_rn(TN, __dadd_rn(TN, FP));
    return __dmul_rn(__dadd_rn(a, b), 0.5);   // /2 is exact
}

// The other evaluation functions of model.h:84 / model.c:462-479, from {TP, FN, FP, TN} in double, operation by operation
// like the reference's C (gcc, no FMA contraction: every product and sum is rounded on its own).  Codes = enum eval_function
// { CA, BA, wBA, GAMMA, TAU_B }.  CA (0) is what the reference EXECUTES for it: `if (!function) function = BA`
// (model.c:465-467) turns code 0 into BA; kEvalCATrue is the documented formula as an extension.  wBA is a TODO in the
// reference (no case in its switch) and is rejected on the host.
constexpr int kEvalCA = 0, kEvalBA = 1, kEvalWBA = 2, kEvalGamma = 3, kEvalTauB = 4, kEvalCATrue = 5;
__device__ __forceinline__ double evaluate_fn(int fn, int tp, int fn_, int fp, int tn) {
    const double TP = (double) tp, FN = (double) fn_, FP = (double) fp, TN = (double) tn;
    if (fn == kEvalGamma || fn == kEvalTauB) {
        const double cross = __dsub_rn(__dmul_rn(TP, TN), __dmul_rn(FP, FN));
        if (fn == kEvalGamma) return __ddiv_rn(cross, __dadd_rn(__dmul_rn(TP, TN), __dmul_rn(FP, FN)));
        const double prod = __dmul_rn(__dmul_rn(__dmul_rn(__dadd_rn(TP, FN), __dadd_rn(TN, FP)), __dadd_rn(TP, FP)), __dadd_rn(TN, FN));
        return __ddiv_rn(cross, __dsqrt_rn(prod));
    }
    if (fn == kEvalCATrue) return __ddiv_rn(__dadd_rn(TP, TN), __dadd_rn(__dadd_rn(__dadd_rn(TP, FN), TN), FP));
    const double a = __ddiv_rn(TP, __dadd_rn(TP, FN));
    const double b = __ddiv_rn(TN, __dadd_rn(TN, FP));
    return __dmul_rn(__dadd_rn(a, b), 0.5);
}
// order-preserving integer image of a (non-NaN) double: the thresholds of the search are integer scores
__device__ __forceinline__ long long value_score(double v) {
    if (isnan(v)) return LLONG_MIN;
    const long long b = __double_as_longlong(__dadd_rn(v, 0.0));     // -0.0 -> +0.0
    return b ^ ((b >> 63) & 0x7FFFFFFFFFFFFFFFLL);
}

// Integer score with the same ordering as the real-valued BA for a fixed fold:
// BA = (TP*N + TN*P) / (2PN)  =>  order by TP*N - FP*P.  Distinct scores differ
// by >= 1/(2PN) in BA, far above double rounding, so score order == BA order;
// equal scores are resolved with the double BA itself.
__device__ __forceinline__ long long ba_score(int tp, int fp, int npos, int nneg) {
    return (long long) tp * nneg - (long long) fp * npos;
}

// canonical order: BA descending, then SNP tuple ascending
__device__ __forceinline__ bool cand_before(double ba_a, int ia, int ja, int ka, double ba_b, int ib, int jb, int kb) {
    if (ba_a != ba_b) return ba_a > ba_b;
    if (ia != ib) return ia < ib;
    if (ja != jb) return ja < jb;
    return ka < kb;
}
__device__ __forceinline__ bool cand_before(const Cand &a, const Cand &b) {
    return cand_before(a.ba, a.i, a.j, a.k, b.ba, b.i, b.j, b.k);
}

// List entries are shared by the warps of one CTA and live either in shared or in
// global memory: volatile generic accesses never hit a stale L1 line.
__device__ __forceinline__ Cand cand_load(const Cand *p) {
    int4 a, b;
    asm volatile("ld.volatile.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
    asm volatile("ld.volatile.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(reinterpret_cast<const int4 *>(p) + 1));
    Cand c;
    c.ba = __hiloint2double(a.y, a.x);
    c.i = a.z; c.j = a.w; c.k = b.x; c.mask = (uint32_t) b.y; c.tp = b.z; c.fp = b.w;
    return c;
}
__device__ __forceinline__ void cand_store(Cand *p, const Cand &c) {
    int4 a, b;
    a.x = __double2loint(c.ba); a.y = __double2hiint(c.ba); a.z = c.i; a.w = c.j;
    b.x = c.k; b.y = (int) c.mask; b.z = c.tp; b.w = c.fp;
    asm volatile("st.volatile.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
    asm volatile("st.volatile.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<int4 *>(p) + 1), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// linear index of the pair (i, j), i < j, in lexicographic order over n SNPs
__host__ __device__ __forceinline__ uint64_t pair_index(uint64_t n, uint64_t i, uint64_t j) {
    return i * (2 * n - i - 1) / 2 + (j - i - 1);
}
__host__ __device__ __forceinline__ uint64_t choose2(uint64_t m) { return m * (m - 1) / 2; }
__host__ __device__ __forceinline__ uint64_t choose3(uint64_t m) {
    // m(m-1)(m-2)/6 without overflow for m up to ~3.3e6
    uint64_t a = m, b = m - 1, c = m - 2;
    if (m < 3) return 0;
    if (a % 2 == 0) a /= 2; else b /= 2;
    if (a % 3 == 0) a /= 3; else if (b % 3 == 0) b /= 3; else c /= 3;
    return a * b * c;
}
// linear index of the triple (i, j, k), i < j < k
__host__ __device__ __forceinline__ uint64_t triple_index(uint64_t n, uint64_t i, uint64_t j, uint64_t k) {
    return (choose3(n) - choose3(n - i)) + (choose2(n - i - 1) - choose2(n - j)) + (k - j - 1);
}

}  // namespace hpgv
