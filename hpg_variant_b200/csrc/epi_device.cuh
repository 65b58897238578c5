// epi_device.cuh -- device helpers: mbarrier / bulk-copy PTX, LOP3 carry-save
// counting, the exact high-risk rule, balanced accuracy, candidate ordering.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion signalled on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ----------------------------------------------------------------------------
// bit counting: AND the planes, compress with carry-save adders (LOP3), POPC
// ----------------------------------------------------------------------------
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

// number of set bits in x[0..BW): 4 words -> CSA(3)+1 (3 POPC), 8 words -> CSA tree of 7 + 1 (4 POPC)
template <int BW>
__device__ __forceinline__ uint32_t count_words(const uint32_t (&x)[BW]) {
    if constexpr (BW == 4) {
        uint32_t ones = xor3(x[0], x[1], x[2]);
        uint32_t twos = maj3(x[0], x[1], x[2]);
        return __popc(ones) + __popc(x[3]) + 2u * __popc(twos);
    } else {
        static_assert(BW == 8, "block width must be 4 or 8 words");
        uint32_t s1 = xor3(x[0], x[1], x[2]), c1 = maj3(x[0], x[1], x[2]);
        uint32_t s2 = xor3(x[3], x[4], x[5]), c2 = maj3(x[3], x[4], x[5]);
        uint32_t s3 = xor3(s1, s2, x[6]),     c3 = maj3(s1, s2, x[6]);
        uint32_t t  = xor3(c1, c2, c3),       f  = maj3(c1, c2, c3);
        return __popc(s3) + __popc(x[7]) + 2u * __popc(t) + 4u * __popc(f);
    }
}

template <int BW>
__device__ __forceinline__ uint32_t cell_count2(const uint32_t (&a)[BW], const uint32_t (&b)[BW]) {
    uint32_t x[BW];
#pragma unroll
    for (int w = 0; w < BW; w++) x[w] = a[w] & b[w];
    return count_words<BW>(x);
}
template <int BW>
__device__ __forceinline__ uint32_t cell_count3(const uint32_t (&a)[BW], const uint32_t (&b)[BW], const uint32_t (&c)[BW]) {
    uint32_t x[BW];
#pragma unroll
    for (int w = 0; w < BW; w++) x[w] = and3(a[w], b[w], c[w]);
    return count_words<BW>(x);
}

// One plane (BW words) of a staged row.  For BW == 8 the two 16-byte halves of a
// plane are stored swapped when bit 2 of the SNP index is set, which makes the
// per-lane LDS.128 of 32 consecutive rows (96-byte stride) bank-conflict free.
template <int BW>
__device__ __forceinline__ void load_plane(const uint32_t *row, int g, int swz, uint32_t (&p)[BW]) {
    if constexpr (BW == 4) {
        uint4 v = *reinterpret_cast<const uint4 *>(row + g * 4);
        p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    } else {
        uint4 lo = *reinterpret_cast<const uint4 *>(row + g * 8 + (swz ? 4 : 0));
        uint4 hi = *reinterpret_cast<const uint4 *>(row + g * 8 + (swz ? 0 : 4));
        p[0] = lo.x; p[1] = lo.y; p[2] = lo.z; p[3] = lo.w;
        p[4] = hi.x; p[5] = hi.y; p[6] = hi.z; p[7] = hi.w;
    }
}
__host__ __device__ __forceinline__ int swizzle_of(int64_t snp) { return static_cast<int>((snp >> 2) & 1); }

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

// list entries live in global memory and are shared by the warps of one CTA:
// always go through L2 (ld.cg / st.cg) so no warp sees a stale L1 line.
__device__ __forceinline__ Cand cand_load(const Cand *p) {
    const int4 *q = reinterpret_cast<const int4 *>(p);
    int4 a = __ldcg(q), b = __ldcg(q + 1);
    Cand c;
    c.ba = __hiloint2double(a.y, a.x);
    c.i = a.z; c.j = a.w; c.k = b.x; c.mask = (uint32_t) b.y; c.tp = b.z; c.fp = b.w;
    return c;
}
__device__ __forceinline__ void cand_store(Cand *p, const Cand &c) {
    int4 a, b;
    a.x = __double2loint(c.ba); a.y = __double2hiint(c.ba); a.z = c.i; a.w = c.j;
    b.x = c.k; b.y = (int) c.mask; b.z = c.tp; b.w = c.fp;
    int4 *q = reinterpret_cast<int4 *>(p);
    __stcg(q, a);
    __stcg(q + 1, b);
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
