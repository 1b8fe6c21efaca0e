// Montgomery-product variants that were MEASURED AND REJECTED in round 1: kept here as
// microbenchmark evidence, not part of the product header csrc/fr.cuh.
//   fr_mul2           two interleaved products per thread: ptxas keeps the chains apart, -10 % in K2
//   fr_mul_karatsuba  48 + 72 instead of 64 + 72 multiplications, 410 instead of 245 instructions: 50.4 vs 66.9 Gprod/s
//   fr_mul_u64        (round 2) the north star's formulation: 4 x 64-bit limbs, CIOS with mad.lo.cc.u64 / madc.hi.cc.u64
//                     chains -- ptxas lowers every 64-bit multiply-add to four 32-bit IMAD.WIDE plus carry fix-ups;
//                     measured side by side with the shipped 8 x 32-bit split-accumulator form
// (The third variant, BN254's p0 = 2^32 - 2^28 + 1 by shifts, was a four-line branch of mont_round: one multiplication
// less and six more ALU operations per round, 65.2 vs 66.9 Gprod/s; removed from the header.)
#pragma once
#include "fr.cuh"

namespace acg {

// Two independent Montgomery products with their rounds alternating in program order: every carry chain is
// self-contained (it starts without carry-in and ends without carry-out), so the interleaving is legal at the
// PTX level, and ptxas -- which renames the carry flag into predicates -- overlaps the two dependency chains.
// For latency-bound callers that have two products to do (the second one costs much less than a full product).
template <class P>
ACG_HD void fr_mul2(fr_t& r0, fr_t& r1, const fr_t& a0, const fr_t& b0, const fr_t& a1, const fr_t& b1) {
    uint32_t e0[9], o0[8], e1[9], o1[8], f0[9], g0[8], f1[9], g1[8];
    mont_round<P, true>(e0, o0, e0, o0, a0.l, b0.l[0]);
    mont_round<P, true>(f0, g0, f0, g0, a1.l, b1.l[0]);
    mont_round<P, false>(e1, o1, e0, o0, a0.l, b0.l[1]);
    mont_round<P, false>(f1, g1, f0, g0, a1.l, b1.l[1]);
    mont_round<P, false>(e0, o0, e1, o1, a0.l, b0.l[2]);
    mont_round<P, false>(f0, g0, f1, g1, a1.l, b1.l[2]);
    mont_round<P, false>(e1, o1, e0, o0, a0.l, b0.l[3]);
    mont_round<P, false>(f1, g1, f0, g0, a1.l, b1.l[3]);
    mont_round<P, false>(e0, o0, e1, o1, a0.l, b0.l[4]);
    mont_round<P, false>(f0, g0, f1, g1, a1.l, b1.l[4]);
    mont_round<P, false>(e1, o1, e0, o0, a0.l, b0.l[5]);
    mont_round<P, false>(f1, g1, f0, g0, a1.l, b1.l[5]);
    mont_round<P, false>(e0, o0, e1, o1, a0.l, b0.l[6]);
    mont_round<P, false>(f0, g0, f1, g1, a1.l, b1.l[6]);
    mont_round<P, false>(e1, o1, e0, o0, a0.l, b0.l[7]);
    mont_round<P, false>(f1, g1, f0, g0, a1.l, b1.l[7]);
    uint32_t t[8], u[8];
    t[0] = ptx::add_cc(o1[0], e1[1]);
#pragma unroll
    for (int i = 1; i < 7; ++i) t[i] = ptx::addc_cc(o1[i], e1[i + 1]);
    t[7] = ptx::addc(o1[7], e1[8]);
    u[0] = ptx::add_cc(g1[0], f1[1]);
#pragma unroll
    for (int i = 1; i < 7; ++i) u[i] = ptx::addc_cc(g1[i], f1[i + 1]);
    u[7] = ptx::addc(g1[7], f1[8]);
    r0 = fr_reduce_once<P>(t);
    r1 = fr_reduce_once<P>(u);
}
// ------------------------------------------------------------------------------------------------
// Karatsuba variant (experiment): the a*b half of the product with 48 instead of 64
// 32x32->64 multiplications, then a separate Montgomery reduction (72).  Bit-exact (tests/test_fr_host.py, op 9)
// and measured SLOWER on B200: 50.4 vs 66.9 G products/s -- 113 instead of 160 multiplier-pipe instructions per
// product in SASS, but 410 instead of 245 instructions in total, and it is the issue rate of carry-chain
// instructions that bounds the product, not the multiplier count.  Not used by any kernel; kept for the microbenchmark
// (tools/microbench/fr_mul_throughput.cu) so the measurement can be repeated.
// ------------------------------------------------------------------------------------------------
// r[0..7] = a[0..3] * b[0..3].  Split accumulators as in mont_round: E holds the limbs at positions 0..7, O the limbs
// at positions 1..8, so that every product lands on a 64-bit-aligned pair of its array.
ACG_HD void mul_4x4(uint32_t r[8], const uint32_t a[4], const uint32_t b[4]) {
    uint32_t E[8], O[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) E[i] = O[i] = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t bi = b[i];
        if ((i & 1) == 0) {
            // a0, a2 -> positions i, i+2 (even): E[i..i+3];  a1, a3 -> positions i+1, i+3 (odd): O[i..i+3]
            E[i] = ptx::mad_lo_cc(a[0], bi, E[i]);
            E[i + 1] = ptx::madc_hi_cc(a[0], bi, E[i + 1]);
            E[i + 2] = ptx::madc_lo_cc(a[2], bi, E[i + 2]);
            E[i + 3] = ptx::madc_hi_cc(a[2], bi, E[i + 3]);
#pragma unroll
            for (int k = i + 4; k < 8; ++k) E[k] = ptx::addc_cc(E[k], 0u);
            O[i] = ptx::mad_lo_cc(a[1], bi, O[i]);
            O[i + 1] = ptx::madc_hi_cc(a[1], bi, O[i + 1]);
            O[i + 2] = ptx::madc_lo_cc(a[3], bi, O[i + 2]);
            O[i + 3] = ptx::madc_hi_cc(a[3], bi, O[i + 3]);
#pragma unroll
            for (int k = i + 4; k < 8; ++k) O[k] = ptx::addc_cc(O[k], 0u);
        } else {
            // a0, a2 -> positions i, i+2 (odd): O[i-1..i+2];  a1, a3 -> positions i+1, i+3 (even): E[i+1..i+4]
            O[i - 1] = ptx::mad_lo_cc(a[0], bi, O[i - 1]);
            O[i] = ptx::madc_hi_cc(a[0], bi, O[i]);
            O[i + 1] = ptx::madc_lo_cc(a[2], bi, O[i + 1]);
            O[i + 2] = ptx::madc_hi_cc(a[2], bi, O[i + 2]);
#pragma unroll
            for (int k = i + 3; k < 8; ++k) O[k] = ptx::addc_cc(O[k], 0u);
            E[i + 1] = ptx::mad_lo_cc(a[1], bi, E[i + 1]);
            E[i + 2] = ptx::madc_hi_cc(a[1], bi, E[i + 2]);
            E[i + 3] = ptx::madc_lo_cc(a[3], bi, E[i + 3]);
            E[i + 4] = ptx::madc_hi_cc(a[3], bi, E[i + 4]);
#pragma unroll
            for (int k = i + 5; k < 8; ++k) E[k] = ptx::addc_cc(E[k], 0u);
        }
    }
    // r = E + (O << 32); the product is < 2^256, so nothing is carried out of limb 7 (and O[7] == 0)
    r[0] = E[0];
    r[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 8; ++k) r[k] = ptx::addc_cc(E[k], O[k - 1]);
}

// One round of a Montgomery reduction on the split accumulator of mont_round (same representation: the pair holds
// S = T * 2^32 with a zero low limb, T position k = ev[k + 1] + od[k]):
//   T <- (T + m*p) / 2^32,  m = -T * p^-1 mod 2^32        -- mont_round without the a*bi products.  T < 2^256 stays so.
template <class P>
ACG_HD void mont_reduce_round(uint32_t nev[9], uint32_t nod[8], const uint32_t ev[9], const uint32_t od[8]) {
    nev[0] = ptx::add_cc(od[0], ev[1]);  // carry has weight 2^32: consumed by the od chain
    nod[0] = ptx::addc_cc(ev[2], 0u);
    nod[1] = ptx::addc_cc(ev[3], 0u);
    nod[2] = ptx::addc_cc(ev[4], 0u);
    nod[3] = ptx::addc_cc(ev[5], 0u);
    nod[4] = ptx::addc_cc(ev[6], 0u);
    nod[5] = ptx::addc_cc(ev[7], 0u);
    nod[6] = ptx::addc_cc(ev[8], 0u);
    nod[7] = ptx::addc(0u, 0u);
#pragma unroll
    for (int i = 1; i < 8; ++i) nev[i] = od[i];
    nev[8] = 0u;
    const uint32_t m = nev[0] * P::NINV32;
    nod[0] = ptx::mad_lo_cc(P::p(1), m, nod[0]);
    nod[1] = ptx::madc_hi_cc(P::p(1), m, nod[1]);
    nod[2] = ptx::madc_lo_cc(P::p(3), m, nod[2]);
    nod[3] = ptx::madc_hi_cc(P::p(3), m, nod[3]);
    nod[4] = ptx::madc_lo_cc(P::p(5), m, nod[4]);
    nod[5] = ptx::madc_hi_cc(P::p(5), m, nod[5]);
    nod[6] = ptx::madc_lo_cc(P::p(7), m, nod[6]);
    nod[7] = ptx::madc_hi(P::p(7), m, nod[7]);
    nev[0] = ptx::mad_lo_cc(P::p(0), m, nev[0]);  // == 0
    nev[1] = ptx::madc_hi_cc(P::p(0), m, nev[1]);
    nev[2] = ptx::madc_lo_cc(P::p(2), m, nev[2]);
    nev[3] = ptx::madc_hi_cc(P::p(2), m, nev[3]);
    nev[4] = ptx::madc_lo_cc(P::p(4), m, nev[4]);
    nev[5] = ptx::madc_hi_cc(P::p(4), m, nev[5]);
    nev[6] = ptx::madc_lo_cc(P::p(6), m, nev[6]);
    nev[7] = ptx::madc_hi_cc(P::p(6), m, nev[7]);
    nev[8] = ptx::addc(nev[8], 0u);
}

// Same contract as fr_mul (a <= p, b any 256-bit value; result in [0, p)).
template <class P>
ACG_HD fr_t fr_mul_karatsuba(const fr_t& a, const fr_t& b) {
    uint32_t z0[8], z2[8], zm[9], sa[4], sb[4];
    mul_4x4(z0, a.l, b.l);
    mul_4x4(z2, a.l + 4, b.l + 4);
    // sa = a_lo + a_hi, sb = b_lo + b_hi (4 limbs + carry bit each)
    sa[0] = ptx::add_cc(a.l[0], a.l[4]);
    sa[1] = ptx::addc_cc(a.l[1], a.l[5]);
    sa[2] = ptx::addc_cc(a.l[2], a.l[6]);
    sa[3] = ptx::addc_cc(a.l[3], a.l[7]);
    const uint32_t ca = ptx::addc(0u, 0u);
    sb[0] = ptx::add_cc(b.l[0], b.l[4]);
    sb[1] = ptx::addc_cc(b.l[1], b.l[5]);
    sb[2] = ptx::addc_cc(b.l[2], b.l[6]);
    sb[3] = ptx::addc_cc(b.l[3], b.l[7]);
    const uint32_t cb = ptx::addc(0u, 0u);
    // zm = (ca*2^128 + sa)(cb*2^128 + sb) = sa*sb + (ca ? sb : 0)*2^128 + (cb ? sa : 0)*2^128 + ca*cb*2^256
    mul_4x4(zm, sa, sb);
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    zm[4] = ptx::add_cc(zm[4], sb[0] & ma);
    zm[5] = ptx::addc_cc(zm[5], sb[1] & ma);
    zm[6] = ptx::addc_cc(zm[6], sb[2] & ma);
    zm[7] = ptx::addc_cc(zm[7], sb[3] & ma);
    zm[8] = ptx::addc(ca & cb, 0u);
    zm[4] = ptx::add_cc(zm[4], sa[0] & mb);
    zm[5] = ptx::addc_cc(zm[5], sa[1] & mb);
    zm[6] = ptx::addc_cc(zm[6], sa[2] & mb);
    zm[7] = ptx::addc_cc(zm[7], sa[3] & mb);
    zm[8] = ptx::addc(zm[8], 0u);
    // z1 = zm - z0 - z2  (>= 0, < 2^258)
    zm[0] = ptx::sub_cc(zm[0], z0[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) zm[i] = ptx::subc_cc(zm[i], z0[i]);
    zm[8] = ptx::subc(zm[8], 0u);
    zm[0] = ptx::sub_cc(zm[0], z2[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) zm[i] = ptx::subc_cc(zm[i], z2[i]);
    zm[8] = ptx::subc(zm[8], 0u);
    // Z = z0 + z1 * 2^128 + z2 * 2^256  (16 limbs: low half in z0, high half in z2)
    z0[4] = ptx::add_cc(z0[4], zm[0]);
    z0[5] = ptx::addc_cc(z0[5], zm[1]);
    z0[6] = ptx::addc_cc(z0[6], zm[2]);
    z0[7] = ptx::addc_cc(z0[7], zm[3]);
    z2[0] = ptx::addc_cc(z2[0], zm[4]);
    z2[1] = ptx::addc_cc(z2[1], zm[5]);
    z2[2] = ptx::addc_cc(z2[2], zm[6]);
    z2[3] = ptx::addc_cc(z2[3], zm[7]);
    z2[4] = ptx::addc_cc(z2[4], zm[8]);
    z2[5] = ptx::addc_cc(z2[5], 0u);
    z2[6] = ptx::addc_cc(z2[6], 0u);
    z2[7] = ptx::addc(z2[7], 0u);
    // (Z + M*p) / 2^256 = Z_hi + (Z_lo + M*p) / 2^256: eight reduction rounds on the low half alone (the window stays
    // below 2^256), then the high half is added.  Z_hi < p and the reduced low half is <= p: the sum is < 2p.
    uint32_t e0[9], o0[8], e1[9], o1[8];
    e0[0] = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        e0[i + 1] = z0[i];
        o0[i] = 0u;
    }
    mont_reduce_round<P>(e1, o1, e0, o0);
    mont_reduce_round<P>(e0, o0, e1, o1);
    mont_reduce_round<P>(e1, o1, e0, o0);
    mont_reduce_round<P>(e0, o0, e1, o1);
    mont_reduce_round<P>(e1, o1, e0, o0);
    mont_reduce_round<P>(e0, o0, e1, o1);
    mont_reduce_round<P>(e1, o1, e0, o0);
    mont_reduce_round<P>(e0, o0, e1, o1);
    uint32_t t[8];
    t[0] = ptx::add_cc(o0[0], e0[1]);
#pragma unroll
    for (int i = 1; i < 7; ++i) t[i] = ptx::addc_cc(o0[i], e0[i + 1]);
    t[7] = ptx::addc(o0[7], e0[8]);
    t[0] = ptx::add_cc(t[0], z2[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) t[i] = ptx::addc_cc(t[i], z2[i]);
    t[7] = ptx::addc(t[7], z2[7]);
    return fr_reduce_once<P>(t);
}

#if defined(__CUDACC__)
// 4 x 64-bit-limb CIOS Montgomery product (device only): a, b < 2^256 with a * b < p * 2^256; result in [0, p).
template <class P>
__device__ __forceinline__ fr_t fr_mul_u64(const fr_t& a, const fr_t& b) {
    uint64_t A[4], B[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        A[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
        B[i] = (uint64_t)b.l[2 * i] | ((uint64_t)b.l[2 * i + 1] << 32);
    }
    constexpr uint64_t p0 = (uint64_t)P::p(0) | ((uint64_t)P::p(1) << 32), p1 = (uint64_t)P::p(2) | ((uint64_t)P::p(3) << 32),
                       p2 = (uint64_t)P::p(4) | ((uint64_t)P::p(5) << 32), p3 = (uint64_t)P::p(6) | ((uint64_t)P::p(7) << 32);
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint64_t t5, m, z;
        // t += A * B[i]
        asm("mad.lo.cc.u64 %0, %6, %10, %0;\n\t"
            "madc.lo.cc.u64 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u64 %2, %8, %10, %2;\n\t"
            "madc.lo.cc.u64 %3, %9, %10, %3;\n\t"
            "addc.cc.u64 %4, %4, 0;\n\t"
            "addc.u64 %5, 0, 0;\n\t"
            "mad.hi.cc.u64 %1, %6, %10, %1;\n\t"
            "madc.hi.cc.u64 %2, %7, %10, %2;\n\t"
            "madc.hi.cc.u64 %3, %8, %10, %3;\n\t"
            "madc.hi.cc.u64 %4, %9, %10, %4;\n\t"
            "addc.u64 %5, %5, 0;"
            : "+l"(t0), "+l"(t1), "+l"(t2), "+l"(t3), "+l"(t4), "=l"(t5)
            : "l"(A[0]), "l"(A[1]), "l"(A[2]), "l"(A[3]), "l"(B[i]));
        m = t0 * P::NINV64;
        // t = (t + m * p) >> 64
        asm("mad.lo.cc.u64 %5, %6, %7, %0;\n\t"
            "madc.lo.cc.u64 %0, %6, %8, %1;\n\t"
            "madc.lo.cc.u64 %1, %6, %9, %2;\n\t"
            "madc.lo.cc.u64 %2, %6, %10, %3;\n\t"
            "addc.cc.u64 %3, %4, 0;\n\t"
            "addc.u64 %4, %11, 0;\n\t"
            "mad.hi.cc.u64 %0, %6, %7, %0;\n\t"
            "madc.hi.cc.u64 %1, %6, %8, %1;\n\t"
            "madc.hi.cc.u64 %2, %6, %9, %2;\n\t"
            "madc.hi.cc.u64 %3, %6, %10, %3;\n\t"
            "addc.u64 %4, %4, 0;"
            : "+l"(t0), "+l"(t1), "+l"(t2), "+l"(t3), "+l"(t4), "=l"(z)
            : "l"(m), "l"(p0), "l"(p1), "l"(p2), "l"(p3), "l"(t5));
    }
    // one conditional subtraction of p
    uint64_t s0, s1, s2, s3, brw;
    asm("sub.cc.u64 %0, %5, %9;\n\t"
        "subc.cc.u64 %1, %6, %10;\n\t"
        "subc.cc.u64 %2, %7, %11;\n\t"
        "subc.cc.u64 %3, %8, %12;\n\t"
        "subc.u64 %4, 0, 0;"
        : "=l"(s0), "=l"(s1), "=l"(s2), "=l"(s3), "=l"(brw)
        : "l"(t0), "l"(t1), "l"(t2), "l"(t3), "l"(p0), "l"(p1), "l"(p2), "l"(p3));
    const bool keep = brw != 0 && t4 == 0;  // t < p
    const uint64_t r[4] = {keep ? t0 : s0, keep ? t1 : s1, keep ? t2 : s2, keep ? t3 : s3};
    fr_t out;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        out.l[2 * i] = (uint32_t)r[i];
        out.l[2 * i + 1] = (uint32_t)(r[i] >> 32);
    }
    return out;
}
#endif

}  // namespace acg
