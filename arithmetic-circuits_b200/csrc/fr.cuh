// K1 -- prime-field arithmetic for the R1CS / QAP path: 256-bit Montgomery Fr over 8x32-bit limbs.
//
// Replaces `Prime r` arithmetic of galois-field-1.0.2 (+ - * negate recip pow fromP), as instantiated
// by the reference at bench/Circuit.hs:10, test/Test/QAP.hs:12 (BN254 Fr) -- SURVEY.md section 8a, R1.
// Elements live in HBM as 32-byte records (8 LE uint32 limbs == 4 LE uint64 limbs), in Montgomery
// form (x*2^256 mod r) on the device and canonical at the C ABI.
//
// Every routine is __host__ __device__.  On the device the carry chains are inline PTX
// (add.cc / addc / mad.lo.cc / madc.hi.cc -- ptxas fuses each lo/hi pair into one IMAD.WIDE.U32 with a
// predicate carry); on the host the same instruction sequence runs against an emulated carry flag, so
// the exact device algorithm is unit-tested on a CPU-only box (tests/test_fr_host.py).
//
// Both supported moduli are < 2^255, which the multiplier relies on (9-limb accumulator, no 10th limb).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ACG_HD __host__ __device__ __forceinline__
#else
#define ACG_HD inline
#endif

// BN254 Fr's p0 = 2^32 - 2^28 + 1 lets m * p0 be built from shifts (see mont_round).  Measured on B200 (round 1, twice):
// one IMAD.WIDE less per round, but SLOWER everywhere -- K1 ceiling 65.2 vs 66.9 Gproducts/s, tiled check 61.1 vs
// 59.6 us, 2^22-point NTT 1.03 vs 0.96 ms: the six dependent ALU operations sit on the cross-round critical path.
// Kept behind this switch (off) so the measurement can be repeated.
#ifndef ACG_BN254_P0_SHIFTS
#define ACG_BN254_P0_SHIFTS 0
#endif

namespace acg {

#include "fr_constants.inc"

// ------------------------------------------------------------------------------------------------
// carry-chain primitives
// ------------------------------------------------------------------------------------------------
namespace ptx {
#if !defined(__CUDA_ARCH__)
inline uint32_t& host_cf() {
    static thread_local uint32_t f = 0;
    return f;
}
#endif

ACG_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t s = (uint64_t)a + b;
    host_cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ACG_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t s = (uint64_t)a + b + host_cf();
    host_cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ACG_HD uint32_t addc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    return a + b + host_cf();
#endif
}
ACG_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    host_cf() = a < b ? 1u : 0u;   // borrow
    return a - b;
#endif
}
ACG_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t d = (uint64_t)a - b - host_cf();
    host_cf() = (uint32_t)(d >> 63);
    return (uint32_t)d;
#endif
}
ACG_HD uint32_t subc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    return a - b - host_cf();
#endif
}
ACG_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c;
    host_cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ACG_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + host_cf();
    host_cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ACG_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t s = (((uint64_t)a * b) >> 32) + c + host_cf();
    host_cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
#endif
}
ACG_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return (uint32_t)(((uint64_t)a * b) >> 32) + c + host_cf();
#endif
}
}  // namespace ptx

// ------------------------------------------------------------------------------------------------
// element type
// ------------------------------------------------------------------------------------------------
struct alignas(32) fr_t {
    uint32_t l[8];
};

template <class P>
ACG_HD fr_t fr_zero() {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = 0;
    return r;
}
template <class P>
ACG_HD fr_t fr_one() {  // Montgomery 1
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = P::one(i);
    return r;
}
ACG_HD bool fr_is_zero(const fr_t& a) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc |= a.l[i];
    return acc == 0;
}
ACG_HD bool fr_eq(const fr_t& a, const fr_t& b) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc |= a.l[i] ^ b.l[i];
    return acc == 0;
}
template <class P>
ACG_HD bool fr_is_one(const fr_t& a) {  // == Montgomery 1
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc |= a.l[i] ^ P::one(i);
    return acc == 0;
}
template <class P>
ACG_HD bool fr_is_minus_one(const fr_t& a) {  // == Montgomery -1
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc |= a.l[i] ^ P::minus_one(i);
    return acc == 0;
}
// a < p ?  (canonical-range check used at the ABI)
template <class P>
ACG_HD bool fr_is_canonical(const fr_t& a) {
    uint32_t t = ptx::sub_cc(a.l[0], P::p(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) t = ptx::subc_cc(a.l[i], P::p(i));
    uint32_t borrow = ptx::subc(0u, 0u);  // 0xffffffff when a < p
    (void)t;
    return borrow != 0;
}

// r = a + b mod p      (a, b < p; for a, b <= p the result is in [0, p])
template <class P>
ACG_HD fr_t fr_add(const fr_t& a, const fr_t& b) {
    fr_t t, s;
    t.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) t.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    t.l[7] = ptx::addc(a.l[7], b.l[7]);  // p < 2^255: no carry out of 256 bits
    s.l[0] = ptx::sub_cc(t.l[0], P::p(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) s.l[i] = ptx::subc_cc(t.l[i], P::p(i));
    uint32_t borrow = ptx::subc(0u, 0u);
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = borrow ? t.l[i] : s.l[i];
    return r;
}
// r = a - b mod p
template <class P>
ACG_HD fr_t fr_sub(const fr_t& a, const fr_t& b) {
    fr_t t;
    t.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) t.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t mask = ptx::subc(0u, 0u);  // all ones when a < b
    fr_t r;
    r.l[0] = ptx::add_cc(t.l[0], P::p(0) & mask);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.l[i] = ptx::addc_cc(t.l[i], P::p(i) & mask);
    r.l[7] = ptx::addc(t.l[7], P::p(7) & mask);
    return r;
}
template <class P>
ACG_HD fr_t fr_neg(const fr_t& a) {
    return fr_sub<P>(fr_zero<P>(), a);
}

// One CIOS round on the split accumulator  T = ev + od * 2^32:
//   T <- (T + a*bi + m*p) / 2^32,  m = -T * p^-1 mod 2^32.
// `ev` holds the even-aligned 64-bit columns (ev[0..8]), `od` the odd-aligned ones (od[0..7], weight
// 2^32).  Products of even limbs of a (or p) land on ev, of odd limbs on od, so every multiply-add
// is a 64-bit-aligned IMAD.WIDE with a predicate carry.  The division by 2^32 is free: the
// accumulators swap roles (new ev = od + ev[1], new od = ev >> 64), which keeps the 64-bit register
// pairing intact.  Outputs go to (nev, nod).  Bound: T < 2p before, < 2p*2^32 < 2^288 inside.
template <class P, bool FIRST>
ACG_HD void mont_round(uint32_t nev[9], uint32_t nod[8], const uint32_t ev[9], const uint32_t od[8],
                       const uint32_t a[8], uint32_t bi) {
    if (FIRST) {  // ev = od = 0
        nod[0] = ptx::mad_lo_cc(a[1], bi, 0u);
    } else {
        nev[0] = ptx::add_cc(od[0], ev[1]);        // carry has weight 2^32: consumed by the od chain
        nod[0] = ptx::madc_lo_cc(a[1], bi, ev[2]);
    }
    nod[1] = ptx::madc_hi_cc(a[1], bi, FIRST ? 0u : ev[3]);
    nod[2] = ptx::madc_lo_cc(a[3], bi, FIRST ? 0u : ev[4]);
    nod[3] = ptx::madc_hi_cc(a[3], bi, FIRST ? 0u : ev[5]);
    nod[4] = ptx::madc_lo_cc(a[5], bi, FIRST ? 0u : ev[6]);
    nod[5] = ptx::madc_hi_cc(a[5], bi, FIRST ? 0u : ev[7]);
    nod[6] = ptx::madc_lo_cc(a[7], bi, FIRST ? 0u : ev[8]);
    nod[7] = ptx::madc_hi(a[7], bi, 0u);

    nev[0] = ptx::mad_lo_cc(a[0], bi, FIRST ? 0u : nev[0]);
    nev[1] = ptx::madc_hi_cc(a[0], bi, FIRST ? 0u : od[1]);
    nev[2] = ptx::madc_lo_cc(a[2], bi, FIRST ? 0u : od[2]);
    nev[3] = ptx::madc_hi_cc(a[2], bi, FIRST ? 0u : od[3]);
    nev[4] = ptx::madc_lo_cc(a[4], bi, FIRST ? 0u : od[4]);
    nev[5] = ptx::madc_hi_cc(a[4], bi, FIRST ? 0u : od[5]);
    nev[6] = ptx::madc_lo_cc(a[6], bi, FIRST ? 0u : od[6]);
    nev[7] = ptx::madc_hi_cc(a[6], bi, FIRST ? 0u : od[7]);
    nev[8] = ptx::addc(0u, 0u);

    // m = -T * p^-1 mod 2^32.  BLS12-381 Fr has p0 = 1, p1 = 2^32 - 1 and -p^-1 = -1 (mod 2^32): m and the
    // products m*p0, m*p1 then come from adds on the ALU pipe instead of the quarter-rate 32x32->64 multiplier
    // (measured +4% on the tiled check).  The analogous rewrite for BN254 Fr (p0 = 2^32 - 2^28 + 1) is slower:
    // see ACG_BN254_P0_SHIFTS at the top of this file.
    uint32_t m;
    if constexpr (P::NINV32 == 0xffffffffu)
        m = 0u - nev[0];
    else
        m = nev[0] * P::NINV32;

    if constexpr (P::p(1) == 0xffffffffu) {  // m * (2^32 - 1) = (m << 32) - m
        const uint32_t lo = ptx::sub_cc(0u, m);
        const uint32_t hi = ptx::subc(m, 0u);
        nod[0] = ptx::add_cc(nod[0], lo);
        nod[1] = ptx::addc_cc(nod[1], hi);
    } else {
        nod[0] = ptx::mad_lo_cc(P::p(1), m, nod[0]);
        nod[1] = ptx::madc_hi_cc(P::p(1), m, nod[1]);
    }
    nod[2] = ptx::madc_lo_cc(P::p(3), m, nod[2]);
    nod[3] = ptx::madc_hi_cc(P::p(3), m, nod[3]);
    nod[4] = ptx::madc_lo_cc(P::p(5), m, nod[4]);
    nod[5] = ptx::madc_hi_cc(P::p(5), m, nod[5]);
    nod[6] = ptx::madc_lo_cc(P::p(7), m, nod[6]);
    nod[7] = ptx::madc_hi(P::p(7), m, nod[7]);

    if constexpr (P::p(0) == 1u) {  // m * 1: the low word cancels nev[0] (carry iff nev[0] != 0)
        nev[0] = ptx::add_cc(nev[0], m);
        nev[1] = ptx::addc_cc(nev[1], 0u);
#if ACG_BN254_P0_SHIFTS
    } else if constexpr (P::p(0) == 0xf0000001u) {
        // BN254 Fr: p0 = 2^32 - 2^28 + 1, so m * p0 = (m << 32) - (m << 28) + m needs no multiplier:
        //   low word  = m - (m << 28)            (mod 2^32, borrow b)
        //   high word = m - (m >> 4) - b
        // One quarter-rate IMAD.WIDE less per round (8 of 136 per product) for six ALU operations.
        const uint32_t lo = ptx::sub_cc(m, m << 28);
        const uint32_t hi = ptx::subc(m, m >> 4);
        nev[0] = ptx::add_cc(nev[0], lo);   // == 0
        nev[1] = ptx::addc_cc(nev[1], hi);
#endif
    } else {
        nev[0] = ptx::mad_lo_cc(P::p(0), m, nev[0]);   // == 0
        nev[1] = ptx::madc_hi_cc(P::p(0), m, nev[1]);
    }
    nev[2] = ptx::madc_lo_cc(P::p(2), m, nev[2]);
    nev[3] = ptx::madc_hi_cc(P::p(2), m, nev[3]);
    nev[4] = ptx::madc_lo_cc(P::p(4), m, nev[4]);
    nev[5] = ptx::madc_hi_cc(P::p(4), m, nev[5]);
    nev[6] = ptx::madc_lo_cc(P::p(6), m, nev[6]);
    nev[7] = ptx::madc_hi_cc(P::p(6), m, nev[7]);
    nev[8] = ptx::addc(nev[8], 0u);
}

// t (8 limbs, < 2p) -> t mod p
template <class P>
ACG_HD fr_t fr_reduce_once(const uint32_t t[8]) {
    fr_t s;
    s.l[0] = ptx::sub_cc(t[0], P::p(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) s.l[i] = ptx::subc_cc(t[i], P::p(i));
    uint32_t borrow = ptx::subc(0u, 0u);
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = borrow ? t[i] : s.l[i];
    return r;
}

// Montgomery product a*b*2^-256 mod p, result in [0, p).  a <= p (the vector operand: every round keeps
// T < a + p); b is consumed limb by limb and may be ANY 256-bit value (an unreduced sum).
template <class P>
ACG_HD fr_t fr_mul(const fr_t& a, const fr_t& b) {
    uint32_t e0[9], o0[8], e1[9], o1[8];
    mont_round<P, true>(e0, o0, e0, o0, a.l, b.l[0]);
    mont_round<P, false>(e1, o1, e0, o0, a.l, b.l[1]);
    mont_round<P, false>(e0, o0, e1, o1, a.l, b.l[2]);
    mont_round<P, false>(e1, o1, e0, o0, a.l, b.l[3]);
    mont_round<P, false>(e0, o0, e1, o1, a.l, b.l[4]);
    mont_round<P, false>(e1, o1, e0, o0, a.l, b.l[5]);
    mont_round<P, false>(e0, o0, e1, o1, a.l, b.l[6]);
    mont_round<P, false>(e1, o1, e0, o0, a.l, b.l[7]);
    // T/2^32 = od + (ev >> 32)
    uint32_t t[8];
    t[0] = ptx::add_cc(o1[0], e1[1]);
#pragma unroll
    for (int i = 1; i < 7; ++i) t[i] = ptx::addc_cc(o1[i], e1[i + 1]);
    t[7] = ptx::addc(o1[7], e1[8]);
    return fr_reduce_once<P>(t);
}
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

template <class P>
ACG_HD fr_t fr_sqr(const fr_t& a) {
    return fr_mul<P>(a, a);
}
template <class P>
ACG_HD fr_t fr_to_mont(const fr_t& a) {
    fr_t r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.l[i] = P::r2(i);
    return fr_mul<P>(a, r2);
}
template <class P>
ACG_HD fr_t fr_from_mont(const fr_t& a) {
    fr_t one;
#pragma unroll
    for (int i = 0; i < 8; ++i) one.l[i] = i == 0 ? 1u : 0u;
    return fr_mul<P>(a, one);
}
// a^e, e given as 8 LE limbs (not secret; plain square-and-multiply, MSB first)
template <class P>
ACG_HD fr_t fr_pow(const fr_t& a, const uint32_t e[8]) {
    fr_t acc = fr_one<P>();
    for (int i = 255; i >= 0; --i) {
        acc = fr_sqr<P>(acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fr_mul<P>(acc, a);
    }
    return acc;
}
// a^(p-2); inv(0) = 0 (callers that need the reference's `recip` semantics guard a == 0 themselves)
template <class P>
ACG_HD fr_t fr_inv(const fr_t& a) {
    uint32_t e[8];
    e[0] = ptx::sub_cc(P::p(0), 2u);
#pragma unroll
    for (int i = 1; i < 8; ++i) e[i] = ptx::subc_cc(P::p(i), 0u);
    return fr_pow<P>(a, e);
}
template <class P>
ACG_HD fr_t fr_from_u64(uint64_t v) {  // small integer -> Montgomery form
    fr_t r = fr_zero<P>();
    r.l[0] = (uint32_t)v;
    r.l[1] = (uint32_t)(v >> 32);
    return fr_to_mont<P>(r);
}
template <class P>
ACG_HD fr_t fr_dbl(const fr_t& a) {
    return fr_add<P>(a, a);
}

}  // namespace acg
