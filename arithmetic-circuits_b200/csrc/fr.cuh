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
