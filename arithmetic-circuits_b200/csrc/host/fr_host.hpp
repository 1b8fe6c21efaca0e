// Host-side Fr (4x64-bit Montgomery, unsigned __int128) used by the C++ mirror of the reference's host
// logic: witness generation (evalGate needs * and recip, src/Circuit/Arithmetic.hs:120-133) and R1CS row
// construction (affineCircuitToAffineMap needs + and *, src/Circuit/Affine.hs:90-105).  Bulk arithmetic
// never runs here -- it runs in the CUDA kernels; this is per-gate host bookkeeping, like the
// reference's own Haskell host code.  Constants come from the same generated table as the device code.
#pragma once
#include <cstdint>
#include <cstring>

#include "../fr.cuh"

namespace acg {
namespace host {

typedef unsigned __int128 u128;

struct El {
    uint64_t v[4];
    bool operator==(const El& o) const { return std::memcmp(v, o.v, 32) == 0; }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
};

template <class P>
struct Fr {
    static uint64_t limb(uint32_t (*f)(int), int i) { return (uint64_t)f(2 * i) | ((uint64_t)f(2 * i + 1) << 32); }
    static El constant(uint32_t (*f)(int)) {
        El e;
        for (int i = 0; i < 4; ++i) e.v[i] = limb(f, i);
        return e;
    }
    static const El& modulus() {
        static const El m = constant(&P::p);
        return m;
    }
    static El zero() { return El{{0, 0, 0, 0}}; }
    static El one() {  // Montgomery 1
        static const El o = constant(&P::one);
        return o;
    }
    static El minus_one() {
        static const El o = constant(&P::minus_one);
        return o;
    }
    static bool geq_mod(const El& a) {
        const El& p = modulus();
        for (int i = 3; i >= 0; --i) {
            if (a.v[i] > p.v[i]) return true;
            if (a.v[i] < p.v[i]) return false;
        }
        return true;
    }
    static El add(const El& a, const El& b) {
        const El& p = modulus();
        El t, s;
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)a.v[i] + b.v[i];
            t.v[i] = (uint64_t)c;
            c >>= 64;
        }
        uint64_t br = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)t.v[i] - p.v[i] - br;
            s.v[i] = (uint64_t)d;
            br = (uint64_t)(d >> 64) & 1;
        }
        return (c || !br) ? s : t;
    }
    static El sub(const El& a, const El& b) {
        const El& p = modulus();
        El t;
        uint64_t br = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)a.v[i] - b.v[i] - br;
            t.v[i] = (uint64_t)d;
            br = (uint64_t)(d >> 64) & 1;
        }
        if (br) {
            u128 c = 0;
            for (int i = 0; i < 4; ++i) {
                c += (u128)t.v[i] + p.v[i];
                t.v[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return t;
    }
    static El neg(const El& a) { return sub(zero(), a); }
    static El mul(const El& a, const El& b) {  // Montgomery product
        const El& p = modulus();
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)a.v[j] * b.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * P::NINV64;
            c = (u128)m * p.v[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)m * p.v[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        El r{{t[0], t[1], t[2], t[3]}}, s;
        uint64_t br = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)r.v[i] - p.v[i] - br;
            s.v[i] = (uint64_t)d;
            br = (uint64_t)(d >> 64) & 1;
        }
        return (t[4] || !br) ? s : r;
    }
    static El to_mont(const El& a) {
        static const El r2 = constant(&P::r2);
        return mul(a, r2);
    }
    static El from_mont(const El& a) { return mul(a, El{{1, 0, 0, 0}}); }
    static El from_u64(uint64_t x) { return to_mont(El{{x, 0, 0, 0}}); }
    static El pow(El base, const uint64_t e[4]) {
        El acc = one();
        for (int i = 0; i < 256; ++i) {
            if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, base);
            base = mul(base, base);
        }
        return acc;
    }
    static El inv(const El& a) {  // a^(p-2); inv(0) = 0
        const El& p = modulus();
        uint64_t e[4] = {p.v[0] - 2, p.v[1], p.v[2], p.v[3]};  // p is odd and > 2: no borrow
        return pow(a, e);
    }
};

}  // namespace host
}  // namespace acg
