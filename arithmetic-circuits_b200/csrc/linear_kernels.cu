// Linear-constraint check over ANY 256-bit prime modulus (second half).
//
// Replaces checkLinearConstraint of the reference's Bulletproofs backend (src/Circuit/Bulletproofs.hs:329-338):
//     wL . aL + wR . aR + wO . aO == wV . v + c          over the scalar field of secp256k1,
// each dot product as src/Circuit/Affine.hs:121-125 (a wire the assignment lacks counts as 0).  It is the sparse-dot
// shape of K2 with B = 1 and no product -- but the field is not: secp256k1's group order is >= 2^255, so K1's
// split-accumulator Montgomery product (fr.cuh), whose carry analysis and lazy row sums rely on a spare top bit, does
// not apply.  This file therefore carries its own textbook arithmetic on 4 x 64-bit limbs with an explicit carry word
// (CIOS, t of n + 2 words), parameterised on the modulus; it is correct for every odd modulus below 2^256 and is also
// instantiated for the two fields of the main path so that it can be tested against them.  A linear constraint costs
// one product per non-zero weight; the kernel is a thread per constraint over two CSR matrices.
#include "dev.cuh"
#include "kernels.h"

namespace acg {
namespace lin {

struct Secp256k1Fn {  // order of the secp256k1 group (the `Fr` of Data.Curve.Weierstrass.SECP256K1)
    static __host__ __device__ constexpr uint64_t n(int i) {
        constexpr uint64_t v[4] = {0xbfd25e8cd0364141ull, 0xbaaedce6af48a03bull, 0xfffffffffffffffeull, 0xffffffffffffffffull};
        return v[i];
    }
    static __host__ __device__ constexpr uint64_t r2(int i) {
        constexpr uint64_t v[4] = {0x896cf21467d7d140ull, 0x741496c20e7cf878ull, 0xe697f5e45bcd07c6ull, 0x9d671cd581c69bc5ull};
        return v[i];
    }
    static constexpr uint64_t ninv = 0x4b0dff665588b13full;
};
struct Bn254FrMod {
    static __host__ __device__ constexpr uint64_t n(int i) {
        constexpr uint64_t v[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
        return v[i];
    }
    static __host__ __device__ constexpr uint64_t r2(int i) {
        constexpr uint64_t v[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
        return v[i];
    }
    static constexpr uint64_t ninv = 0xc2e1f593efffffffull;
};
struct Bls12381FrMod {
    static __host__ __device__ constexpr uint64_t n(int i) {
        constexpr uint64_t v[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
        return v[i];
    }
    static __host__ __device__ constexpr uint64_t r2(int i) {
        constexpr uint64_t v[4] = {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull};
        return v[i];
    }
    static constexpr uint64_t ninv = 0xfffffffeffffffffull;
};

struct u256 {
    uint64_t w[4];
};
__device__ __forceinline__ u256 load_u256(const uint64_t* p) {
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(p), b = *reinterpret_cast<const ulonglong2*>(p + 2);
    return u256{{a.x, a.y, b.x, b.y}};
}
// a * b + c + carry: low word returned, high word left in carry (never overflows 128 bits)
__device__ __forceinline__ uint64_t mac(uint64_t a, uint64_t b, uint64_t c, uint64_t& carry) {
    uint64_t lo = a * b, hi = __umul64hi(a, b);
    lo += c;
    hi += lo < c;
    lo += carry;
    hi += lo < carry;
    carry = hi;
    return lo;
}
template <class M>
__device__ __forceinline__ bool geq_mod(const u256& a) {  // a >= n
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        if (a.w[i] > M::n(i)) return true;
        if (a.w[i] < M::n(i)) return false;
    }
    return true;
}
template <class M>
__device__ __forceinline__ u256 sub_mod_n(const u256& a) {  // a - n (mod 2^256)
    u256 r;
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t d = a.w[i] - M::n(i), d2 = d - borrow;
        borrow = (a.w[i] < M::n(i)) | (d < borrow);
        r.w[i] = d2;
    }
    return r;
}
// (a + b) mod n for a, b < n: the sum may carry out of 256 bits when n >= 2^255
template <class M>
__device__ __forceinline__ u256 add_mod(const u256& a, const u256& b) {
    u256 s;
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t t = a.w[i] + b.w[i], t2 = t + carry;
        carry = (t < a.w[i]) | (t2 < t);
        s.w[i] = t2;
    }
    return (carry || geq_mod<M>(s)) ? sub_mod_n<M>(s) : s;
}
// Montgomery product a * b * 2^-256 mod n (CIOS, t of 4 + 2 words): a, b < n, result < n
template <class M>
__device__ __forceinline__ u256 mont_mul(const u256& a, const u256& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = mac(a.w[j], b.w[i], t[j], c);
        uint64_t s = t[4] + c;
        t[5] = s < c;
        t[4] = s;
        const uint64_t m = t[0] * M::ninv;
        c = 0;
        (void)mac(m, M::n(0), t[0], c);
#pragma unroll
        for (int j = 1; j < 4; ++j) t[j - 1] = mac(m, M::n(j), t[j], c);
        s = t[4] + c;
        t[3] = s;
        t[4] = t[5] + (s < c);
    }
    const u256 r{{t[0], t[1], t[2], t[3]}};
    return (t[4] || geq_mod<M>(r)) ? sub_mod_n<M>(r) : r;
}

// x[i] <- x[i] * 2^256 mod n (Montgomery form), flagging elements >= n
template <class M>
__global__ void k_lin_to_mont(uint64_t* __restrict__ x, uint64_t n_el, int* __restrict__ bad_flag) {
    const u256 r2{{M::r2(0), M::r2(1), M::r2(2), M::r2(3)}};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_el; i += (uint64_t)gridDim.x * blockDim.x) {
        const u256 v = load_u256(x + 4 * i);
        if (geq_mod<M>(v)) {
            *bad_flag = 1;
            continue;
        }
        const u256 m = mont_mul<M>(v, r2);
#pragma unroll
        for (int k = 0; k < 4; ++k) x[4 * i + k] = m.w[k];
    }
}

// row . x for one CSR row: weights canonical, x in Montgomery form => every product is canonical
template <class M>
__device__ __forceinline__ u256 row_dot(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col,
                                        const uint64_t* __restrict__ val, const uint64_t* __restrict__ x_mont, uint32_t row,
                                        bool& bad) {
    u256 acc{{0, 0, 0, 0}};
    for (uint32_t k = rowptr[row]; k < rowptr[row + 1]; ++k) {
        const u256 wgt = load_u256(val + 4ull * k);
        bad |= geq_mod<M>(wgt);
        acc = add_mod<M>(acc, mont_mul<M>(wgt, load_u256(x_mont + 4ull * col[k])));
    }
    return acc;
}
// constraint i holds  <=>  lhs_i . x == rhs_i . v + c_i
template <class M>
__global__ void k_linear_constraints(const uint32_t* __restrict__ l_rowptr, const uint32_t* __restrict__ l_col,
                                     const uint64_t* __restrict__ l_val, const uint32_t* __restrict__ r_rowptr,
                                     const uint32_t* __restrict__ r_col, const uint64_t* __restrict__ r_val,
                                     const uint64_t* __restrict__ cst, const uint64_t* __restrict__ x_mont,
                                     const uint64_t* __restrict__ v_mont, uint32_t n_constraints,
                                     unsigned long long* __restrict__ result, int* __restrict__ bad_flag) {
    const uint32_t n_groups = (n_constraints + 31u) / 32u;
    const uint32_t warps_per_block = blockDim.x / 32u;
    for (uint32_t g = blockIdx.x * warps_per_block + threadIdx.x / 32u; g < n_groups; g += gridDim.x * warps_per_block) {
        const uint32_t i = g * 32u + lane_id();
        bool violated = false, bad = false;
        if (i < n_constraints) {
            const u256 lhs = row_dot<M>(l_rowptr, l_col, l_val, x_mont, i, bad);
            const u256 c = load_u256(cst + 4ull * i);
            bad |= geq_mod<M>(c);
            const u256 rhs = add_mod<M>(row_dot<M>(r_rowptr, r_col, r_val, v_mont, i, bad), c);
            violated = (lhs.w[0] ^ rhs.w[0]) | (lhs.w[1] ^ rhs.w[1]) | (lhs.w[2] ^ rhs.w[2]) | (lhs.w[3] ^ rhs.w[3]);
        }
        if (bad) *bad_flag = 1;
        const uint32_t bal = __ballot_sync(0xffffffffu, violated);
        if (bal != 0u && lane_id() == 0u) report_bad_rows(result, bal, (uint64_t)g * 32u);
    }
}

}  // namespace lin

static inline unsigned lin_grid(uint64_t n, unsigned block, unsigned max_blocks) {
    uint64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (unsigned)g;
}

#define ACG_DISPATCH_MODULUS(modulus, EXPR)   \
    do {                                      \
        if ((modulus) == 0) {                 \
            using M = lin::Bn254FrMod;        \
            EXPR;                             \
        } else if ((modulus) == 1) {          \
            using M = lin::Bls12381FrMod;     \
            EXPR;                             \
        } else if ((modulus) == 2) {          \
            using M = lin::Secp256k1Fn;       \
            EXPR;                             \
        } else {                              \
            return cudaErrorInvalidValue;     \
        }                                     \
    } while (0)

cudaError_t launch_lin_to_mont(int modulus, uint64_t* x, uint64_t n_el, int* d_bad_flag, cudaStream_t s) {
    if (n_el == 0) return cudaSuccess;
    ACG_DISPATCH_MODULUS(modulus, (lin::k_lin_to_mont<M><<<lin_grid(n_el, 128, 148 * 16), 128, 0, s>>>(x, n_el, d_bad_flag)));
    return cudaGetLastError();
}

// d_result: {violated constraints, first violated constraint} -- must hold {0, ~0} on entry
cudaError_t launch_linear_constraints(int modulus, const uint32_t* l_rowptr, const uint32_t* l_col, const uint64_t* l_val,
                                      const uint32_t* r_rowptr, const uint32_t* r_col, const uint64_t* r_val,
                                      const uint64_t* cst, const uint64_t* x_mont, const uint64_t* v_mont,
                                      uint32_t n_constraints, unsigned long long* d_result, int* d_bad_flag,
                                      cudaStream_t s) {
    if (n_constraints == 0) return cudaSuccess;
    ACG_DISPATCH_MODULUS(modulus, (lin::k_linear_constraints<M><<<lin_grid(n_constraints, 128, 148 * 16), 128, 0, s>>>(
                                      l_rowptr, l_col, l_val, r_rowptr, r_col, r_val, cst, x_mont, v_mont, n_constraints,
                                      d_result, d_bad_flag)));
    return cudaGetLastError();
}

}  // namespace acg
