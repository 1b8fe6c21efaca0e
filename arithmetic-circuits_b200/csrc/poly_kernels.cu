// K7 -- exact polynomial division by the target on the device.
//
// Replaces `quotRem inputOutputPoly qapTarget` of verificationWitnessZk (reference src/QAP.hs:325-327; the
// quotRem itself is Data.Euclidean on poly-0.4 VPoly: schoolbook long division) for a target of ANY shape:
// prod (X - root_i) of a Lagrange-built QAP (src/QAP.hs:492), FFT.fftTargetPoly (:524), X^N - 1.
// Polynomials are dense little-endian coefficient arrays in Montgomery form.
//
// Long division is a chain: quotient coefficient t needs the dividend after the updates of all higher quotient
// coefficients.  The chain is cut into blocks of kDivBlock coefficients:
//   k_divmod_head   (one warp)   the block's quotient coefficients from the top kDivBlock coefficients of the
//                                running dividend and of the target -- the only sequential part, one product deep
//                                per coefficient;
//   k_divmod_update (whole grid) p[x] -= sum_t q_t * T[x - t] for the n coefficients below the block: a rank-
//                                kDivBlock update, one thread per coefficient.
// n * (deg p - n + 1) products in total, as the schoolbook division the reference performs, spread over the chip.
#include "dev.cuh"
#include "kernels.h"

namespace acg {

constexpr uint32_t kDivBlock = 32;

// v[i] += s * t[i]
template <class P>
__global__ void k_axpy1(fr_t* __restrict__ v, const fr_t* __restrict__ t, fr_t s, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        v[i] = fr_add<P>(v[i], fr_mul<P>(s, t[i]));
}

// *flag |= 1 when any of v[0..n) is non-zero (Montgomery zero is all-zero limbs)
__global__ void k_any_nonzero(const fr_t* __restrict__ v, uint64_t n, int* __restrict__ flag) {
    bool nz = false;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        nz |= !fr_is_zero(v[i]);
    if (__any_sync(0xffffffffu, nz) && lane_id() == 0u) atomicOr(flag, 1);
}

__device__ __forceinline__ fr_t shfl_fr(const fr_t& x, uint32_t src) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_sync(0xffffffffu, x.l[i], src);
    return r;
}

// One warp.  Quotient coefficients q[lo .. lo + cnt) of p / T, T of degree n with leading coefficient 1 / lc_inv:
// lane s holds pw_s = p[lo + n + s]; from the top, q_t = pw_t * lc_inv, and every lane below takes its share of
// q_t * T off its own coefficient.  Writes q to h[lo ..].
template <class P>
__global__ void __launch_bounds__(32) k_divmod_head(const fr_t* __restrict__ p, const fr_t* __restrict__ T, uint32_t n,
                                                    uint32_t lo, uint32_t cnt, fr_t lc_inv, int monic,
                                                    fr_t* __restrict__ h) {
    const uint32_t s = threadIdx.x;
    fr_t pw = s < cnt ? p[(uint64_t)lo + n + s] : fr_zero<P>();
    for (uint32_t t = cnt; t-- > 0;) {
        fr_t q = shfl_fr(pw, t);
        if (!monic) q = fr_mul<P>(q, lc_inv);
        if (s == t) h[lo + t] = q;
        const uint32_t d = t - s;  // coefficient of T that meets lane s: T[n - d], d >= 1
        if (s < t && d <= n) pw = fr_sub<P>(pw, fr_mul<P>(q, T[n - d]));
    }
}

// p[x] -= sum_{t < cnt} h[lo + t] * T[x - lo - t]   for x in [lo, lo + n): the n coefficients below the block
template <class P>
__global__ void __launch_bounds__(128) k_divmod_update(fr_t* __restrict__ p, const fr_t* __restrict__ T, uint32_t n,
                                                       uint32_t lo, uint32_t cnt, const fr_t* __restrict__ h) {
    __shared__ fr_t q[kDivBlock];
    if (threadIdx.x < cnt) q[threadIdx.x] = h[lo + threadIdx.x];
    __syncthreads();
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        fr_t acc = p[(uint64_t)lo + j];
        for (uint32_t t = 0; t < cnt && t <= j; ++t) acc = fr_sub<P>(acc, fr_mul<P>(q[t], T[j - t]));
        p[(uint64_t)lo + j] = acc;
    }
}

#define ACG_DISPATCH_FIELD(field, EXPR)   \
    do {                                  \
        if ((field) == 0) {               \
            using P = Bn254Fr;            \
            EXPR;                         \
        } else if ((field) == 1) {        \
            using P = Bls12381Fr;         \
            EXPR;                         \
        } else {                          \
            return cudaErrorInvalidValue; \
        }                                 \
    } while (0)

static inline unsigned grid_for(uint64_t n, unsigned block, unsigned max_blocks) {
    uint64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (unsigned)g;
}

cudaError_t launch_axpy1(int field, fr_t* v, const fr_t* t, fr_t s, uint64_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    ACG_DISPATCH_FIELD(field, (k_axpy1<P><<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(v, t, s, n)));
    return cudaGetLastError();
}

cudaError_t launch_any_nonzero(const fr_t* v, uint64_t n, int* d_flag, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k_any_nonzero<<<grid_for(n, 256, 148 * 8), 256, 0, s>>>(v, n, d_flag);
    return cudaGetLastError();
}

template <class P>
static cudaError_t divmod_impl(fr_t* p, uint32_t len_p, const fr_t* T, uint32_t n, fr_t lc_inv, bool monic, fr_t* h,
                               cudaStream_t s, uint32_t* launches) {
    // quotient coefficients len_p - 1 - n down to 0, kDivBlock at a time
    const uint32_t n_q = len_p - n;
    for (uint32_t done = 0; done < n_q;) {
        const uint32_t cnt = n_q - done < kDivBlock ? n_q - done : kDivBlock;
        const uint32_t lo = n_q - done - cnt;
        k_divmod_head<P><<<1, 32, 0, s>>>(p, T, n, lo, cnt, lc_inv, monic ? 1 : 0, h);
        ++*launches;
        if (n) {
            k_divmod_update<P><<<grid_for(n, 128, 148 * 8), 128, 0, s>>>(p, T, n, lo, cnt, h);
            ++*launches;
        }
        done += cnt;
    }
    return cudaGetLastError();
}

// p: len_p coefficients (destroyed: on return p[0..n) is the remainder), T: n + 1 coefficients with T[n] != 0 and
// lc_inv = 1 / T[n]; h: len_p - n quotient coefficients.  Requires len_p > n.
cudaError_t launch_poly_divmod(int field, fr_t* p, uint32_t len_p, const fr_t* T, uint32_t n, fr_t lc_inv, bool monic,
                               fr_t* h, cudaStream_t s, uint32_t* launches) {
    if (len_p <= n) return cudaErrorInvalidValue;
    ACG_DISPATCH_FIELD(field, return (divmod_impl<P>(p, len_p, T, n, lc_inv, monic, h, s, launches)));
    return cudaErrorInvalidValue;
}

}  // namespace acg
