// K5 -- Lagrange interpolation through arbitrary distinct roots, n <= 4096.
//
// Replaces createPolynomials.lagrangeInterpolate of the reference (src/QAP.hs:495-508):
//   P(X) = sum_i (y_i / phi_i) * (M(X) / (X - x_i)),  M = prod (X - x_i),  phi_i = M'(x_i),
// and the target polynomial prod (X - root) of src/QAP.hs:492.  O(n^2) field products per polynomial
// (the reference's own build is O(n^3) through repeated polynomial division); bound by the integer
// pipe.  Three kernels: master polynomial (one CTA, n sequential rank-1 updates), barycentric weights
// (thread per node, one inversion each), combination (one CTA per polynomial: every thread runs the
// synthetic-division recurrence q_i[j-1] = m_j + x_i q_i[j] for its nodes and the CTA reduces
// sum_i s_i q_i[j] per coefficient with warp shuffles).
#include "dev.cuh"
#include "kernels.h"

namespace acg {

__device__ __forceinline__ fr_t ld_cg(const fr_t* p) {  // L1-bypassing load (data written by other threads)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = __ldcg(q), b = __ldcg(q + 1);
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// M(X) = prod_{i<n} (X - x_i): n+1 coefficients.  Ping-pongs between buf0 and buf1; the result lands
// in buf0 when n is even, buf1 when n is odd.
template <class P>
__global__ void __launch_bounds__(1024) k_lagrange_master(const fr_t* __restrict__ xs, uint32_t n, fr_t* buf0,
                                                          fr_t* buf1) {
    if (threadIdx.x == 0) buf0[0] = fr_one<P>();
    __syncthreads();
    for (uint32_t i = 0; i < n; ++i) {
        const fr_t* src = (i & 1u) ? buf1 : buf0;
        fr_t* dst = (i & 1u) ? buf0 : buf1;
        const fr_t xi = xs[i];
        for (uint32_t j = threadIdx.x; j <= i + 1u; j += blockDim.x) {
            const fr_t lo = j >= 1u ? ld_cg(src + (j - 1u)) : fr_zero<P>();
            fr_t v = lo;
            if (j <= i) {
                const fr_t hi = ld_cg(src + j);
                v = fr_sub<P>(lo, fr_mul<P>(xi, hi));
            }
            dst[j] = v;
        }
        __syncthreads();
    }
}

// wts[i] = 1 / prod_{j != i} (x_i - x_j); *status = 1 when two nodes coincide
template <class P>
__global__ void k_lagrange_weights(const fr_t* __restrict__ xs, uint32_t n, fr_t* __restrict__ wts,
                                   int* __restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr_t xi = xs[i];
    fr_t acc = fr_one<P>();
    for (uint32_t j = 0; j < n; ++j) {
        if (j == i) continue;
        const fr_t xj = xs[j];
        acc = fr_mul<P>(acc, fr_sub<P>(xi, xj));
    }
    if (fr_is_zero(acc)) *status = 1;
    wts[i] = fr_inv<P>(acc);
}

template <class P>
__device__ __forceinline__ fr_t warp_sum(fr_t v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        fr_t o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.l[k] = __shfl_xor_sync(0xffffffffu, v.l[k], off);
        v = fr_add<P>(v, o);
    }
    return v;
}

// One CTA per polynomial.  master: n+1 coefficients (monic).  out: n coefficients.
template <class P>
__global__ void __launch_bounds__(256) k_lagrange_combine(const fr_t* __restrict__ xs, const fr_t* __restrict__ ys,
                                                          const fr_t* __restrict__ wts, const fr_t* __restrict__ master,
                                                          uint32_t n, fr_t* __restrict__ sbuf,
                                                          fr_t* __restrict__ out) {
    extern __shared__ __align__(32) uint8_t lag_smem[];
    fr_t* q = reinterpret_cast<fr_t*>(lag_smem);  // q_i[j] for the current j, i < n  (n * 32 B <= 128 KB)
    fr_t* s = sbuf + (size_t)blockIdx.x * n;      // s_i = y_i * w_i (each thread re-reads only its own writes)
    __shared__ fr_t warp_part[8];
    const uint32_t poly = blockIdx.x;
    const fr_t* y = ys + (size_t)poly * n;
    fr_t* o = out + (size_t)poly * n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const fr_t yi = y[i], wi = wts[i];
        s[i] = fr_mul<P>(yi, wi);
        q[i] = fr_one<P>();  // q_i[n-1] = 1 (M is monic)
    }
    __syncthreads();
    for (uint32_t jj = n; jj-- > 0;) {  // coefficient index jj = n-1 .. 0
        fr_t part = fr_zero<P>();
        const fr_t mj = master[jj];  // used to step q down to index jj-1 (needs m_jj)
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const fr_t qi = q[i];
            const fr_t si = s[i];
            part = fr_add<P>(part, fr_mul<P>(si, qi));
            if (jj > 0) {
                const fr_t xi = xs[i];
                q[i] = fr_add<P>(mj, fr_mul<P>(xi, qi));  // q_i[jj-1] = m_jj + x_i q_i[jj]
            }
        }
        part = warp_sum<P>(part);
        if ((threadIdx.x & 31u) == 0u) warp_part[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x < 32u) {
            fr_t v = threadIdx.x < (blockDim.x >> 5) ? warp_part[threadIdx.x] : fr_zero<P>();
            v = warp_sum<P>(v);
            if (threadIdx.x == 0u) o[jj] = v;
        }
        __syncthreads();
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

template <class P>
static cudaError_t lagrange_impl(const fr_t* xs, const fr_t* ys, uint32_t n, uint32_t n_polys, fr_t* coeffs,
                                 fr_t* target, fr_t* scratch, int* d_status, cudaStream_t s, uint32_t* launches) {
    fr_t* buf0 = scratch;                // n+1
    fr_t* buf1 = scratch + (n + 1);      // n+1
    fr_t* wts = scratch + 2 * (n + 1);   // n
    fr_t* sbuf = wts + n;                // n_polys * n
    k_lagrange_master<P><<<1, 1024, 0, s>>>(xs, n, buf0, buf1);
    const fr_t* master = (n & 1u) ? buf1 : buf0;
    k_lagrange_weights<P><<<(n + 127) / 128, 128, 0, s>>>(xs, n, wts, d_status);
    *launches += 2;
    if (n_polys) {
        const size_t smem = (size_t)n * sizeof(fr_t);
        cudaError_t e = cudaFuncSetAttribute(k_lagrange_combine<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        k_lagrange_combine<P><<<n_polys, 256, smem, s>>>(xs, ys, wts, master, n, sbuf, coeffs);
        *launches += 1;
    }
    if (target) {
        cudaError_t e = cudaMemcpyAsync(target, master, (size_t)(n + 1) * sizeof(fr_t), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

cudaError_t launch_lagrange(int field, const fr_t* xs, const fr_t* ys, uint32_t n, uint32_t n_polys, fr_t* coeffs,
                            fr_t* target, fr_t* scratch, int* d_status, cudaStream_t s, uint32_t* launches) {
    ACG_DISPATCH_FIELD(field, return (lagrange_impl<P>(xs, ys, n, n_polys, coeffs, target, scratch, d_status, s,
                                                       launches)));
    return cudaErrorInvalidValue;
}

}  // namespace acg
