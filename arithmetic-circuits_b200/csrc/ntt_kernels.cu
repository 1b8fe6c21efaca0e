// K3 -- radix-2 Cooley-Tukey NTT over Fr, and K4 -- the pointwise kernels of the QAP quotient.
//
// Replaces FFT.interpolate / the DFT of galois-fft-0.1.0 as called at reference src/QAP.hs:521-523
// (omega = getRootOfUnity k = g^((r-1)/2^k), pairing-1.0.0), and the polynomial product / division
// of src/QAP.hs:325-327 done on a coset of the evaluation domain (T = X^N - 1).
//
// Structure: decimation in frequency, natural order in -> bit-reversed order out, as a sequence of
// passes.  A pass runs k <= 8 (strided "column" pass) or k <= 11 (final contiguous pass) butterfly
// stages on 2048-element tiles held in shared memory (64 KB, split in two 16-byte planes so 128-bit
// accesses of neighbouring lanes never collide on a bank).  Column passes load C adjacent columns so
// every global access is a contiguous C*32-byte segment, and finish with the four-step twiddle
// correction w_M^(column * bitrev(position)), fetched from a two-level power table.
// Global traffic: one read + one write of the vector per pass; the work is bound by the integer
// pipe (one 256-bit Montgomery product per butterfly), not by HBM (DESIGN.md "K3").
#include "dev.cuh"
#include "kernels.h"

#include <cstdlib>
#include <vector>

namespace acg {

constexpr uint32_t kNttTileLog = 11;  // 2048 elements per CTA tile
constexpr uint32_t kNttTile = 1u << kNttTileLog;
constexpr uint32_t kNttThreads = 256;
constexpr uint32_t kNttSmallLog = 11;  // small twiddle table: w_{2^11}^i, i < 2^10

struct NttPass {
    uint32_t k;       // butterfly stages in this pass
    uint32_t log_s;   // log2 stride between sub-transform elements (0 for the final pass)
    uint32_t log_c;   // log2 columns per tile
    uint32_t log_g;   // log2 sub-transforms per tile (final pass)
    uint32_t log_m;   // log2 length of the transform this pass belongs to (k + log_s)
    bool correct;     // apply the four-step twiddle after the stages
};

struct NttPlan {
    int field;
    uint32_t log_n;
    bool inverse;
    std::vector<NttPass> passes;
    fr_t* d_small = nullptr;  // w_{2^11}^i  (or w_{2^log_n}^i when log_n < 11), i < 2^10
    uint32_t small_log = 0;
    fr_t* d_hi = nullptr;     // w_N^(i << lo_bits)
    fr_t* d_lo = nullptr;     // w_N^i, i < 2^lo_bits
    uint32_t lo_bits = 0;
    fr_t scale;               // 1/n (Montgomery) for inverse plans
    // Four-step twiddle corrections of the column passes as full tables, indexed like the data (one factor per
    // element): one product per element instead of two (table lookup product + application).  The kernel is
    // bound by the 32x32->64 multiplier, not by HBM, so trading 32 B of extra read per
    // element for a 256-bit product is a net win; the tables cost 32 * N bytes per column pass.
    // Inverse plans fold 1/n into the table of the last column pass.
    std::vector<fr_t*> d_corr;
    bool scale_folded = false;
};

__device__ __forceinline__ uint32_t brev_bits(uint32_t x, uint32_t bits) {
    return bits ? (__brev(x) >> (32u - bits)) : 0u;
}

// out[i] = base^(i << shift), i < n
template <class P>
__global__ void k_fill_powers(fr_t* __restrict__ out, uint32_t n, fr_t base, uint32_t shift) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t e = (uint64_t)i << shift;
        fr_t acc = fr_one<P>();
        fr_t b = base;
        while (e) {
            if (e & 1ull) acc = fr_mul<P>(acc, b);
            b = fr_sqr<P>(b);
            e >>= 1;
        }
        out[i] = acc;
    }
}

struct SmemPlanes {
    uint4* lo;
    uint4* hi;
    __device__ __forceinline__ fr_t load(uint32_t i) const {
        const uint4 a = lo[i], b = hi[i];
        fr_t r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void store(uint32_t i, const fr_t& v) const {
        lo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};

template <class P>
__global__ void __launch_bounds__(kNttThreads)
    k_ntt_pass(const fr_t* src, fr_t* dst, uint32_t permute_log, const fr_t* __restrict__ tw_small,
               uint32_t small_log, uint32_t k,
               uint32_t log_s, uint32_t log_c, uint32_t log_g, const fr_t* __restrict__ tw_hi,
               const fr_t* __restrict__ tw_lo, uint32_t lo_bits, uint32_t corr_shift, int apply_corr, fr_t scale,
               int apply_scale, const fr_t* __restrict__ corr, uint64_t corr_mask) {
    extern __shared__ __align__(16) uint8_t ntt_smem[];
    const uint32_t tile_log = k + log_c + log_g;
    const uint32_t tile_n = 1u << tile_log;
    SmemPlanes sm{reinterpret_cast<uint4*>(ntt_smem), reinterpret_cast<uint4*>(ntt_smem) + tile_n};

    const uint32_t tid = threadIdx.x;
    const uint64_t tile = blockIdx.x;
    // tile -> (first outer block, column block)
    const uint32_t log_tiles_per_outer = log_s - log_c;  // column tiles per outer block (0 when final)
    const uint64_t outer0 = (tile >> log_tiles_per_outer) << log_g;
    const uint32_t cb = (uint32_t)(tile & ((1ull << log_tiles_per_outer) - 1ull));
    const uint32_t cmask = (1u << log_c) - 1u;
    const uint32_t kmask = (1u << k) - 1u;

    // ---- load: local index L = ((g << k) + j) << log_c | c
    for (uint32_t L = tid; L < tile_n; L += kNttThreads) {
        const uint32_t c = L & cmask;
        const uint32_t j = (L >> log_c) & kmask;
        const uint32_t g = L >> (log_c + k);
        const uint64_t addr = ((((outer0 + g) << k) + j) << log_s) + ((uint64_t)cb << log_c) + c;
        const fr_t v = src[addr];
        sm.store(L, v);
    }
    __syncthreads();

    // ---- k DIF stages
    const uint32_t half_n = tile_n >> 1;
    for (uint32_t t = 0; t < k; ++t) {
        const uint32_t hb = k - 1u - t;  // log2 of the half distance (in j units)
        for (uint32_t b = tid; b < half_n; b += kNttThreads) {
            const uint32_t c = b & cmask;
            const uint32_t jb = (b >> log_c) & (kmask >> 1);
            const uint32_t g = b >> (log_c + k - 1u);
            const uint32_t o = jb & ((1u << hb) - 1u);
            const uint32_t q = jb >> hb;
            const uint32_t j0 = (q << (hb + 1u)) | o;
            const uint32_t L0 = ((((g << k) + j0)) << log_c) | c;
            const uint32_t L1 = L0 + ((1u << hb) << log_c);
            const fr_t u = sm.load(L0);
            const fr_t v = sm.load(L1);
            sm.store(L0, fr_add<P>(u, v));
            fr_t d = fr_sub<P>(u, v);
            if (hb != 0u) {  // last stage: twiddle is 1
                const fr_t tw = tw_small[(o << t) << (small_log - k)];
                d = fr_mul<P>(d, tw);
            }
            sm.store(L1, d);
        }
        __syncthreads();
    }

    // ---- store (+ four-step twiddle, + 1/n scale)
    for (uint32_t L = tid; L < tile_n; L += kNttThreads) {
        const uint32_t c = L & cmask;
        const uint32_t j = (L >> log_c) & kmask;
        const uint32_t g = L >> (log_c + k);
        const uint64_t addr = ((((outer0 + g) << k) + j) << log_s) + ((uint64_t)cb << log_c) + c;
        fr_t v = sm.load(L);
        if (corr) {
            v = fr_mul<P>(v, corr[addr & corr_mask]);
        } else if (apply_corr) {
            const uint64_t n2 = ((uint64_t)cb << log_c) + c;
            const uint64_t e = (n2 * (uint64_t)brev_bits(j, k)) << corr_shift;
            if (e != 0ull) {
                const fr_t th = tw_hi[e >> lo_bits];
                const fr_t tl = tw_lo[e & ((1ull << lo_bits) - 1ull)];
                v = fr_mul<P>(v, fr_mul<P>(th, tl));
            }
        }
        if (apply_scale) v = fr_mul<P>(v, scale);
        if (permute_log) {  // last pass of a natural-order transform: undo the bit reversal on the way out
            const uint64_t m = (1ull << permute_log) - 1ull;
            dst[(addr & ~m) | brev_bits((uint32_t)(addr & m), permute_log)] = v;
        } else {
            dst[addr] = v;
        }
    }
}

// corr[addr] = w_N^((n2 * bitrev_k(j)) << corr_shift) (* scale), addr = (((outer << k) + j) << log_s) + n2
template <class P>
__global__ void k_fill_corr(fr_t* __restrict__ corr, uint64_t n, uint32_t k, uint32_t log_s, uint32_t corr_shift,
                            const fr_t* __restrict__ tw_hi, const fr_t* __restrict__ tw_lo, uint32_t lo_bits, fr_t scale,
                            int use_scale) {
    for (uint64_t a = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; a < n; a += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n2 = a & ((1ull << log_s) - 1ull);
        const uint32_t j = (uint32_t)(a >> log_s) & ((1u << k) - 1u);
        const uint64_t e = (n2 * (uint64_t)brev_bits(j, k)) << corr_shift;
        fr_t f = fr_mul<P>(tw_hi[e >> lo_bits], tw_lo[e & ((1ull << lo_bits) - 1ull)]);
        if (use_scale) f = fr_mul<P>(f, scale);
        corr[a] = f;
    }
}

// out[i] = in[bitrev(i)] within each 2^log_n block
__global__ void k_bitrev_permute(const fr_t* __restrict__ in, fr_t* __restrict__ out, uint32_t log_n, uint64_t total) {
    const uint64_t mask = (1ull << log_n) - 1ull;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t blk = i & ~mask;
        const uint32_t r = brev_bits((uint32_t)(i & mask), log_n);
        out[i] = in[blk | r];
    }
}

// v[i] *= pow_hi[e >> lo_bits] * pow_lo[e & mask] * scale, e = i or bitrev(i)
template <class P>
__global__ void k_scale_by_powers(fr_t* __restrict__ v, uint64_t n, const fr_t* __restrict__ pow_hi,
                                  const fr_t* __restrict__ pow_lo, uint32_t lo_bits, fr_t scale, int use_scale,
                                  int index_bitrev, uint32_t log_n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t e = index_bitrev ? brev_bits((uint32_t)i, log_n) : (uint32_t)i;
        fr_t f = fr_mul<P>(pow_hi[e >> lo_bits], pow_lo[e & ((1u << lo_bits) - 1u)]);
        if (use_scale) f = fr_mul<P>(f, scale);
        const fr_t x = v[i];
        v[i] = fr_mul<P>(x, f);
    }
}

template <class P>
__global__ void k_quotient_pointwise(const fr_t* __restrict__ a, const fr_t* __restrict__ b, const fr_t* __restrict__ c,
                                     fr_t* __restrict__ h, uint64_t n, fr_t zinv) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const fr_t x = a[i], y = b[i], z = c[i];
        h[i] = fr_mul<P>(fr_sub<P>(fr_mul<P>(x, y), z), zinv);
    }
}

template <class P>
__global__ void k_axpy2(fr_t* __restrict__ h, const fr_t* __restrict__ a, const fr_t* __restrict__ b, fr_t d2, fr_t d1,
                        uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const fr_t x = a[i], y = b[i], z = h[i];
        h[i] = fr_add<P>(z, fr_add<P>(fr_mul<P>(d2, x), fr_mul<P>(d1, y)));
    }
}

// out[i] = sum_k weight[k] * polys[k * len + i]: the scale-and-sum of verificationWitnessZk over a per-wire QAP
// (foldQapSet / combineWithDefaults, src/QAP.hs:163-181, 314-324).  Thread per coefficient, coalesced across i.
template <class P>
__global__ void k_poly_combine(const fr_t* __restrict__ polys, const fr_t* __restrict__ weight, uint32_t n_polys,
                               uint64_t len, fr_t* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
        fr_t acc = fr_zero<P>();
        for (uint32_t k = 0; k < n_polys; ++k) acc = fr_add<P>(acc, fr_mul<P>(polys[(uint64_t)k * len + i], weight[k]));
        out[i] = acc;
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

// ---- host-side field helpers for plan construction (same algorithms, host path of fr.cuh) ----------
template <class P>
static fr_t host_pow(fr_t base, uint64_t e) {
    fr_t acc = fr_one<P>();
    while (e) {
        if (e & 1ull) acc = fr_mul<P>(acc, base);
        base = fr_sqr<P>(base);
        e >>= 1;
    }
    return acc;
}
template <class P>
static fr_t host_root_of_unity(uint32_t k, bool inverse) {  // Montgomery form
    fr_t w;
    for (int i = 0; i < 8; ++i) w.l[i] = inverse ? P::two_adic_root_inv(i) : P::two_adic_root(i);
    for (uint32_t i = k; i < (uint32_t)P::TWO_ADICITY; ++i) w = fr_sqr<P>(w);
    return w;
}

template <class P>
static cudaError_t plan_fill(NttPlan* p) {
    const uint32_t log_n = p->log_n;
    cudaError_t e;
    // small table
    p->small_log = log_n < kNttSmallLog ? log_n : kNttSmallLog;
    const uint32_t small_n = p->small_log ? (1u << (p->small_log - 1)) : 1u;
    if ((e = cudaMalloc(&p->d_small, (size_t)small_n * sizeof(fr_t))) != cudaSuccess) return e;
    k_fill_powers<P><<<grid_for(small_n, 128, 1024), 128>>>(p->d_small, small_n,
                                                            host_root_of_unity<P>(p->small_log, p->inverse), 0);
    // two-level table of w_N powers
    p->lo_bits = (log_n + 1) / 2;
    const uint32_t lo_n = 1u << p->lo_bits, hi_n = 1u << (log_n - p->lo_bits);
    if ((e = cudaMalloc(&p->d_lo, (size_t)lo_n * sizeof(fr_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&p->d_hi, (size_t)hi_n * sizeof(fr_t))) != cudaSuccess) return e;
    const fr_t wn = host_root_of_unity<P>(log_n, p->inverse);
    k_fill_powers<P><<<grid_for(lo_n, 128, 1024), 128>>>(p->d_lo, lo_n, wn, 0);
    k_fill_powers<P><<<grid_for(hi_n, 128, 1024), 128>>>(p->d_hi, hi_n, wn, p->lo_bits);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // 1/n
    fr_t n_m = fr_from_u64<P>(1ull << log_n);
    p->scale = fr_inv<P>(n_m);
    // correction tables of the column passes (ACG_NTT_NO_CORR_TABLES=1 keeps the two-product on-the-fly form)
    p->d_corr.assign(p->passes.size(), nullptr);
    if (!getenv("ACG_NTT_NO_CORR_TABLES")) {
        size_t last_col = p->passes.size();
        for (size_t i = 0; i < p->passes.size(); ++i)
            if (p->passes[i].correct) last_col = i;
        for (size_t i = 0; i < p->passes.size(); ++i) {
            const NttPass& ps = p->passes[i];
            if (!ps.correct) continue;
            const uint64_t n = 1ull << log_n;
            if ((e = cudaMalloc(&p->d_corr[i], n * sizeof(fr_t))) != cudaSuccess) return e;
            const bool fold = p->inverse && i == last_col;
            k_fill_corr<P><<<grid_for(n, 256, 148 * 16), 256>>>(p->d_corr[i], n, ps.k, ps.log_s, log_n - ps.log_m, p->d_hi,
                                                               p->d_lo, p->lo_bits, p->scale, fold ? 1 : 0);
            if (fold) p->scale_folded = true;
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaDeviceSynchronize();
}

cudaError_t ntt_plan_create(int field, uint32_t log_n, bool inverse, NttPlan** out) {
    NttPlan* p = new NttPlan();
    p->field = field;
    p->log_n = log_n;
    p->inverse = inverse;
    // pass decomposition: column passes of <= 8 stages while more than 11 remain, then a contiguous pass
    uint32_t remaining = log_n;
    while (remaining > kNttTileLog) {
        uint32_t k = remaining - 6u < 8u ? remaining - 6u : 8u;
        NttPass ps;
        ps.k = k;
        ps.log_s = remaining - k;
        ps.log_c = kNttTileLog - k;
        if (ps.log_c > ps.log_s) ps.log_c = ps.log_s;
        ps.log_g = 0;
        ps.log_m = remaining;
        ps.correct = true;
        p->passes.push_back(ps);
        remaining -= k;
    }
    {
        NttPass ps;
        ps.k = remaining;
        ps.log_s = 0;
        ps.log_c = 0;
        ps.log_g = kNttTileLog - remaining;  // clipped at run time to the amount of data
        ps.log_m = remaining;
        ps.correct = false;
        p->passes.push_back(ps);
    }
    cudaError_t e = cudaErrorInvalidValue;
    if (field == 0) e = plan_fill<Bn254Fr>(p);
    if (field == 1) e = plan_fill<Bls12381Fr>(p);
    if (e != cudaSuccess) {
        ntt_plan_destroy(p);
        return e;
    }
    *out = p;
    return cudaSuccess;
}

void ntt_plan_destroy(NttPlan* p) {
    if (!p) return;
    cudaFree(p->d_small);
    cudaFree(p->d_hi);
    cudaFree(p->d_lo);
    for (fr_t* c : p->d_corr) cudaFree(c);
    delete p;
}

template <class P>
// permute: write natural order.  Passes work tile by tile (load a tile, transform it in shared memory, store it
// to the same addresses), so any pass can also run out of place; with more than one pass the one before the last
// stores to `scratch` and the last one reads it and scatters into `data`.  A single pass holds whole transforms in
// its tile and permutes in place.
static cudaError_t run_dif(const NttPlan* p, fr_t* data, fr_t* scratch, bool permute, uint32_t batch, bool scale_last,
                           cudaStream_t s, uint32_t* launches) {
    {
        cudaError_t e = cudaFuncSetAttribute(k_ntt_pass<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(kNttTile * sizeof(fr_t)));
        if (e != cudaSuccess) return e;
    }
    const uint64_t total = (uint64_t)batch << p->log_n;
    uint32_t total_log = 0;
    while ((1ull << total_log) < total) ++total_log;  // batch is a power of two or 1 in our callers
    if ((1ull << total_log) != total) return cudaErrorInvalidValue;
    for (size_t i = 0; i < p->passes.size(); ++i) {
        NttPass ps = p->passes[i];
        const bool last = i + 1 == p->passes.size();
        if (last) {
            const uint32_t avail = total_log - ps.k;  // log2 of sub-transforms available
            if (ps.log_g > avail) ps.log_g = avail;
        }
        const uint32_t tile_log = ps.k + ps.log_c + ps.log_g;
        const uint64_t n_tiles = total >> tile_log;
        const size_t smem = (size_t)sizeof(fr_t) << tile_log;
        const bool via_scratch = permute && p->passes.size() > 1;
        const fr_t* src = (via_scratch && last) ? scratch : data;
        fr_t* dst = (via_scratch && i + 2 == p->passes.size()) ? scratch : data;
        k_ntt_pass<P><<<(unsigned)n_tiles, kNttThreads, smem, s>>>(
            src, dst, (permute && last) ? p->log_n : 0u, p->d_small, p->small_log, ps.k, ps.log_s, ps.log_c, ps.log_g, p->d_hi, p->d_lo, p->lo_bits,
            p->log_n - ps.log_m, ps.correct ? 1 : 0, p->scale, (last && scale_last && !p->scale_folded) ? 1 : 0,
            p->d_corr.empty() ? nullptr : p->d_corr[i], (1ull << p->log_n) - 1ull);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t ntt_run_dif(const NttPlan* p, fr_t* data, uint32_t batch, cudaStream_t s, uint32_t* launches) {
    if (p->scale_folded) return cudaErrorInvalidValue;  // an inverse plan always scales: use ntt_run
    ACG_DISPATCH_FIELD(p->field, return (run_dif<P>(p, data, nullptr, false, batch, false, s, launches)));
    return cudaErrorInvalidValue;
}

cudaError_t ntt_run(const NttPlan* p, fr_t* data, fr_t* scratch, uint32_t batch, cudaStream_t s, uint32_t* launches) {
    ACG_DISPATCH_FIELD(p->field, return (run_dif<P>(p, data, scratch, p->log_n > 0, batch, p->inverse, s, launches)));
    return cudaErrorInvalidValue;
}

const fr_t* ntt_plan_pow_hi(const NttPlan* p) { return p->d_hi; }
const fr_t* ntt_plan_pow_lo(const NttPlan* p) { return p->d_lo; }
uint32_t ntt_plan_lo_bits(const NttPlan* p) { return p->lo_bits; }

cudaError_t launch_bitrev_permute(const fr_t* in, fr_t* out, uint32_t log_n, uint32_t batch, cudaStream_t s) {
    const uint64_t total = (uint64_t)batch << log_n;
    k_bitrev_permute<<<grid_for(total, 256, 148 * 32), 256, 0, s>>>(in, out, log_n, total);
    return cudaGetLastError();
}

cudaError_t launch_fill_powers(int field, fr_t* out, uint32_t n, fr_t base, uint32_t shift, cudaStream_t s) {
    ACG_DISPATCH_FIELD(field, (k_fill_powers<P><<<grid_for(n, 128, 1024), 128, 0, s>>>(out, n, base, shift)));
    return cudaGetLastError();
}

cudaError_t launch_scale_by_powers(int field, fr_t* v, uint64_t n, const fr_t* d_pow_hi, const fr_t* d_pow_lo,
                                   uint32_t lo_bits, fr_t scale, bool use_scale, bool index_bitrev, uint32_t log_n,
                                   cudaStream_t s) {
    ACG_DISPATCH_FIELD(field, (k_scale_by_powers<P><<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(
                                  v, n, d_pow_hi, d_pow_lo, lo_bits, scale, use_scale ? 1 : 0, index_bitrev ? 1 : 0,
                                  log_n)));
    return cudaGetLastError();
}

cudaError_t launch_quotient_pointwise(int field, const fr_t* a, const fr_t* b, const fr_t* c, fr_t* h, uint64_t n,
                                      fr_t zinv, cudaStream_t s) {
    ACG_DISPATCH_FIELD(field,
                       (k_quotient_pointwise<P><<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(a, b, c, h, n, zinv)));
    return cudaGetLastError();
}

cudaError_t launch_poly_combine(int field, const fr_t* polys, const fr_t* weight, uint32_t n_polys, uint64_t len,
                                fr_t* out, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    ACG_DISPATCH_FIELD(field,
                       (k_poly_combine<P><<<grid_for(len, 128, 148 * 16), 128, 0, s>>>(polys, weight, n_polys, len, out)));
    return cudaGetLastError();
}

cudaError_t launch_axpy2(int field, fr_t* h, const fr_t* a, const fr_t* b, fr_t d2, fr_t d1, uint64_t n,
                         cudaStream_t s) {
    ACG_DISPATCH_FIELD(field, (k_axpy2<P><<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(h, a, b, d2, d1, n)));
    return cudaGetLastError();
}

}  // namespace acg
