// K2 -- R1CS witness check  (A.w) o (B.w) - C.w == 0  over Fr, plus the small conversion kernels.
//
// Replaces, per row g, dotProduct (reference src/Circuit/Affine.hs:121-125) of the sparse A/B/C row
// with the witness and the predicate of verificationWitnessZk (src/QAP.hs:309-327) in its
// evaluation-domain form.  Everything on the device is in Montgomery form, so
// the test  a*b*R^-1 == c  is exactly  (A.w)(B.w) == (C.w)  on canonical residues.
//
// Two kernels compute the same thing:
//   k_r1cs_rowwise : thread per row, direct global loads.  Used for one-shot checks (no preprocessing)
//                    and for rows too long to stage in shared memory.
//   k_r1cs_tiled   : persistent CTAs over tiles of <= 128 (or 256) rows built at upload time.  Per tile:
//                    (load) one thread issues TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) of
//                           the tile's slices -- tagged columns and row pointers of A, B, C, the values
//                           of the general-coefficient entries and their static index list -- into
//                           shared memory; several CTAs per SM hide each other's load latency;
//                           plus a second bulk copy of the tile's WITNESS WINDOW, the contiguous slice of
//                           w most references of the tile fall into;
//                    (P2)   one lane per general entry: the 256-bit Montgomery product -- dense, no
//                           divergence between coefficient kinds;
//                    (P3)   thread per row over the tile's ELL (slot-major, padded) entry words: the three
//                           sums A.w, B.w, C.w advance together with warp-uniform control flow; +-1 terms
//                           are gathered from the witness straight into registers, general terms come
//                           from the product planes (one generic-address load path); then a*b == c.
//                    Shared memory holds only what is reused (blob + products), so occupancy is bounded
//                    by registers, not by staging 32 bytes per entry.
//                    The coefficient classification (+1 / -1 / general) lives in two tag bits of the
//                    column word and is computed once at upload: the sparsity pattern is static, and the
//                    32-byte encodings of +-1 never need to be re-read.  HBM traffic per check is one
//                    pass over columns, row pointers and general values; the witness is gathered via L2.
#include "dev.cuh"
#include "kernels.h"

namespace acg {

// ------------------------------------------------------------------------------------------------
// conversions and field self-test kernels
// ------------------------------------------------------------------------------------------------
template <class P>
__global__ void k_to_mont(fr_t* __restrict__ v, uint64_t n, int* __restrict__ bad_flag) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        fr_t x = v[i];
        if (!fr_is_canonical<P>(x)) {
            *bad_flag = 1;  // benign race: every writer stores 1
        } else {
            v[i] = fr_to_mont<P>(x);
        }
    }
}
template <class P>
__global__ void k_from_mont(fr_t* __restrict__ v, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        fr_t x = v[i];
        v[i] = fr_from_mont<P>(x);
    }
}
// a, b, o in Montgomery form.  op: 0 add 1 sub 2 mul 3 inv(a) (inv 0 = 0)
template <class P>
__global__ void k_fr_binop(int op, const fr_t* __restrict__ a, const fr_t* __restrict__ b, fr_t* __restrict__ o,
                           uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        fr_t x = a[i], y = b[i], z;
        switch (op) {
            case 0: z = fr_add<P>(x, y); break;
            case 1: z = fr_sub<P>(x, y); break;
            case 2: z = fr_mul<P>(x, y); break;
            default: z = fr_inv<P>(x); break;
        }
        o[i] = z;
    }
}
__global__ void k_init_result(unsigned long long* r) {
    r[0] = 0ull;
    r[1] = ~0ull;
}

// ------------------------------------------------------------------------------------------------
// row-wise kernel
// ------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ fr_t row_dot(const DevCsr& M, uint32_t row, const fr_t* __restrict__ w, bool tagged) {
    fr_t acc = fr_zero<P>();
    const uint32_t s = M.rowptr[row], e = M.rowptr[row + 1];
    for (uint32_t k = s; k < e; ++k) {
        const uint32_t c = M.col[k];
        const fr_t x = w[c & kColMask];
        uint32_t tag = c >> 30;
        fr_t v;
        if (!tagged || tag == kTagGeneral) v = M.val[k];
        if (!tagged) tag = fr_is_one<P>(v) ? kTagPlusOne : (fr_is_minus_one<P>(v) ? kTagMinusOne : kTagGeneral);
        if (tag == kTagPlusOne) {
            acc = fr_add<P>(acc, x);
        } else if (tag == kTagMinusOne) {
            acc = fr_sub<P>(acc, x);
        } else {
            acc = fr_add<P>(acc, fr_mul<P>(v, x));
        }
    }
    return acc;
}

template <class P, bool EMIT>
__global__ void __launch_bounds__(256) k_r1cs_rowwise(DevR1cs m, const fr_t* __restrict__ w, uint32_t row_lo,
                                                      uint32_t row_hi, uint64_t row_base,
                                                      unsigned long long* __restrict__ result, fr_t* __restrict__ Aw,
                                                      fr_t* __restrict__ Bw, fr_t* __restrict__ Cw) {
    // rows are handed out in warp-sized groups so the ballot below is warp-uniform
    const uint32_t n = row_hi - row_lo;
    const uint32_t n_groups = (n + 31u) / 32u;
    const uint32_t warps_per_block = blockDim.x / 32u;
    for (uint32_t g = blockIdx.x * warps_per_block + threadIdx.x / 32u; g < n_groups; g += gridDim.x * warps_per_block) {
        const uint32_t row = row_lo + g * 32u + lane_id();
        bool bad = false;
        if (row < row_hi) {
            const fr_t a = row_dot<P>(m.m[0], row, w, m.tagged != 0);
            const fr_t b = row_dot<P>(m.m[1], row, w, m.tagged != 0);
            const fr_t c = row_dot<P>(m.m[2], row, w, m.tagged != 0);
            if (EMIT) {
                if (Aw) Aw[row] = a;
                if (Bw) Bw[row] = b;
                if (Cw) Cw[row] = c;
            }
            bad = !fr_eq(fr_mul<P>(a, b), c);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, bad);
        if (bal != 0u && lane_id() == 0u) report_bad_rows(result, bal, row_base + row_lo + (uint64_t)g * 32u);
    }
}

// ------------------------------------------------------------------------------------------------
// structural validation (one-shot path: the host does not walk the arrays)
// ------------------------------------------------------------------------------------------------
__global__ void k_validate_csr(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, uint32_t n_rows,
                               uint64_t nnz, uint32_t n_cols, int* __restrict__ flag) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t t0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    bool bad = false;
    for (uint64_t r = t0; r < n_rows; r += stride) bad |= rowptr[r] > rowptr[r + 1];
    if (t0 == 0) bad |= rowptr[0] != 0u || (uint64_t)rowptr[n_rows] != nnz;
    for (uint64_t e = t0; e < nnz; e += stride) bad |= col[e] >= n_cols;
    if (bad) atomicOr(flag, 2);
}

// ------------------------------------------------------------------------------------------------
// tiled kernel
// ------------------------------------------------------------------------------------------------
namespace tiled {
constexpr uint32_t align_up(uint32_t x, uint32_t a) {
    return (x + a - 1) / a * a;
}
// shared-memory layout of one CTA: STAGES x (blob buffer + witness window), then the products of the
// general-coefficient entries (+ the zero slot padding words point at)
template <int V, int STAGES>
struct Cfg {
    static constexpr uint32_t kThreads = kTileGeom[V].threads;
    static constexpr uint32_t kMaxGen = kTileGeom[V].max_gen;
    static constexpr uint32_t kBlobCap = align_up(tile_blob_capacity(kTileGeom[V]), 128);
    static constexpr uint32_t kWinBytes = kTileGeom[V].window * 32;
    static constexpr uint32_t kStageBytes = kBlobCap + kWinBytes;
    static constexpr uint32_t kOffProd = kStageBytes * STAGES;
    static constexpr uint32_t kBytes = kOffProd + (kMaxGen + 1) * 32;
    // resident CTAs per SM: bounded by shared memory (227 KB, 1 KB per CTA reserved), by 64 registers per
    // thread (1024 threads) and by the hardware limit of 32
    static constexpr uint32_t kCtasBySmem = (227u * 1024u) / (kBytes + 1024u + 64u);
    static constexpr uint32_t kCtasByRegs = 1024u / kThreads;
    static constexpr uint32_t kCtasPerSm = kCtasBySmem < kCtasByRegs ? kCtasBySmem : kCtasByRegs;
};

// 2 x 16-byte asynchronous gather global -> shared (LDGSTS), no register staging
__device__ __forceinline__ void cp_async_fr_planes(uint4* lo, uint4* hi, const fr_t* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(lo)), "l"(gmem_src) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(hi)),
                 "l"(reinterpret_cast<const uint8_t*>(gmem_src) + 16)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ fr_t load_planes(const uint4* lo, const uint4* hi, uint32_t i) {
    const uint4 a = lo[i], b = hi[i];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_planes(uint4* lo, uint4* hi, uint32_t i, const fr_t& v) {
    lo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ fr_t load_fr16(const uint8_t* p) {  // 16-byte aligned shared-memory element
    const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
// p - x without the zero fix-up: the result is in (0, p] and only ever feeds fr_add, which accepts it
template <class P>
__device__ __forceinline__ fr_t neg_lazy(const fr_t& x) {
    fr_t r;
    r.l[0] = ptx::sub_cc(P::p(0), x.l[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.l[i] = ptx::subc_cc(P::p(i), x.l[i]);
    r.l[7] = ptx::subc(P::p(7), x.l[7]);
    return r;
}
// one thread: bulk copies of the tile blob and of the tile's witness window
__device__ __forceinline__ void issue_blob_load(const DevTileStream& ts, const fr_t* __restrict__ w, uint32_t tile,
                                                uint8_t* dst, uint32_t blob_cap, uint64_t* bar) {
    const uint32_t o0 = ts.offsets[tile], o1 = ts.offsets[tile + 1];
    const uint32_t bytes = (o1 - o0) * 16u;
    const uint2 win = ts.windows[tile];  // {win_lo, win_n}
    mbar_arrive_expect_tx(bar, bytes + win.y * 32u);
    tma_load_1d(dst, ts.blobs + (size_t)o0 * 16u, bytes, bar);
    if (win.y) tma_load_1d(dst + blob_cap, w + win.x, win.y * 32u, bar);
}
}  // namespace tiled

template <class P, bool EMIT, int V, int STAGES>
__global__ void __launch_bounds__(kTileGeom[V].threads, tiled::Cfg<V, STAGES>::kCtasPerSm)
    k_r1cs_tiled(DevTileStream ts, const fr_t* __restrict__ w, uint64_t row_base,
                 unsigned long long* __restrict__ result, fr_t* __restrict__ Aw, fr_t* __restrict__ Bw,
                 fr_t* __restrict__ Cw) {
    using namespace tiled;
    using C = Cfg<V, STAGES>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t n_tiles = ts.n_tiles;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && blockIdx.x < n_tiles) issue_blob_load(ts, w, blockIdx.x, smem, C::kBlobCap, &full_bar[0]);

    uint4* prod = reinterpret_cast<uint4*>(smem + C::kOffProd);  // 32-byte product slots

    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t stage = it % STAGES;
        const uint8_t* blob = smem + (size_t)stage * C::kStageBytes;
        const uint4* win = reinterpret_cast<const uint4*>(blob + C::kBlobCap);
        if (STAGES == 2 && tid == 0 && tile + gridDim.x < n_tiles)  // prefetch: that buffer was released by the
            issue_blob_load(ts, w, tile + gridDim.x,                // barrier that ended the previous tile
                            smem + (size_t)((it + 1) % STAGES) * C::kStageBytes, C::kBlobCap,
                            &full_bar[(it + 1) % STAGES]);
        mbar_wait(&full_bar[stage], (it / STAGES) & 1u);

        const TileHeader h = *reinterpret_cast<const TileHeader*>(blob);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(blob + h.off_words);
        const uint32_t* gcol = reinterpret_cast<const uint32_t*>(blob + h.off_gcol);
        const uint8_t* gval = blob + h.off_gval;

        // ---- P2: dense 256-bit Montgomery products, one general entry per lane: operand from the witness
        //          window (shared memory) or one 256-bit global load
        for (uint32_t j = tid; j < h.n_general; j += C::kThreads) {
            const uint32_t gw = gcol[j];
            const uint4* gsrc = reinterpret_cast<const uint4*>(w + (gw & kColMask));
            const uint4* src = (gw & kWinFlag) ? win + 2u * (gw & kColMask) : gsrc;
            const uint4 a4 = src[0], b4 = src[1];
            fr_t x;
            x.l[0] = a4.x; x.l[1] = a4.y; x.l[2] = a4.z; x.l[3] = a4.w;
            x.l[4] = b4.x; x.l[5] = b4.y; x.l[6] = b4.z; x.l[7] = b4.w;
            const fr_t pr = fr_mul<P>(load_fr16(gval + (size_t)j * 32u), x);
            prod[2u * j] = make_uint4(pr.l[0], pr.l[1], pr.l[2], pr.l[3]);
            prod[2u * j + 1u] = make_uint4(pr.l[4], pr.l[5], pr.l[6], pr.l[7]);
        }
        if (tid == 0) {  // the zero slot that padding words reference
            prod[2u * h.n_general] = make_uint4(0u, 0u, 0u, 0u);
            prod[2u * h.n_general + 1u] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();

        // ---- P3: thread per row, warp-uniform: the sums A.w, B.w, C.w advance together slot by slot.  A term
        //          is loaded through one generic address that points into the witness window or the product
        //          slots (shared memory) or, for far references, into the witness in global memory; a -1
        //          coefficient negates under a predicate.
        bool bad = false;
        if (tid < h.nrows) {
            auto term = [&](uint32_t word) -> fr_t {
                const uint32_t idx = word & kColMask;
                const uint4* src = reinterpret_cast<const uint4*>(w + idx);
                if (word & kWinFlag) src = win + 2u * idx;
                if (word >> 31) src = prod + 2u * idx;  // tags 2 (general) and 3 (padding)
                const uint4 a4 = src[0], b4 = src[1];
                fr_t x;
                x.l[0] = a4.x; x.l[1] = a4.y; x.l[2] = a4.z; x.l[3] = a4.w;
                x.l[4] = b4.x; x.l[5] = b4.y; x.l[6] = b4.z; x.l[7] = b4.w;
                if ((word >> 30) == kTagMinusOne) x = neg_lazy<P>(x);
                return x;
            };
            const uint32_t wA = h.width[0], wB = h.width[1], wC = h.width[2];
            const uint32_t* pa = words + tid;
            const uint32_t* pb = pa + wA * h.nrows;
            const uint32_t* pc = pb + wB * h.nrows;
            const uint32_t wmax = max(wA, max(wB, wC));
            fr_t a = fr_zero<P>(), b = fr_zero<P>(), c = fr_zero<P>();
            for (uint32_t j = 0; j < wmax; ++j) {  // trip count and the three predicates are warp-uniform
                const bool ua = j < wA, ub = j < wB, uc = j < wC;
                fr_t ta, tb, tc;
                if (ua) ta = term(pa[j * h.nrows]);
                if (ub) tb = term(pb[j * h.nrows]);
                if (uc) tc = term(pc[j * h.nrows]);
                if (ua) a = fr_add<P>(a, ta);
                if (ub) b = fr_add<P>(b, tb);
                if (uc) c = fr_add<P>(c, tc);
            }
            if (EMIT) {
                const uint32_t row = h.row0 + tid;
                if (Aw) Aw[row] = a;
                if (Bw) Bw[row] = b;
                if (Cw) Cw[row] = c;
            }
            bad = !fr_eq(fr_mul<P>(a, b), c);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, bad);
        if (bal != 0u && lane == 0u) report_bad_rows(result, bal, row_base + h.row0 + (tid & ~31u));

        // the blob was read through the generic proxy; order that before the next TMA (async proxy) refill
        fence_proxy_async_smem();
        __syncthreads();
        if (STAGES == 1 && tid == 0 && tile + gridDim.x < n_tiles)
            issue_blob_load(ts, w, tile + gridDim.x, smem, C::kBlobCap, &full_bar[0]);
    }
}

template <class P>
__global__ void k_to_mont_scattered(uint8_t* __restrict__ blobs, const uint32_t* __restrict__ offs, uint64_t n,
                                    int* __restrict__ bad_flag) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4* p = reinterpret_cast<uint4*>(blobs + (size_t)offs[i] * 16u);
        const uint4 a = p[0], b = p[1];
        fr_t x;
        x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
        x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
        if (!fr_is_canonical<P>(x)) {
            *bad_flag = 1;
        } else {
            const fr_t y = fr_to_mont<P>(x);
            p[0] = make_uint4(y.l[0], y.l[1], y.l[2], y.l[3]);
            p[1] = make_uint4(y.l[4], y.l[5], y.l[6], y.l[7]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
#define ACG_DISPATCH_FIELD(field, EXPR)                    \
    do {                                                   \
        if ((field) == 0) {                                \
            using P = Bn254Fr;                             \
            EXPR;                                          \
        } else if ((field) == 1) {                         \
            using P = Bls12381Fr;                          \
            EXPR;                                          \
        } else {                                           \
            return cudaErrorInvalidValue;                  \
        }                                                  \
    } while (0)

static inline unsigned grid_for(uint64_t n, unsigned block, unsigned max_blocks) {
    uint64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (unsigned)g;
}

cudaError_t launch_to_mont(int field, fr_t* v, uint64_t n, int* d_bad_flag, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    ACG_DISPATCH_FIELD(field, (k_to_mont<P><<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(v, n, d_bad_flag)));
    return cudaGetLastError();
}
cudaError_t launch_from_mont(int field, fr_t* v, uint64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    ACG_DISPATCH_FIELD(field, (k_from_mont<P><<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(v, n)));
    return cudaGetLastError();
}
cudaError_t launch_fr_binop(int field, int op, const fr_t* a, const fr_t* b, fr_t* o, uint64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    ACG_DISPATCH_FIELD(field, (k_fr_binop<P><<<grid_for(n, 128, 148 * 16), 128, 0, s>>>(op, a, b, o, n)));
    return cudaGetLastError();
}
cudaError_t launch_init_result(unsigned long long* d_result, cudaStream_t s) {
    k_init_result<<<1, 1, 0, s>>>(d_result);
    return cudaGetLastError();
}

cudaError_t launch_r1cs_rowwise(int field, const DevR1cs& m, const fr_t* w, uint32_t row_lo, uint32_t row_hi,
                                uint64_t row_base, unsigned long long* d_result, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                                cudaStream_t s) {
    if (row_hi <= row_lo) return cudaSuccess;
    const bool emit = Aw || Bw || Cw;
    const unsigned grid = grid_for(row_hi - row_lo, 256, 148 * 64);
    if (emit) {
        ACG_DISPATCH_FIELD(field, (k_r1cs_rowwise<P, true><<<grid, 256, 0, s>>>(m, w, row_lo, row_hi, row_base,
                                                                                 d_result, Aw, Bw, Cw)));
    } else {
        ACG_DISPATCH_FIELD(field, (k_r1cs_rowwise<P, false><<<grid, 256, 0, s>>>(m, w, row_lo, row_hi, row_base,
                                                                                  d_result, Aw, Bw, Cw)));
    }
    return cudaGetLastError();
}

template <class P, bool EMIT, int V, int STAGES>
static cudaError_t launch_tiled_impl(const DevTileStream& ts, const fr_t* w, uint64_t row_base,
                                     unsigned long long* d_result, fr_t* Aw, fr_t* Bw, fr_t* Cw, int sm_count,
                                     cudaStream_t s) {
    using C = tiled::Cfg<V, STAGES>;
    {
        cudaError_t e = cudaFuncSetAttribute(k_r1cs_tiled<P, EMIT, V, STAGES>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kBytes);
        if (e != cudaSuccess) return e;
    }
    unsigned grid = (unsigned)(sm_count * (int)C::kCtasPerSm);
    if (grid > ts.n_tiles) grid = ts.n_tiles;
    k_r1cs_tiled<P, EMIT, V, STAGES><<<grid, kTileGeom[V].threads, C::kBytes, s>>>(ts, w, row_base, d_result, Aw, Bw,
                                                                                   Cw);
    return cudaGetLastError();
}

cudaError_t launch_r1cs_tiled(int field, const DevTileStream& ts, const fr_t* w, uint64_t row_base,
                              unsigned long long* d_result, fr_t* Aw, fr_t* Bw, fr_t* Cw, int sm_count, int stages,
                              cudaStream_t s) {
    if (ts.n_tiles == 0) return cudaSuccess;
    const bool emit = Aw || Bw || Cw;
#define ACG_TILED(EMITV, VAR, STG)                                                                               \
    ACG_DISPATCH_FIELD(field, return (launch_tiled_impl<P, EMITV, VAR, STG>(ts, w, row_base, d_result, Aw, Bw, Cw, \
                                                                            sm_count, s)))
#define ACG_TILED_VS(VAR, STG)        \
    do {                              \
        if (emit) ACG_TILED(true, VAR, STG); \
        ACG_TILED(false, VAR, STG);   \
    } while (0)
    switch (ts.variant) {
        case 1:
            if (stages == 2) ACG_TILED_VS(1, 2);
            ACG_TILED_VS(1, 1);
            break;
        case 2:
            if (stages == 2) ACG_TILED_VS(2, 2);
            ACG_TILED_VS(2, 1);
            break;
        case 3:
            if (stages == 2) ACG_TILED_VS(3, 2);
            ACG_TILED_VS(3, 1);
            break;
        default:
            if (stages == 2) ACG_TILED_VS(0, 2);
            ACG_TILED_VS(0, 1);
            break;
    }
#undef ACG_TILED_VS
#undef ACG_TILED
    return cudaErrorInvalidValue;
}

cudaError_t launch_to_mont_scattered(int field, uint8_t* blobs, const uint32_t* offs, uint64_t n, int* d_bad_flag,
                                     cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    ACG_DISPATCH_FIELD(field,
                       (k_to_mont_scattered<P><<<grid_for(n, 256, 148 * 16), 256, 0, s>>>(blobs, offs, n, d_bad_flag)));
    return cudaGetLastError();
}

cudaError_t launch_validate_csr(const uint32_t* rowptr, const uint32_t* col, uint32_t n_rows, uint64_t nnz,
                                uint32_t n_cols, int* d_flag, cudaStream_t s) {
    const uint64_t work = nnz > n_rows ? nnz : n_rows;
    k_validate_csr<<<grid_for(work ? work : 1, 256, 148 * 8), 256, 0, s>>>(rowptr, col, n_rows, nnz, n_cols, d_flag);
    return cudaGetLastError();
}

}  // namespace acg
