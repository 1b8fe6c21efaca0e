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
//                    (P1)   thread per ENTRY: the witness element of every entry is gathered with
//                           cp.async (LDGSTS, no register staging, all gathers of the tile in flight at
//                           once) into two 16-byte planes (bank-conflict-free 128-bit accesses);
//                    (P2)   one lane per general entry: the 256-bit Montgomery product, in place --
//                           dense, no divergence between coefficient kinds;
//                    (P3)   thread per row: signed sum of the row's terms, then the test a*b == c.
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
// shared-memory layout of one CTA for geometry variant V
template <int V>
struct Cfg {
    static constexpr uint32_t kThreads = kTileGeom[V].threads;
    static constexpr uint32_t kPool = kTileGeom[V].pool;
    static constexpr uint32_t kMaxGen = kTileGeom[V].max_gen;
    static constexpr uint32_t kCtasPerSm = kTileGeom[V].ctas_per_sm;
    static constexpr uint32_t kColsCap = kPool + 24;   // 3 chunks, each <= ne + 6 after 16-byte alignment
    static constexpr uint32_t kRpCap = kThreads + 8;   // per matrix, (nrows + 1) rounded up to 4
    // gathered witness terms, later the per-entry products: two 16-byte planes so that 128-bit accesses of
    // neighbouring lanes fall on distinct banks
    static constexpr uint32_t kOffLo = 0;
    static constexpr uint32_t kOffHi = kPool * 16;
    static constexpr uint32_t kOffCols = kPool * 32;
    static constexpr uint32_t kOffRp = kOffCols + kColsCap * 4;
    static constexpr uint32_t kOffGval = align_up(kOffRp + 3 * kRpCap * 4, 32);
    static constexpr uint32_t kOffGlist = kOffGval + kMaxGen * 32;
    static constexpr uint32_t kOffDesc = align_up(kOffGlist + kMaxGen * 2, 16);
    static constexpr uint32_t kBytes = align_up(kOffDesc + 64, 128);
    static_assert(kOffCols % 16 == 0 && kOffRp % 16 == 0 && kOffGlist % 16 == 0 && (kRpCap * 4) % 16 == 0, "align");
};

__device__ __forceinline__ uint32_t round_up4(uint32_t x) {
    return (x + 3u) & ~3u;
}
__device__ __forceinline__ uint32_t round_up8(uint32_t x) {
    return (x + 7u) & ~7u;
}
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

// one thread: stage tile `tile` into `sb`, completion on `bar`
template <int V>
__device__ __forceinline__ void issue_tile_load(const DevR1cs& m, const Tile* __restrict__ tiles, uint32_t tile,
                                                uint8_t* sb, uint64_t* bar) {
    using C = Cfg<V>;
    const Tile t = tiles[tile];
    *reinterpret_cast<Tile*>(sb + C::kOffDesc) = t;
    const uint32_t rp_bytes = round_up4(t.nrows + 1u) * 4u;
    const uint32_t gl_bytes = round_up8(t.ng) * 2u;
    uint32_t total = 3u * rp_bytes + gl_bytes + t.ng * 32u;
    uint32_t cb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        cb[k] = t.ne[k] ? round_up4((t.e0[k] & 3u) + t.ne[k]) * 4u : 0u;
        total += cb[k];
    }
    mbar_arrive_expect_tx(bar, total);
    uint32_t coff = 0, goff = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (t.ne[k]) tma_load_1d(sb + C::kOffCols + (size_t)coff * 4u, m.m[k].col + (t.e0[k] & ~3u), cb[k], bar);
        if (t.ngv[k])
            tma_load_1d(sb + C::kOffGval + (size_t)goff * 32u, m.m[k].gval + t.gv0[k], t.ngv[k] * 32u, bar);
        tma_load_1d(sb + C::kOffRp + (size_t)k * C::kRpCap * 4u, m.m[k].rowptr + t.row0, rp_bytes, bar);
        coff += cb[k] / 4u;
        goff += t.ngv[k];
    }
    if (gl_bytes) tma_load_1d(sb + C::kOffGlist, m.glist + t.g0, gl_bytes, bar);
}
}  // namespace tiled

size_t r1cs_tiled_smem_bytes(int variant) {
    return variant == 0 ? tiled::Cfg<0>::kBytes : tiled::Cfg<1>::kBytes;
}

template <class P, bool EMIT, int V>
__global__ void __launch_bounds__(kTileGeom[V].threads, kTileGeom[V].ctas_per_sm)
    k_r1cs_tiled(DevR1cs m, const fr_t* __restrict__ w, const Tile* __restrict__ tiles, uint32_t n_tiles,
                 uint64_t row_base, unsigned long long* __restrict__ result, fr_t* __restrict__ Aw,
                 fr_t* __restrict__ Bw, fr_t* __restrict__ Cw) {
    using namespace tiled;
    using C = Cfg<V>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    if (tid == 0) {
        mbar_init(&full_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && blockIdx.x < n_tiles) issue_tile_load<V>(m, tiles, blockIdx.x, smem, &full_bar);

    uint4* lo = reinterpret_cast<uint4*>(smem + C::kOffLo);
    uint4* hi = reinterpret_cast<uint4*>(smem + C::kOffHi);
    const uint32_t* cols = reinterpret_cast<const uint32_t*>(smem + C::kOffCols);
    const uint32_t* rp = reinterpret_cast<const uint32_t*>(smem + C::kOffRp);
    const fr_t* gval = reinterpret_cast<const fr_t*>(smem + C::kOffGval);
    const uint16_t* glist = reinterpret_cast<const uint16_t*>(smem + C::kOffGlist);

    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(&full_bar, it & 1u);
        const Tile t = *reinterpret_cast<const Tile*>(smem + C::kOffDesc);

        const uint32_t nA = t.ne[0], nB = t.ne[1], nC = t.ne[2];
        const uint32_t nAB = nA + nB;
        const uint32_t E = nAB + nC;
        // column word of pool entry idx:  cols[idx + cadj[M]]
        const uint32_t c0 = t.e0[0] & 3u;
        const uint32_t off1 = nA ? round_up4(c0 + nA) : 0u;
        const uint32_t c1 = off1 + (t.e0[1] & 3u);
        const uint32_t off2 = off1 + (nB ? round_up4((t.e0[1] & 3u) + nB) : 0u);
        const uint32_t c2 = off2 + (t.e0[2] & 3u);
        const uint32_t cadj[3] = {c0, c1 - nA, c2 - nAB};  // may wrap; used modulo 2^32

        // ---- P1: gather the witness element of every entry, asynchronously, into the term planes
        for (uint32_t e = tid; e < E; e += C::kThreads) {
            const uint32_t c = cols[e + (e < nA ? cadj[0] : (e < nAB ? cadj[1] : cadj[2]))];
            cp_async_fr_planes(lo + e, hi + e, w + (c & kColMask));
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- P2: dense Montgomery products of the general entries, one per lane, in place
        for (uint32_t j = tid; j < t.ng; j += C::kThreads) {
            const uint32_t e = glist[j];
            const fr_t v = gval[j];
            const fr_t x = load_planes(lo, hi, e);
            store_planes(lo, hi, e, fr_mul<P>(v, x));
        }
        __syncthreads();

        // ---- P3: thread per row: signed sum of the terms, test a*b == c
        bool bad = false;
        if (tid < t.nrows) {
            fr_t abc[3];
            uint32_t vstart = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t* rpk = rp + k * C::kRpCap;
                const uint32_t s = rpk[tid] - t.e0[k] + vstart;
                const uint32_t e = rpk[tid + 1] - t.e0[k] + vstart;
                fr_t acc = fr_zero<P>();
                for (uint32_t j = s; j < e; ++j) {
                    const fr_t term = load_planes(lo, hi, j);
                    const uint32_t tag = cols[j + cadj[k]] >> 30;
                    if (tag == kTagMinusOne) {
                        acc = fr_sub<P>(acc, term);
                    } else {
                        acc = fr_add<P>(acc, term);
                    }
                }
                abc[k] = acc;
                vstart += t.ne[k];
            }
            if (EMIT) {
                const uint32_t row = t.row0 + tid;
                if (Aw) Aw[row] = abc[0];
                if (Bw) Bw[row] = abc[1];
                if (Cw) Cw[row] = abc[2];
            }
            bad = !fr_eq(fr_mul<P>(abc[0], abc[1]), abc[2]);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, bad);
        if (bal != 0u && lane == 0u) report_bad_rows(result, bal, row_base + t.row0 + (tid & ~31u));

        // the staged arrays were read through the generic proxy; order that before the next TMA (async
        // proxy) refill of the same bytes, then let one thread issue it
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < n_tiles) issue_tile_load<V>(m, tiles, tile + gridDim.x, smem, &full_bar);
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

template <class P, bool EMIT, int V>
static cudaError_t launch_tiled_impl(const DevR1cs& m, const fr_t* w, const Tile* d_tiles, uint32_t n_tiles,
                                     uint64_t row_base, unsigned long long* d_result, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                                     int sm_count, cudaStream_t s) {
    const size_t smem = tiled::Cfg<V>::kBytes;
    {
        cudaError_t e = cudaFuncSetAttribute(k_r1cs_tiled<P, EMIT, V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    unsigned grid = (unsigned)(sm_count * (int)kTileGeom[V].ctas_per_sm);
    if (grid > n_tiles) grid = n_tiles;
    k_r1cs_tiled<P, EMIT, V><<<grid, kTileGeom[V].threads, smem, s>>>(m, w, d_tiles, n_tiles, row_base, d_result, Aw,
                                                                       Bw, Cw);
    return cudaGetLastError();
}

cudaError_t launch_r1cs_tiled(int field, const DevR1cs& m, const fr_t* w, const Tile* d_tiles, uint32_t n_tiles,
                              uint64_t row_base, unsigned long long* d_result, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                              int sm_count, int variant, cudaStream_t s) {
    if (n_tiles == 0) return cudaSuccess;
    if (!m.tagged || !m.glist) return cudaErrorInvalidValue;
    const bool emit = Aw || Bw || Cw;
#define ACG_TILED(EMITV, VAR)                                                                                     \
    ACG_DISPATCH_FIELD(field, return (launch_tiled_impl<P, EMITV, VAR>(m, w, d_tiles, n_tiles, row_base, d_result, \
                                                                       Aw, Bw, Cw, sm_count, s)))
    if (variant == 1) {
        if (emit) ACG_TILED(true, 1);
        ACG_TILED(false, 1);
    } else {
        if (emit) ACG_TILED(true, 0);
        ACG_TILED(false, 0);
    }
#undef ACG_TILED
    return cudaErrorInvalidValue;
}

cudaError_t launch_validate_csr(const uint32_t* rowptr, const uint32_t* col, uint32_t n_rows, uint64_t nnz,
                                uint32_t n_cols, int* d_flag, cudaStream_t s) {
    const uint64_t work = nnz > n_rows ? nnz : n_rows;
    k_validate_csr<<<grid_for(work ? work : 1, 256, 148 * 8), 256, 0, s>>>(rowptr, col, n_rows, nnz, n_cols, d_flag);
    return cudaGetLastError();
}

}  // namespace acg
