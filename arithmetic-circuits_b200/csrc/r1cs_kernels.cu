// K2 -- R1CS witness check  (A.w) o (B.w) - C.w == 0  over Fr, plus the small conversion kernels.
//
// Replaces, per row g, dotProduct (reference src/Circuit/Affine.hs:121-125) of the sparse A/B/C row
// with the witness and the predicate of verificationWitnessZk (src/QAP.hs:309-327) in its
// evaluation-domain form.  Everything on the device is in Montgomery form, so
// the test  a*b*R^-1 == c  is exactly  (A.w)(B.w) == (C.w)  on canonical residues.
//
// Two kernels compute the same thing:
//   k_r1cs_rowwise : thread per row, direct global loads.  Simple; also the path for rows too long to
//                    stage in shared memory (Split gates wider than a tile).
//   k_r1cs_tiled   : persistent CTAs; a producer warp streams each tile's CSR slices (values, columns,
//                    row pointers of A, B and C) into shared memory with TMA bulk copies
//                    (cp.async.bulk + mbarrier, SASS UBLKCP) through a 2-stage ring; 256 consumer
//                    threads then (P1) turn every +-1-coefficient entry into its term +-w[col] in place,
//                    thread-per-ENTRY so the witness gathers are independent and the work is balanced,
//                    and queue the general-coefficient entries, (P2) run the queued 256-bit Montgomery
//                    products densely -- one entry per lane, no divergence between coefficient kinds --
//                    and (P3) sum each row's terms thread-per-row and test a*b == c.
//                    HBM traffic is exactly one pass over the CSR arrays; the witness is gathered
//                    through L2.
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
__device__ __forceinline__ fr_t row_dot(const DevCsr& M, uint32_t row, const fr_t* __restrict__ w) {
    fr_t acc = fr_zero<P>();
    const uint32_t s = M.rowptr[row], e = M.rowptr[row + 1];
    for (uint32_t k = s; k < e; ++k) {
        const fr_t v = M.val[k];
        const fr_t x = w[M.col[k]];
        if (fr_is_one<P>(v)) {
            acc = fr_add<P>(acc, x);
        } else if (fr_is_minus_one<P>(v)) {
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
            const fr_t a = row_dot<P>(m.m[0], row, w);
            const fr_t b = row_dot<P>(m.m[1], row, w);
            const fr_t c = row_dot<P>(m.m[2], row, w);
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
// tiled kernel
// ------------------------------------------------------------------------------------------------
namespace tiled {
constexpr uint32_t kConsumers = kTileRows;                 // 256
constexpr uint32_t kPool = kTilePoolEntries;               // entries
constexpr uint32_t kColsCap = kPool + 24;                  // 3 chunks, each <= ne + 6 after 16-byte alignment
constexpr uint32_t kRpCap = kTileRows + 8;                 // per matrix, (nrows + 1) rounded up to 4
constexpr uint32_t kOffVals = 0;
constexpr uint32_t kOffCols = kOffVals + kPool * 32;
constexpr uint32_t kOffRp = kOffCols + kColsCap * 4;
constexpr uint32_t kOffWork = kOffRp + 3 * kRpCap * 4;
constexpr uint32_t kOffDesc = kOffWork + kPool * 2;
constexpr uint32_t kStageBytes = ((kOffDesc + 32 + 127) / 128) * 128;
static_assert(kOffCols % 16 == 0 && kOffRp % 16 == 0 && kOffWork % 16 == 0 && kOffDesc % 16 == 0, "align");
static_assert((kRpCap * 4) % 16 == 0, "align");
constexpr uint32_t kUnroll = 3;

__device__ __forceinline__ uint32_t round_up4(uint32_t x) {
    return (x + 3u) & ~3u;
}
}  // namespace tiled

size_t r1cs_tiled_smem_bytes() {
    return (size_t)tiled::kStageBytes * kTileStages;
}

template <class P, bool EMIT>
__global__ void __launch_bounds__(kTiledThreads, kTiledCtasPerSm)
    k_r1cs_tiled(DevR1cs m, const fr_t* __restrict__ w, const Tile* __restrict__ tiles, uint32_t n_tiles,
                 uint64_t row_base, unsigned long long* __restrict__ result, fr_t* __restrict__ Aw,
                 fr_t* __restrict__ Bw, fr_t* __restrict__ Cw) {
    using namespace tiled;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kTileStages];
    __shared__ __align__(8) uint64_t empty_bar[kTileStages];
    __shared__ uint32_t s_nwork;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kTileStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        s_nwork = 0;
        mbar_fence_init();
    }
    __syncthreads();

    if (threadIdx.x >= kConsumers) {
        // ===== producer warp: one lane drives the TMA bulk copies =====
        if (threadIdx.x == kConsumers) {
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t stage = it % kTileStages;
                const uint32_t par = (it / kTileStages) & 1u;
                mbar_wait(&empty_bar[stage], par ^ 1u);
                uint8_t* sb = smem + (size_t)stage * kStageBytes;
                const Tile t = tiles[tile];
                *reinterpret_cast<Tile*>(sb + kOffDesc) = t;
                const uint32_t rp_bytes = round_up4(t.nrows + 1u) * 4u;
                uint32_t total = 3u * rp_bytes;
                uint32_t cb[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    cb[k] = t.ne[k] ? round_up4((t.e0[k] & 3u) + t.ne[k]) * 4u : 0u;
                    total += t.ne[k] * 32u + cb[k];
                }
                mbar_arrive_expect_tx(&full_bar[stage], total);
                uint32_t vstart = 0, coff = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (t.ne[k]) {
                        tma_load_1d(sb + kOffVals + (size_t)vstart * 32u, m.m[k].val + t.e0[k], t.ne[k] * 32u,
                                    &full_bar[stage]);
                        tma_load_1d(sb + kOffCols + (size_t)coff * 4u, m.m[k].col + (t.e0[k] & ~3u), cb[k],
                                    &full_bar[stage]);
                    }
                    tma_load_1d(sb + kOffRp + (size_t)k * kRpCap * 4u, m.m[k].rowptr + t.row0, rp_bytes,
                                &full_bar[stage]);
                    vstart += t.ne[k];
                    coff += cb[k] / 4u;
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t stage = it % kTileStages;
        const uint32_t par = (it / kTileStages) & 1u;
        mbar_wait(&full_bar[stage], par);
        uint8_t* sb = smem + (size_t)stage * kStageBytes;
        fr_t* vals = reinterpret_cast<fr_t*>(sb + kOffVals);
        const uint32_t* cols = reinterpret_cast<const uint32_t*>(sb + kOffCols);
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(sb + kOffRp);
        uint16_t* work = reinterpret_cast<uint16_t*>(sb + kOffWork);
        const Tile t = *reinterpret_cast<const Tile*>(sb + kOffDesc);

        const uint32_t nA = t.ne[0], nB = t.ne[1], nC = t.ne[2];
        const uint32_t nAB = nA + nB;
        const uint32_t E = nAB + nC;
        // column index of pool entry idx:  cols[idx + cadj[M]]
        const uint32_t c0 = t.e0[0] & 3u;
        const uint32_t off1 = nA ? round_up4(c0 + nA) : 0u;
        const uint32_t c1 = off1 + (t.e0[1] & 3u);
        const uint32_t off2 = off1 + (nB ? round_up4((t.e0[1] & 3u) + nB) : 0u);
        const uint32_t c2 = off2 + (t.e0[2] & 3u);
        const uint32_t cadjA = c0, cadjB = c1 - nA, cadjC = c2 - nAB;  // may wrap; used modulo 2^32

        // ---- P1: thread per entry.  +-1 coefficients: replace the value by +-w[col].  Others: queue.
        for (uint32_t base = 0; base < E; base += kConsumers * kUnroll) {
            uint32_t idx[kUnroll], col[kUnroll];
            bool plain[kUnroll], neg[kUnroll];
#pragma unroll
            for (uint32_t u = 0; u < kUnroll; ++u) {
                idx[u] = base + u * kConsumers + tid;
                const bool valid = idx[u] < E;
                bool gen = false;
                plain[u] = false;
                neg[u] = false;
                col[u] = 0;
                if (valid) {
                    const fr_t v = vals[idx[u]];
                    const uint32_t adj = idx[u] < nA ? cadjA : (idx[u] < nAB ? cadjB : cadjC);
                    col[u] = cols[idx[u] + adj];
                    const bool one = fr_is_one<P>(v);
                    neg[u] = fr_is_minus_one<P>(v);
                    plain[u] = one || neg[u];
                    gen = !plain[u];
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, gen);
                if (bal != 0u) {
                    const uint32_t leader = (uint32_t)__ffs(bal) - 1u;
                    uint32_t pos0 = 0;
                    if (lane == leader) pos0 = atomicAdd(&s_nwork, (uint32_t)__popc(bal));
                    pos0 = __shfl_sync(0xffffffffu, pos0, leader);
                    if (gen) work[pos0 + __popc(bal & lanemask_lt())] = (uint16_t)idx[u];
                }
            }
            fr_t x[kUnroll];
#pragma unroll
            for (uint32_t u = 0; u < kUnroll; ++u)
                if (plain[u]) x[u] = w[col[u]];
#pragma unroll
            for (uint32_t u = 0; u < kUnroll; ++u)
                if (plain[u]) vals[idx[u]] = neg[u] ? fr_neg<P>(x[u]) : x[u];
        }
        named_bar_sync(1, kConsumers);

        // ---- P2: dense Montgomery products of the queued entries, one per lane
        const uint32_t n_work = s_nwork;
        for (uint32_t i = tid; i < n_work; i += kConsumers) {
            const uint32_t e = work[i];
            const uint32_t adj = e < nA ? cadjA : (e < nAB ? cadjB : cadjC);
            const fr_t x = w[cols[e + adj]];
            const fr_t v = vals[e];
            vals[e] = fr_mul<P>(v, x);
        }
        named_bar_sync(1, kConsumers);
        if (tid == 0) s_nwork = 0;

        // ---- P3: thread per row: sum the terms, test a*b == c
        bool bad = false;
        if (tid < t.nrows) {
            fr_t abc[3];
            uint32_t vstart = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t* rpk = rp + k * kRpCap;
                const uint32_t s = rpk[tid] - t.e0[k] + vstart;
                const uint32_t e = rpk[tid + 1] - t.e0[k] + vstart;
                fr_t acc = fr_zero<P>();
                for (uint32_t j = s; j < e; ++j) {
                    const fr_t term = vals[j];
                    acc = fr_add<P>(acc, term);
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

        // the stage's shared memory was written through the generic proxy (P1/P2); order those writes
        // before the TMA (async proxy) refill, then hand the stage back to the producer
        fence_proxy_async_smem();
        named_bar_sync(1, kConsumers);
        if (tid == 0) mbar_arrive(&empty_bar[stage]);
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

template <class P, bool EMIT>
static cudaError_t launch_tiled_impl(const DevR1cs& m, const fr_t* w, const Tile* d_tiles, uint32_t n_tiles,
                                     uint64_t row_base, unsigned long long* d_result, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                                     int sm_count, cudaStream_t s) {
    const size_t smem = r1cs_tiled_smem_bytes();
    {
        cudaError_t e = cudaFuncSetAttribute(k_r1cs_tiled<P, EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    unsigned grid = (unsigned)(sm_count * kTiledCtasPerSm);
    if (grid > n_tiles) grid = n_tiles;
    k_r1cs_tiled<P, EMIT><<<grid, kTiledThreads, smem, s>>>(m, w, d_tiles, n_tiles, row_base, d_result, Aw, Bw, Cw);
    return cudaGetLastError();
}

cudaError_t launch_r1cs_tiled(int field, const DevR1cs& m, const fr_t* w, const Tile* d_tiles, uint32_t n_tiles,
                              uint64_t row_base, unsigned long long* d_result, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                              int sm_count, cudaStream_t s) {
    if (n_tiles == 0) return cudaSuccess;
    const bool emit = Aw || Bw || Cw;
    if (emit) {
        ACG_DISPATCH_FIELD(field, return (launch_tiled_impl<P, true>(m, w, d_tiles, n_tiles, row_base, d_result, Aw,
                                                                     Bw, Cw, sm_count, s)));
    } else {
        ACG_DISPATCH_FIELD(field, return (launch_tiled_impl<P, false>(m, w, d_tiles, n_tiles, row_base, d_result, Aw,
                                                                      Bw, Cw, sm_count, s)));
    }
    return cudaErrorInvalidValue;
}

}  // namespace acg
