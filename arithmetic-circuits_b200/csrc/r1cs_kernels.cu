// K2 -- R1CS witness check  (A.w) o (B.w) - C.w == 0  over Fr, plus the small conversion kernels.
//
// Replaces, per row g, dotProduct (reference src/Circuit/Affine.hs:121-125) of the sparse A/B/C row
// with the witness and the predicate of verificationWitnessZk (src/QAP.hs:309-327) in its
// evaluation-domain form.  Everything on the device is in Montgomery form, so
// the test  a*b*R^-1 == c  is exactly  (A.w)(B.w) == (C.w)  on canonical residues.
//
// Three kernels compute the same thing:
//   k_r1cs_rowwise : thread per row, direct global loads.  Used for one-shot checks (no preprocessing).
//   k_r1cs_longrows: warp per row over a list of rows too long for a tile (Split gates); only for tile geometries
//                    without the LONG form of the tiled kernel, and for shards without tiles.
//   k_r1cs_tiled   : persistent CTAs, each walking a contiguous run of the tiles (<= 128 / 256 / 64 / 32 rows) that
//                    upload laid out as a stream of execution-ready blobs (kernels.h).  Per tile:
//                    (load) two TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP): the blob and the tile's
//                           WITNESS WINDOW, the contiguous slice of w most of its references fall into -- issued
//                           the moment the previous tile is done, from L2 (prefetched a tile earlier); every term
//                           then lives in one shared-memory array of 32-byte slots
//                           (window | far buffer 0 | far buffer 1 | products | zero);
//                    (far)  the distinct references outside the window were gathered during the previous tile
//                           (16-byte cp.async copies into the far buffer of this tile's parity, from the column
//                           list the previous blob carried); the gathers of the next tile start now;
//                    (P2)   one lane per general entry off the constant wire: the 256-bit Montgomery product --
//                           dense, no divergence between coefficient kinds;
//                    (P3)   thread per row over the tile's ELL (slot-major, padded) entry words: the three
//                           sums A.w, B.w, C.w advance together with warp-uniform control flow and touch
//                           shared memory only (word = sign | term address); then a*b == c.
//                    The coefficient classification (+1 / -1 / general / general on the constant wire) is computed
//                    once at upload: the sparsity pattern is static, and the 32-byte encodings of +-1 never need to
//                    be re-read.  HBM traffic per check is one pass over the blobs; the witness is read via L2.
//                    (long)  LONG instantiations: once a CTA has run out of tiles its warps claim the rows too long
//                           for a tile from a counter and check them on the plain CSR arrays (kernels.h DevLongRows).
// Hand-over of the result (kernels.h CheckEpilogue): a plain single-GPU check starts by writing {0, none} and opening
// a gate, violating warps update the result themselves, and nothing runs at the end; sharded and overlapped checks end
// with finish_check(): the last CTA of the last launch finalises the result pair and, for row shards on several GPUs,
// all-reduces it over peer memory (PeerSlots).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

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
// dst <- src unless *bad_flag is set: the commit step of a staged upload (a rejected update leaves dst untouched)
__global__ void k_copy_if_clean(uint4* __restrict__ dst, const uint4* __restrict__ src, uint64_t n16,
                                const int* __restrict__ bad_flag) {
    if (*bad_flag) return;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}
__global__ void k_init_result(unsigned long long* r) {
    r[0] = 0ull;
    r[1] = ~0ull;
}

// ------------------------------------------------------------------------------------------------
// all-reduce of the {violation count, first bad row} pair over peer memory (NVLink / NVSwitch P2P)
// ------------------------------------------------------------------------------------------------
// The only exchange of the row-sharded check is 16 bytes per rank per check.  Instead of a NCCL collective (a
// separate launch with ~15 us of latency next to a ~65 us kernel) every rank STORES its pair straight into every
// peer's exchange buffer -- plain system-scope stores through the NVLink peer mapping (CUDA IPC), data first, then a
// release store of the step's sequence number -- and then reads the world's pairs from its OWN buffer, spinning on
// the sequence numbers with acquire loads.  One warp, one lane per rank.  Buffers are double-buffered by sequence
// parity: a rank can publish step s + 1 while a slower peer still reads step s, and it cannot get two steps ahead
// because finishing s + 1 needs that peer's s + 1 publication, which follows its step-s read in stream order.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// one warp: (cnt, first) of this rank in, the world's {sum, min} out (valid on every lane); cnt = ~0 on a timeout
__device__ __forceinline__ void peer_allreduce_warp(const PeerSlots& ps, unsigned long long seq, unsigned long long& cnt,
                                                    unsigned long long& first) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t par = (uint32_t)(seq & 1ull);
    unsigned long long* base_lane = lane < ps.world ? ps.base[lane] : nullptr;
    unsigned long long* base_own = ps.own;
    if (lane < ps.world) {
        unsigned long long* dst = base_lane + (size_t)(par * kMaxPeers + ps.rank) * 4u;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(cnt) : "memory");
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(first) : "memory");
        st_release_sys(dst + 2, seq);
    }
    unsigned long long c = 0ull, f = ~0ull;
    bool timed_out = false;
    if (lane < ps.world) {
        const unsigned long long* src = base_own + (size_t)(par * kMaxPeers + lane) * 4u;
        const long long t0 = clock64();
        while (ld_acquire_sys(src + 2) != seq) {
            if (clock64() - t0 > 8000000000ll) {  // ~4 s: a peer never arrived; report instead of hanging the GPU
                timed_out = true;
                break;
            }
            __nanosleep(100);
        }
        c = src[0];
        f = src[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        const unsigned long long g = __shfl_xor_sync(0xffffffffu, f, o);
        f = g < f ? g : f;
    }
    const bool any_timeout = __any_sync(0xffffffffu, timed_out);
    cnt = any_timeout ? ~0ull : c;
    first = f;
}
// stand-alone form (a check that launched no kernel on this rank still has to take part)
__global__ void k_peer_allreduce(PeerSlots ps, unsigned long long seq, unsigned long long* __restrict__ result) {
    unsigned long long cnt = result[0], first = result[1];
    peer_allreduce_warp(ps, seq, cnt, first);
    if (threadIdx.x == 0) {
        result[0] = cnt;
        result[1] = first;
    }
}

// Direct hand-over (kernels.h CheckEpilogue), one thread of block 0: out <- what earlier launches of this check
// accumulated, scratch reset, gate opened.
__device__ __forceinline__ void open_gate(const CheckEpilogue& ep) {
    const unsigned long long cnt = __ldcg(ep.accum), first = __ldcg(ep.accum + 1);
    ep.out[0] = cnt;
    ep.out[1] = first;
    if (cnt != 0ull || first != ~0ull) {
        ep.accum[0] = 0ull;
        ep.accum[1] = ~0ull;
    }
    __threadfence();
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(ep.gate), "l"(ep.gate_seq) : "memory");
}
// One lane of a warp that found violated rows (rare, and kept out of line so that it costs the tile loop no
// registers): into `out` behind the gate (direct hand-over), else into the scratch pair.
__device__ __noinline__ void report_violations(unsigned long long* gate, unsigned long long gate_seq,
                                               unsigned long long* out, unsigned long long* accum, uint32_t overlap,
                                               unsigned long long n_bad, unsigned long long first) {
    if (gate != nullptr && out != nullptr) {
        const long long t0 = clock64();
        for (;;) {
            unsigned long long g;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(g) : "l"(gate) : "memory");
            if (g == gate_seq) break;
            // block 0 is dispatched before any other block of the grid, so this wait is microseconds; a word that
            // never matches means the ring was lapped (kGateRing later checks started while this one ran): fail loudly
            if (clock64() - t0 > 8000000000ll) __trap();
            __nanosleep(64);
        }
        atomicAdd(&out[0], n_bad);
        atomicMin(&out[1], first);
        return;
    }
    if (overlap) griddep_wait();  // the scratch pair belongs to the previous check until it completes
    atomicAdd(&accum[0], n_bad);
    atomicMin(&accum[1], first);
}

// End of a check kernel (every thread of every CTA calls it): see CheckEpilogue in kernels.h.
__device__ __forceinline__ void finish_check(const CheckEpilogue& ep) {
    if (ep.out == nullptr) return;  // an earlier launch of the same check: it only accumulates
    if (ep.gate != nullptr) return;  // direct hand-over (kernels.h CheckEpilogue): `out` is complete when the grid is
    if (ep.overlap) griddep_wait();  // the previous check has finalised (and reset) the scratch pair
    __shared__ uint32_t s_last;
    __syncthreads();  // every warp of this CTA has reported
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ep.ticket, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (s_last == 0u || threadIdx.x >= 32u) return;
    __threadfence();
    unsigned long long cnt = __ldcg(ep.accum), first = __ldcg(ep.accum + 1);
    if (threadIdx.x == 0) {  // scratch and ticket ready for the next check
        ep.accum[0] = 0ull;
        ep.accum[1] = ~0ull;
        *ep.ticket = 0u;
    }
    if (ep.peers.world > 1u) peer_allreduce_warp(ep.peers, ep.seq, cnt, first);
    if (threadIdx.x == 0) {
        ep.out[0] = cnt;
        ep.out[1] = first;
    }
}

cudaError_t launch_peer_allreduce(const PeerSlots& ps, unsigned long long seq, unsigned long long* d_result,
                                  cudaStream_t s) {
    k_peer_allreduce<<<1, 32, 0, s>>>(ps, seq, d_result);
    return cudaGetLastError();
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
                                                      uint32_t row_hi, uint64_t row_base, CheckEpilogue ep,
                                                      fr_t* __restrict__ Aw, fr_t* __restrict__ Bw,
                                                      fr_t* __restrict__ Cw) {
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
        if (bal != 0u && lane_id() == 0u) report_bad_rows(ep.accum, bal, row_base + row_lo + (uint64_t)g * 32u);
    }
    finish_check(ep);
}

// ------------------------------------------------------------------------------------------------
// long rows: warp per row over a LIST of rows
// ------------------------------------------------------------------------------------------------
// Rows too long for a tile -- row 0 of a Split gate carries one general coefficient 2^i per output bit, 256 of them in
// the reference's circuits (src/QAP.hs:443-473, test/Test/Circuit/Arithmetic.hs:112-120) -- are listed at upload and
// all handled by ONE launch: a warp per row, the lanes stride over the row's entries (coalesced column and value
// loads, 32 products in flight per row), the three partial sums are folded with warp shuffles, lane 0 tests a*b == c.
__device__ __forceinline__ fr_t shfl_xor_fr(const fr_t& x, int mask) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_xor_sync(0xffffffffu, x.l[i], mask);
    return r;
}
// The lanes take the row's entries kBatch at a time each (k, k + 32, ...): the column words are loaded first, then the
// witness elements and coefficient values -- 2 * kBatch independent 32-byte loads in flight per lane -- and only then
// the products: the kernel is bound by the latency of the two dependent loads (column -> witness element), so the
// loads of a batch must not wait for the arithmetic of the previous entry.
template <class P, int kBatch = 2>
__device__ __forceinline__ fr_t row_dot_warp(const DevCsr& M, uint32_t row, const fr_t* __restrict__ w, bool tagged) {
    fr_t acc = fr_zero<P>();
    const uint32_t s = M.rowptr[row], e = M.rowptr[row + 1];
    for (uint32_t k0 = s + lane_id(); k0 < e; k0 += 32u * kBatch) {
        uint32_t c[kBatch];
        fr_t x[kBatch], v[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) c[j] = k0 + 32u * j < e ? M.col[k0 + 32u * j] : kColMask + 1u;  // (tag 1, unused)
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            if (k0 + 32u * j < e) {
                x[j] = w[c[j] & kColMask];
                if (!tagged || (c[j] >> 30) == kTagGeneral) v[j] = M.val[k0 + 32u * j];
            }
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            if (k0 + 32u * j < e) {
                uint32_t tag = c[j] >> 30;
                if (!tagged) tag = fr_is_one<P>(v[j]) ? kTagPlusOne : (fr_is_minus_one<P>(v[j]) ? kTagMinusOne : kTagGeneral);
                if (tag == kTagPlusOne) {
                    acc = fr_add<P>(acc, x[j]);
                } else if (tag == kTagMinusOne) {
                    acc = fr_sub<P>(acc, x[j]);
                } else {
                    acc = fr_add<P>(acc, fr_mul<P>(v[j], x[j]));
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = fr_add<P>(acc, shfl_xor_fr(acc, o));
    return acc;  // on every lane
}
template <class P, bool EMIT>
__global__ void __launch_bounds__(128, 3) k_r1cs_longrows(DevR1cs m, const fr_t* __restrict__ w,
                                                       const uint32_t* __restrict__ rows, uint32_t n_rows,
                                                       uint64_t row_base, CheckEpilogue ep, fr_t* __restrict__ Aw,
                                                       fr_t* __restrict__ Bw, fr_t* __restrict__ Cw) {
    const uint32_t warps_per_block = blockDim.x / 32u;
    for (uint32_t i = blockIdx.x * warps_per_block + threadIdx.x / 32u; i < n_rows; i += gridDim.x * warps_per_block) {
        const uint32_t row = rows[i];
        const fr_t a = row_dot_warp<P>(m.m[0], row, w, m.tagged != 0);
        const fr_t b = row_dot_warp<P>(m.m[1], row, w, m.tagged != 0);
        const fr_t c = row_dot_warp<P>(m.m[2], row, w, m.tagged != 0);
        if (lane_id() == 0u) {
            if (EMIT) {
                if (Aw) Aw[row] = a;
                if (Bw) Bw[row] = b;
                if (Cw) Cw[row] = c;
            }
            if (!fr_eq(fr_mul<P>(a, b), c)) report_bad_rows(ep.accum, 1u, row_base + row);
        }
    }
    finish_check(ep);
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
// shared-memory layout of one CTA: the tile blob, then the term array (window | far | products | zero)
template <int V>
struct Cfg {
    static constexpr uint32_t kThreads = kTileGeom[V].threads;
    static constexpr uint32_t kWin = kTileGeom[V].window;
    static constexpr uint32_t kFar0 = kWin;                              // first far slot of even tiles
    static constexpr uint32_t kFarN = kTileGeom[V].max_far;              // odd tiles: kFar0 + kFarN
    static constexpr uint32_t kProd0 = tile_prod_slot0(kTileGeom[V]);    // first product slot
    static constexpr uint32_t kZero = tile_term_slots(kTileGeom[V]) - 1u;  // the zero slot
    static constexpr uint32_t kOffTerms = tile_terms_offset(kTileGeom[V]);
    static constexpr uint32_t kBytes = tile_smem_bytes(kTileGeom[V]);
    static constexpr uint32_t kCtasPerSm = tile_ctas_per_sm(kTileGeom[V]);
    static constexpr bool kSwizzle = kTileGeom[V].swizzle != 0;
    static constexpr bool kFarDouble = kTileGeom[V].far_bufs == 2u;
    static constexpr bool kProdInPlace = kTileGeom[V].prod_in_place != 0;
    // chunk (16-byte unit from the start of shared memory) of the low half of term slot `slot` outside the window
    static __device__ __forceinline__ uint32_t term_chunk(uint32_t slot) {
        const uint32_t c = kOffTerms / 16u + 2u * slot;
        return kSwizzle ? swz16(c) : c;
    }
    // .. and of the 32-byte value `j` of a blob section that starts `off` bytes (a multiple of 32) into the blob
    static __device__ __forceinline__ uint32_t blob_chunk(uint32_t off, uint32_t j) {
        const uint32_t c = off / 16u + 2u * j;
        return kSwizzle ? swz16(c) : c;
    }
};

// a term by the chunk of its low half (kernels.h swz16): the high half is the other chunk of the 32-byte unit
__device__ __forceinline__ fr_t load_term(const uint4* smem16, uint32_t chunk) {
    const uint4 a = smem16[chunk], b = smem16[chunk ^ 1u];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_term(uint4* smem16, uint32_t chunk, const fr_t& v) {
    smem16[chunk] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    smem16[chunk ^ 1u] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
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
// one thread: bulk copies of a tile blob and of the tile's witness window (into the first term slots)
__device__ __forceinline__ void issue_tile_load(const DevTileStream& ts, const fr_t* __restrict__ w, uint32_t blob_off16,
                                                uint32_t blob_bytes, uint32_t win_lo, uint32_t win_n,
                                                uint8_t* blob_dst, uint8_t* win_dst, uint64_t* bar,
                                                bool first_of_run = false) {
    mbar_arrive_expect_tx(bar, blob_bytes + win_n * 32u);
    if (first_of_run)
        tma_load_1d_evict_last(blob_dst, ts.blobs + (size_t)blob_off16 * 16u, blob_bytes, bar);
    else
        tma_load_1d_stream(blob_dst, ts.blobs + (size_t)blob_off16 * 16u, blob_bytes, bar);
    if (win_n) tma_load_1d_keep(win_dst, w + win_lo, win_n * 32u, bar);
}
// one thread: pull what most likely comes after the tile being loaded towards L2 -- the stream is linear per
// CTA and windows of successive tiles are (for circuits built gate by gate) successive slices of w, so the
// guess "as much again, right behind" is good; a wrong guess only wastes a prefetch
__device__ __forceinline__ void prefetch_behind(const DevTileStream& ts, const fr_t* __restrict__ w, uint32_t blob_off16,
                                                uint32_t blob_bytes, uint32_t win_lo, uint32_t win_n) {
    const uint32_t o = blob_off16 + blob_bytes / 16u;
    if (o < ts.blobs_len16) prefetch_l2_bulk_stream(ts.blobs + (size_t)o * 16u, min(blob_bytes, (ts.blobs_len16 - o) * 16u));
    const uint32_t c = win_lo + win_n;
    if (c < ts.n_cols) prefetch_l2_bulk(w + c, min(win_n, ts.n_cols - c) * 32u);
}
}  // namespace tiled

namespace tiled {
template <class P>
__device__ __forceinline__ fr_t signed_term(const uint4* terms, uint32_t word) {
    fr_t x = load_term(terms, word & ~kTermSign);
    if (word & kTermSign) x = neg_lazy<P>(x);
    return x;  // <= p
}
// a + b with no reduction (the caller bounds the sum below 2^256)
__device__ __forceinline__ fr_t add_wide(const fr_t& a, const fr_t& b) {
    fr_t t;
    t.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) t.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    t.l[7] = ptx::addc(a.l[7], b.l[7]);
    return t;
}
// how many values <= p add up below 2^256: 5 for BN254 Fr, 2 for BLS12-381 Fr
template <class P>
struct Lazy {
    static constexpr int kTerms = (int)(0x100000000ull / ((uint64_t)P::p(7) + 1ull));
};
// One accumulation step of a row sum.  LAZY: the running value may stay unreduced (it ends as the scalar
// operand of the Montgomery product, which accepts any 256-bit value); otherwise it stays <= p.
template <class P, bool LAZY>
__device__ __forceinline__ fr_t acc_step(const fr_t& acc, const fr_t& t) {
    return LAZY ? add_wide(acc, t) : fr_add<P>(acc, t);  // fr_add maps [0, 2p] -> [0, p]
}
// Sums of one row over the tile's ELL words (pa[j * nrows] = word of slot j), compile-time widths >= 1:
// straight-line, the three sums advance together slot by slot (three independent carry chains).
//   a: A-row . w, possibly unreduced (< 2^256, congruent mod p);  b in [0, p];  c in [0, p)
template <class P, int WA, int WB, int WC>
__device__ __forceinline__ void row_sums_fixed(const uint4* terms, const uint32_t* pa, uint32_t nrows, fr_t& a,
                                               fr_t& b, fr_t& c) {
    const uint32_t* pb = pa + WA * nrows;
    const uint32_t* pc = pb + WB * nrows;
    constexpr int WM = WA > WB ? (WA > WC ? WA : WC) : (WB > WC ? WB : WC);
    constexpr int L = Lazy<P>::kTerms;
#pragma unroll
    for (int j = 0; j < WM; ++j) {
        fr_t ta, tb, tc;
        if (j < WA) ta = signed_term<P>(terms, pa[j * nrows]);
        if (j < WB) tb = signed_term<P>(terms, pb[j * nrows]);
        if (j < WC) tc = load_term(terms, pc[j * nrows]);  // C entries carry no sign (upload tags -1 in C general)
        if (j == 0) {
            a = ta;
            b = tb;
            c = tc;
        } else {
            // the last L-1 steps of a may skip the reduction: value <= p * (1 + remaining steps) <= L * p
            if (j < WA) a = (WA - j <= L - 1) ? acc_step<P, true>(a, ta) : acc_step<P, false>(a, ta);
            if (j < WB) b = acc_step<P, false>(b, tb);
            if (j < WC) c = fr_add<P>(c, tc);  // both < p: canonical
        }
    }
}
// Run-time widths.  a, b, c in [0, p).
template <class P>
__device__ __forceinline__ void row_sums_any(const uint4* terms, const uint32_t* pa, uint32_t nrows, uint32_t wa,
                                             uint32_t wb, uint32_t wc, fr_t& a, fr_t& b, fr_t& c) {
    a = fr_zero<P>();
    b = fr_zero<P>();
    c = fr_zero<P>();
    const uint32_t* pb = pa + wa * nrows;
    const uint32_t* pc = pb + wb * nrows;
    const uint32_t wmax = max(wa, max(wb, wc));
    for (uint32_t j = 0; j < wmax; ++j) {  // trip count and the three predicates are warp-uniform
        const bool ua = j < wa, ub = j < wb, uc = j < wc;
        fr_t ta, tb, tc;
        if (ua) ta = signed_term<P>(terms, pa[j * nrows]);
        if (ub) tb = signed_term<P>(terms, pb[j * nrows]);
        if (uc) tc = load_term(terms, pc[j * nrows]);
        if (ua) a = fr_add<P>(a, ta);
        if (ub) b = fr_add<P>(b, tb);
        if (uc) c = fr_add<P>(c, tc);
    }
}
// canonical representative of a value < 2^256 (EMIT path only)
template <class P>
__device__ __forceinline__ fr_t canonical(fr_t x) {
#pragma unroll
    for (int i = 0; i < Lazy<P>::kTerms; ++i) x = fr_add<P>(x, fr_zero<P>());
    return x;
}
}  // namespace tiled

// phase cycle counters of the TIMING instantiation (a measurement aid, compiled only with -DACG_TILED_TIMING_BUILD and
// switched on with ACG_TILED_TIMING=1; the shipped library does not contain it)
__device__ unsigned long long g_tiled_phase_cycles[2][8];
// .. and per-CTA wall-clock marks (globaltimer, ns): kernel entry, first tile staged, last tile done, exit
constexpr uint32_t kMaxTimedCtas = 2048;
__device__ unsigned long long g_tiled_cta_marks[kMaxTimedCtas][6];  // [4] = SM id, [5] = tiles of the CTA
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}


// One CTA walks a contiguous run of tiles.  While tile i computes:
//   * its far witness elements are already in the far buffer (i & 1): they were gathered with 16-byte cp.async
//     copies issued during tile i - 1, from the far witness columns that blob i - 1 carried;
//   * the gathers of tile i + 1 are in flight into the other far buffer;
//   * the blob and window of tile i + 1 are in (or on their way to) L2 (cp.async.bulk.prefetch.L2 issued a tile
//     earlier), so that the bulk copies issued the moment P3 of tile i is done complete after one L2 round trip.
// The long rows of a system, checked by the tiled kernel's CTAs after their tiles (kernels.h DevLongRows): warp per row,
// rows claimed from the system's counter.  (Only in the LONG instantiations of the kernel: the one every Split-free
// system runs is untouched.  __noinline__ here crashes nvcc 12.9.)
template <class P, bool EMIT>
__device__ __forceinline__ void check_long_rows(const DevLongRows* lrp, uint32_t claim_base, const fr_t* w,
                                                uint64_t row_base, unsigned long long* gate, unsigned long long gate_seq,
                                                unsigned long long* out, unsigned long long* accum, uint32_t overlap,
                                                fr_t* __restrict__ Aw, fr_t* __restrict__ Bw, fr_t* __restrict__ Cw) {
    const DevLongRows& lr = *lrp;
    for (;;) {
        uint32_t i = 0;
        if (lane_id() == 0u) i = atomicAdd(lr.counter, 1u) - claim_base;  // (wraps consistently with the host's count)
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= lr.n_rows) break;
        const uint32_t row = lr.rows[i];
        const fr_t a = row_dot_warp<P, 1>(lr.m.m[0], row, w, lr.m.tagged != 0);
        const fr_t b = row_dot_warp<P, 1>(lr.m.m[1], row, w, lr.m.tagged != 0);
        const fr_t c = row_dot_warp<P, 1>(lr.m.m[2], row, w, lr.m.tagged != 0);
        if (lane_id() == 0u) {
            if (EMIT) {
                if (Aw) Aw[row] = a;
                if (Bw) Bw[row] = b;
                if (Cw) Cw[row] = c;
            }
            if (!fr_eq(fr_mul<P>(a, b), c)) report_violations(gate, gate_seq, out, accum, overlap, 1ull, row_base + row);
        }
    }
}

template <class P, bool EMIT, int V, bool TIMING = false, bool LONG = false>
__global__ void __launch_bounds__(kTileGeom[V].threads, tiled::Cfg<V>::kCtasPerSm)
    k_r1cs_tiled(DevTileStream ts, const fr_t* __restrict__ w, uint64_t row_base, CheckEpilogue ep,
                 fr_t* __restrict__ Aw, fr_t* __restrict__ Bw, fr_t* __restrict__ Cw,
                 const DevLongRows* __restrict__ long_rows, uint32_t long_claim_base) {
    using namespace tiled;
    using C = Cfg<V>;
    static_assert(kTileGeom[V].max_far <= kFarPerThread * C::kThreads, "far slots exceed the per-thread gathers");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar;
    __shared__ unsigned long long ph[2][8];  // TIMING only: per-phase cycles seen by warp 0 and by the last warp

    if (TIMING && threadIdx.x == 0 && blockIdx.x < kMaxTimedCtas) g_tiled_cta_marks[blockIdx.x][0] = globaltimer_ns();
    if (!ep.overlap) griddep_wait();   // (an overlapped launch reads only what the previous check also only read)
    griddep_launch_dependents();       // the next check may move in as CTAs of this one exit
    const uint32_t tid = threadIdx.x;
    if (ep.gate != nullptr && blockIdx.x == 0u && tid == C::kThreads - 1u) open_gate(ep);  // (not the thread that loads)
    const uint32_t lane = tid & 31u;
    const bool rec = TIMING && (tid == 0u || tid == C::kThreads - 32u);
    const uint32_t rw = tid == 0u ? 0u : 1u;
    long long t_last = 0;
    auto mark = [&](int k) {
        if (TIMING && rec) {
            const long long now = clock64();
            ph[rw][k] += (unsigned long long)(now - t_last);
            t_last = now;
        }
    };
    if (TIMING && tid < 16u) ph[tid >> 3][tid & 7u] = 0ull;
    uint8_t* blob = smem;
    uint4* smem4 = reinterpret_cast<uint4*>(smem);  // entry / operand words address 16-byte chunks from here
    // this CTA's run of tiles
    uint32_t t_begin, t_end;
    TileMeta tm;
    const bool planned = ts.runs != nullptr && ts.run_far != nullptr && gridDim.x == ts.n_runs;  // weighted runs (CtaRun)
    uint32_t first_far_col[kFarPerThread];  // this thread's far columns of the run's first tile (DevTileStream::run_far)
    if (planned) {
#pragma unroll
        for (uint32_t k = 0; k < kFarPerThread; ++k) {
            const uint32_t f = tid + k * C::kThreads;
            first_far_col[k] = f < C::kFarN ? __ldg(ts.run_far + (size_t)blockIdx.x * C::kFarN + f) : 0u;
        }
        const CtaRun run = ts.runs[blockIdx.x];
        t_begin = run.t_begin;
        t_end = run.t_end;
        tm = run.first;
    } else {
        t_begin = (uint32_t)((uint64_t)blockIdx.x * ts.n_tiles / gridDim.x);
        t_end = (uint32_t)((uint64_t)(blockIdx.x + 1u) * ts.n_tiles / gridDim.x);
    }
    if (t_begin >= t_end) {  // (the launcher never makes the grid larger than the tile count)
        if (LONG)
            check_long_rows<P, EMIT>(long_rows, long_claim_base, w, row_base, ep.gate, ep.gate_seq,
                                     ep.out, ep.accum, ep.overlap, Aw, Bw, Cw);
        finish_check(ep);
        return;
    }
    if (!planned) tm = ts.meta[t_begin];  // the first tile of the run is described from outside
    // column 0 is the constant wire: w[0] == 1 for every witness the reference builds, and then a general
    // coefficient on column 0 is its own product (kernels.h); anything else takes the multiply-in-place path
    // (loaded here, behind the tile record, and first looked at in P2: off the start-up critical path)
    const fr_t w0_first = ld_witness(w);
    // start the gather of this thread's far witness elements of tile `tile` (far slot f = tid + k * threads)
    // from the column list `cols`
    auto gather_far_async = [&](uint32_t tile, uint32_t n_far, const uint32_t* cols) {
        const uint32_t far0 = C::kFar0 + (C::kFarDouble ? (tile & 1u) * C::kFarN : 0u);
#pragma unroll
        for (uint32_t k = 0; k < kFarPerThread; ++k) {
            const uint32_t f = tid + k * C::kThreads;
            if (f < n_far) {
                const uint4* src = reinterpret_cast<const uint4*>(w + cols[f]);
                const uint32_t c = C::term_chunk(far0 + f);
                cp_async16(smem4 + c, src);
                cp_async16(smem4 + (c ^ 1u), src + 1);
            }
        }
    };

    uint32_t next_off16;  // where the next blob starts (thread 0)
    {
        if (tid == 0) {
            mbar_init(&full_bar, 1);
            mbar_fence_init();
            store_term(smem4, C::term_chunk(C::kZero), fr_zero<P>());
            issue_tile_load(ts, w, tm.blob_off16, tm.blob_bytes, tm.win_lo, tm.win_n, smem, smem + C::kOffTerms,
                            &full_bar, true);
            prefetch_behind(ts, w, tm.blob_off16, tm.blob_bytes, tm.win_lo, tm.win_n);
        }
        next_off16 = tm.blob_off16 + tm.blob_bytes / 16u;
        if (planned) {
            const uint32_t far0 = C::kFar0 + (C::kFarDouble ? (t_begin & 1u) * C::kFarN : 0u);
#pragma unroll
            for (uint32_t k = 0; k < kFarPerThread; ++k) {
                const uint32_t f = tid + k * C::kThreads;
                if (f < tm.n_far) {
                    const uint4* src = reinterpret_cast<const uint4*>(w + first_far_col[k]);
                    const uint32_t c = C::term_chunk(far0 + f);
                    cp_async16(smem4 + c, src);
                    cp_async16(smem4 + (c ^ 1u), src + 1);
                }
            }
        } else {
            gather_far_async(t_begin, tm.n_far, ts.far_cols + tm.far_off);
        }
    }
    const bool w0_is_one = fr_is_one<P>(w0_first);
    __syncthreads();  // mbarrier initialised before anyone waits on it

    uint32_t it = 0;
    if (TIMING) t_last = clock64();
    for (uint32_t tile = t_begin; tile < t_end; ++tile, ++it) {
        mbar_wait(&full_bar, it & 1u);
        cp_async_wait_all();  // this thread's far gathers of the tile
        if (TIMING && it == 0u && tid == 0u && blockIdx.x < kMaxTimedCtas)
            g_tiled_cta_marks[blockIdx.x][1] = globaltimer_ns();
        mark(0);
        __syncthreads();      // .. and everybody else's
        mark(1);
        // (a copy in registers on purpose: reading the fields from shared memory where they are used -- fewer live
        // registers in P3 -- measured 2 us SLOWER at 2^20 rows)
        const TileHeader h = *reinterpret_cast<const TileHeader*>(blob);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(blob + h.off_words);
        const uint16_t* gop = reinterpret_cast<const uint16_t*>(blob + h.off_gop);
        // the far witness elements of the NEXT tile start their way into the other far buffer
        if (C::kFarDouble && tile + 1u < t_end)
            gather_far_async(tile + 1u, h.next_n_far, reinterpret_cast<const uint32_t*>(blob + h.off_next_far));

        // ---- P2: dense 256-bit Montgomery products, one general entry per lane (no divergence between
        //          coefficient kinds): product slot <- coefficient * operand slot
        const uint32_t n_general = h.n_general, off_gval = h.off_gval;  // (the stores below could alias the header)
        for (uint32_t j = tid; j < n_general; j += C::kThreads)
            store_term(smem4, C::kProdInPlace ? C::blob_chunk(off_gval, j) : C::term_chunk(C::kProd0 + j),
                       fr_mul<P>(load_term(smem4, C::blob_chunk(off_gval, j)), load_term(smem4, gop[j])));
        if (!w0_is_one) {  // not a witness of the reference: coefficient * w[0] in place, inside the blob
            const fr_t w0 = ld_witness(w);
            const uint32_t n_all = n_general + h.n_const;
            for (uint32_t j = n_general + tid; j < n_all; j += C::kThreads) {
                const uint32_t c = C::blob_chunk(off_gval, j);
                store_term(smem4, c, fr_mul<P>(load_term(smem4, c), w0));
            }
        }
        mark(2);
        __syncthreads();
        mark(3);

        // ---- P3: thread per row, warp-uniform, shared memory only: the sums A.w, B.w, C.w advance together
        //          slot by slot (three independent carry chains); a -1 coefficient negates under a predicate
        // (rows are sorted by shape and the words laid out per warp, kernels.h TileWarp: this warp's widths, not the tile's)
        bool bad = false;
        const TileWarp tw = reinterpret_cast<const TileWarp*>(blob + kTileWarpsOffset)[tid >> 5];
        if (lane < tw.nrows) {
            const uint32_t wA = tw.width[0], wB = tw.width[1], wC = tw.width[2];
            const uint32_t* wp = words + tw.words0 + lane;
            fr_t a, b, c;
            if (wA == 2u && wB == 2u && wC == 1u) {  // the common shapes: straight-line code, no predicates
                row_sums_fixed<P, 2, 2, 1>(smem4, wp, tw.nrows, a, b, c);
            } else if (wA == 3u && wB == 2u && wC == 1u) {
                row_sums_fixed<P, 3, 2, 1>(smem4, wp, tw.nrows, a, b, c);
            } else if (wA == 2u && wB == 3u && wC == 1u) {
                row_sums_fixed<P, 2, 3, 1>(smem4, wp, tw.nrows, a, b, c);
            } else if (wA == 3u && wB == 3u && wC == 1u) {
                row_sums_fixed<P, 3, 3, 1>(smem4, wp, tw.nrows, a, b, c);
            } else {
                row_sums_any<P>(smem4, wp, tw.nrows, wA, wB, wC, a, b, c);
            }
            if (EMIT) {
                a = canonical<P>(a);
                b = fr_add<P>(b, fr_zero<P>());
                const uint32_t row = h.row0 + blob[kTilePermOffset + tid];
                if (Aw) Aw[row] = a;
                if (Bw) Bw[row] = b;
                if (Cw) Cw[row] = c;
            }
            bad = !fr_eq(fr_mul<P>(b, a), c);  // b (<= p) is the vector operand, a the limb-wise scalar
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, bad);
        if (bal != 0u) {  // rare: violated rows -- count them and find the smallest ORIGINAL row among them
            const uint32_t first = __reduce_min_sync(0xffffffffu, bad ? (uint32_t)blob[kTilePermOffset + tid] : 0xFFFFFFFFu);
            if (lane == 0u)
                report_violations(ep.gate, ep.gate_seq, ep.out, ep.accum, ep.overlap,
                                  (unsigned long long)__popc(bal), (unsigned long long)(row_base + h.row0 + first));
        }
        mark(4);
        if (tile + 1u == t_end) break;
        const uint32_t nx_bytes = h.next_bytes, nx_win_lo = h.next_win_lo, nx_win_n = h.next_win_n;
        const uint32_t nx_n_far = C::kFarDouble ? 0u : h.next_n_far;
        // one far buffer: this thread's far columns of the next tile leave the blob before it is refilled
        uint32_t next_far_col[kFarPerThread];
        if (!C::kFarDouble) {
#pragma unroll
            for (uint32_t k = 0; k < kFarPerThread; ++k) {
                const uint32_t f = tid + k * C::kThreads;
                next_far_col[k] = f < nx_n_far ? reinterpret_cast<const uint32_t*>(blob + h.off_next_far)[f] : 0u;
            }
        }
        // blob and window were read (and the term array written) through the generic proxy; order that before
        // the TMA (async proxy) refill of the same bytes
        fence_proxy_async_smem();
        __syncthreads();
        mark(5);
        if (!C::kFarDouble) {  // .. and are gathered now that every row of this tile has read the far buffer
#pragma unroll
            for (uint32_t k = 0; k < kFarPerThread; ++k) {
                const uint32_t f = tid + k * C::kThreads;
                if (f < nx_n_far) {
                    const uint4* src = reinterpret_cast<const uint4*>(w + next_far_col[k]);
                    const uint32_t c = C::term_chunk(C::kFar0 + f);
                    cp_async16(smem4 + c, src);
                    cp_async16(smem4 + (c ^ 1u), src + 1);
                }
            }
        }
        if (tid == 0)
            issue_tile_load(ts, w, next_off16, nx_bytes, nx_win_lo, nx_win_n, smem, smem + C::kOffTerms, &full_bar);
        else if (tid == C::kThreads - 1u)  // another warp: the prefetches do not delay the loads
            prefetch_behind(ts, w, next_off16, nx_bytes, nx_win_lo, nx_win_n);
        next_off16 += nx_bytes / 16u;
        mark(6);
    }
    if (TIMING && rec) {
        for (int k = 0; k < 7; ++k) atomicAdd(&g_tiled_phase_cycles[rw][k], ph[rw][k]);
        atomicAdd(&g_tiled_phase_cycles[rw][7], (unsigned long long)it);
    }
    if (TIMING && tid == 0u && blockIdx.x < kMaxTimedCtas) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_tiled_cta_marks[blockIdx.x][2] = globaltimer_ns();
        g_tiled_cta_marks[blockIdx.x][4] = smid;
        g_tiled_cta_marks[blockIdx.x][5] = t_end - t_begin;
    }
    if (LONG)
        check_long_rows<P, EMIT>(long_rows, long_claim_base, w, row_base, ep.gate, ep.gate_seq,
                                 ep.out, ep.accum, ep.overlap, Aw, Bw, Cw);
    finish_check(ep);
    if (TIMING && tid == 0u && blockIdx.x < kMaxTimedCtas) g_tiled_cta_marks[blockIdx.x][3] = globaltimer_ns();
}

template <class P>
__global__ void k_to_mont_scattered(uint8_t* __restrict__ blobs, const uint32_t* __restrict__ offs, uint64_t n,
                                    int* __restrict__ bad_flag) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        // offs[i]: chunk (16-byte unit of the stream) of the low half; bit 31: the high half is the chunk before it
        // instead of the one after it (a swizzled value, kernels.h swz16)
        const uint32_t o = offs[i];
        uint4* lo = reinterpret_cast<uint4*>(blobs + (size_t)(o & 0x7FFFFFFFu) * 16u);
        uint4* hi = (o >> 31) ? lo - 1 : lo + 1;
        const uint4 a = *lo, b = *hi;
        fr_t x;
        x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
        x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
        if (!fr_is_canonical<P>(x)) {
            *bad_flag = 1;
        } else {
            const fr_t y = fr_to_mont<P>(x);
            *lo = make_uint4(y.l[0], y.l[1], y.l[2], y.l[3]);
            *hi = make_uint4(y.l[4], y.l[5], y.l[6], y.l[7]);
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
cudaError_t launch_copy_if_clean(fr_t* dst, const fr_t* src, uint64_t n, const int* d_bad_flag, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k_copy_if_clean<<<grid_for(2 * n, 256, 148 * 16), 256, 0, s>>>(reinterpret_cast<uint4*>(dst),
                                                                    reinterpret_cast<const uint4*>(src), 2 * n, d_bad_flag);
    return cudaGetLastError();
}
cudaError_t launch_init_result(unsigned long long* d_result, cudaStream_t s) {
    k_init_result<<<1, 1, 0, s>>>(d_result);
    return cudaGetLastError();
}

cudaError_t launch_r1cs_rowwise(int field, const DevR1cs& m, const fr_t* w, uint32_t row_lo, uint32_t row_hi,
                                uint64_t row_base, const CheckEpilogue& ep, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                                cudaStream_t s) {
    if (row_hi <= row_lo) return cudaSuccess;
    const bool emit = Aw || Bw || Cw;
    const unsigned grid = grid_for(row_hi - row_lo, 256, 148 * 64);
    if (emit) {
        ACG_DISPATCH_FIELD(field, (k_r1cs_rowwise<P, true><<<grid, 256, 0, s>>>(m, w, row_lo, row_hi, row_base, ep,
                                                                                 Aw, Bw, Cw)));
    } else {
        ACG_DISPATCH_FIELD(field, (k_r1cs_rowwise<P, false><<<grid, 256, 0, s>>>(m, w, row_lo, row_hi, row_base, ep,
                                                                                  Aw, Bw, Cw)));
    }
    return cudaGetLastError();
}

cudaError_t launch_r1cs_longrows(int field, const DevR1cs& m, const fr_t* w, const uint32_t* d_rows, uint32_t n_rows,
                                 uint64_t row_base, const CheckEpilogue& ep, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                                 cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    const bool emit = Aw || Bw || Cw;
    const unsigned grid = grid_for((uint64_t)n_rows * 32u, 128, 148 * 16);
    if (emit) {
        ACG_DISPATCH_FIELD(field, (k_r1cs_longrows<P, true><<<grid, 128, 0, s>>>(m, w, d_rows, n_rows, row_base, ep, Aw,
                                                                                  Bw, Cw)));
    } else {
        ACG_DISPATCH_FIELD(field, (k_r1cs_longrows<P, false><<<grid, 128, 0, s>>>(m, w, d_rows, n_rows, row_base, ep, Aw,
                                                                                   Bw, Cw)));
    }
    return cudaGetLastError();
}

// Placement probe (kernels.h CtaRun): the tiled kernel's launch geometry on an empty body
template <int V>
__global__ void __launch_bounds__(kTileGeom[V].threads, tiled::Cfg<V>::kCtasPerSm)
    k_probe_placement(uint32_t* __restrict__ smid_out, unsigned int* __restrict__ arrived) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        smid_out[blockIdx.x] = smid;
        smem[0] = (uint8_t)smid;  // (keeps the dynamic shared memory allocation alive)
        __threadfence();
        atomicAdd(arrived, 1u);
        // stay resident until every block has been placed (they all fit at once: grid = SMs x resident CTAs per SM);
        // give up after ~50 ms rather than hang if something else occupies the device
        const long long t0 = clock64();
        while (atomicAdd(arrived, 0u) < gridDim.x && clock64() - t0 < 100000000ll) __nanosleep(200);
    }
    __syncthreads();
}
template <int V>
static cudaError_t launch_probe_impl(int sm_count, uint32_t* d_smid, unsigned int* d_arrived, uint32_t* grid_out,
                                     cudaStream_t s) {
    using C = tiled::Cfg<V>;
    cudaError_t e = cudaFuncSetAttribute(k_probe_placement<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kBytes);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)sm_count * C::kCtasPerSm;
    *grid_out = grid;
    k_probe_placement<V><<<grid, kTileGeom[V].threads, C::kBytes, s>>>(d_smid, d_arrived);
    return cudaGetLastError();
}
cudaError_t launch_probe_placement(int variant, int sm_count, uint32_t* d_smid, unsigned int* d_arrived,
                                   uint32_t* grid_out, cudaStream_t s) {
    switch (variant) {
        case 1: return launch_probe_impl<1>(sm_count, d_smid, d_arrived, grid_out, s);
        case 2: return launch_probe_impl<2>(sm_count, d_smid, d_arrived, grid_out, s);
        case 3: return launch_probe_impl<3>(sm_count, d_smid, d_arrived, grid_out, s);
        case 4: return launch_probe_impl<4>(sm_count, d_smid, d_arrived, grid_out, s);
        case 5: return launch_probe_impl<5>(sm_count, d_smid, d_arrived, grid_out, s);
        case 6: return launch_probe_impl<6>(sm_count, d_smid, d_arrived, grid_out, s);
        case 7: return launch_probe_impl<7>(sm_count, d_smid, d_arrived, grid_out, s);
        case 8: return launch_probe_impl<8>(sm_count, d_smid, d_arrived, grid_out, s);
        default: return launch_probe_impl<0>(sm_count, d_smid, d_arrived, grid_out, s);
    }
}
uint32_t tiled_ctas_per_sm(int variant) {
    return variant >= 0 && variant < kNumTileVariants ? tile_ctas_per_sm(kTileGeom[variant]) : 0u;
}

// geometries whose kernel also exists in the form that checks the long rows in the same launch (the default and the
// dense one; the others take the separate warp-per-row launch first)
bool tiled_checks_long_rows(int variant) {
    return variant == 0 || variant == kDenseTileVariant;
}

template <class P, bool EMIT, int V, bool LONG = false>
static cudaError_t launch_tiled_impl(const DevTileStream& ts, const fr_t* w, uint64_t row_base,
                                     const CheckEpilogue& ep, fr_t* Aw, fr_t* Bw, fr_t* Cw, int sm_count,
                                     cudaStream_t s, const DevLongRows* long_rows = nullptr, uint32_t n_long_rows = 0,
                                     uint32_t* long_claims = nullptr) {
    using C = tiled::Cfg<V>;
    // ACG_K2_CTAS_PER_SM=k (a measurement aid): pad the dynamic shared memory so that only k CTAs fit on an SM
    static const int limit_ctas = getenv("ACG_K2_CTAS_PER_SM") ? atoi(getenv("ACG_K2_CTAS_PER_SM")) : 0;
    unsigned smem_bytes = C::kBytes, ctas = C::kCtasPerSm;
    if (limit_ctas > 0 && (unsigned)limit_ctas < ctas) {
        ctas = (unsigned)limit_ctas;
        smem_bytes = (227u * 1024u) / ctas - 1024u - 256u;
    }
    {
        cudaError_t e = cudaFuncSetAttribute(k_r1cs_tiled<P, EMIT, V, false, LONG>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    unsigned grid = (unsigned)sm_count * ctas;
    if (grid > ts.n_tiles) grid = ts.n_tiles;
    const DevTileStream& tsw = ts;
#ifdef ACG_TILED_TIMING_BUILD  // measurement build only (ACG_NVCC_EXTRA=-DACG_TILED_TIMING_BUILD python build.py --force)
    if (V == 0 && !EMIT && !LONG) {  // ACG_TILED_TIMING=1: run the instrumented instantiation and print its counters
        static const bool timing = getenv("ACG_TILED_TIMING") != nullptr;
        if (timing) {
            cudaFuncSetAttribute(k_r1cs_tiled<P, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)C::kBytes);
            unsigned long long z[2][8] = {};
            cudaMemcpyToSymbolAsync(g_tiled_phase_cycles, z, sizeof z, 0, cudaMemcpyHostToDevice, s);
            k_r1cs_tiled<P, false, 0, true><<<grid, kTileGeom[0].threads, C::kBytes, s>>>(tsw, w, row_base, ep, Aw, Bw,
                                                                                          Cw, nullptr, 0u);
            cudaMemcpyFromSymbolAsync(z, g_tiled_phase_cycles, sizeof z, 0, cudaMemcpyDeviceToHost, s);
            static unsigned long long marks[kMaxTimedCtas][6];
            cudaMemcpyFromSymbolAsync(marks, g_tiled_cta_marks, sizeof marks, 0, cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            static int printed = 0;
            if (printed < 3) {  // CTA timeline: percentiles over the CTAs, ns after the first CTA entered the kernel
                const unsigned n = grid < kMaxTimedCtas ? grid : kMaxTimedCtas;
                unsigned long long t0 = ~0ull;
                for (unsigned i = 0; i < n; ++i) t0 = marks[i][0] < t0 ? marks[i][0] : t0;
                if (const char* path = getenv("ACG_TILED_TIMING_DUMP")) {  // raw marks, one line per CTA
                    if (FILE* f = fopen(path, printed ? "a" : "w")) {
                        fprintf(f, "# launch %d: cta smid tiles entry first_staged last_done exit (ns)\n", printed);
                        for (unsigned i = 0; i < n; ++i)
                            fprintf(f, "%u %llu %llu %llu %llu %llu %llu\n", i, marks[i][4], marks[i][5], marks[i][0] - t0,
                                    marks[i][1] - t0, marks[i][2] - t0, marks[i][3] - t0);
                        fclose(f);
                    }
                }
                static const char* mn[4] = {"entry", "first tile staged", "last tile done", "exit"};
                for (int k = 0; k < 4; ++k) {
                    std::vector<unsigned long long> v(n);
                    for (unsigned i = 0; i < n; ++i) v[i] = marks[i][k] - t0;
                    std::sort(v.begin(), v.end());
                    fprintf(stderr, "[cta timeline ns, %s] min=%llu p10=%llu p50=%llu p90=%llu max=%llu\n", mn[k], v[0],
                            v[n / 10], v[n / 2], v[(size_t)n * 9 / 10], v[n - 1]);
                }
            }
            if (printed++ < 3) {
                static const char* nm[7] = {"tma_wait", "bar1", "p2", "bar2", "p3", "bar3", "refill"};
                for (int wsel = 0; wsel < 2; ++wsel) {
                    fprintf(stderr, "[phase cycles/tile, %s] tiles=%llu:", wsel ? "last warp" : "warp 0", z[wsel][7]);
                    unsigned long long tot = 0;
                    for (int k = 0; k < 7; ++k) {
                        fprintf(stderr, " %s=%.0f", nm[k], (double)z[wsel][k] / (double)z[wsel][7]);
                        tot += z[wsel][k];
                    }
                    fprintf(stderr, " total=%.0f\n", (double)tot / (double)z[wsel][7]);
                }
            }
            return cudaGetLastError();
        }
    }
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTileGeom[V].threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ep.overlap ? 1u : 0u;
    const uint32_t claim_base = LONG && long_claims ? *long_claims : 0u;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_r1cs_tiled<P, EMIT, V, false, LONG>, tsw, w, row_base, ep, Aw, Bw, Cw,
                                             long_rows, claim_base);
    // this check's claims: one per long row and one (past the end) per warp of the grid -- counted only for a launch
    // that was accepted, so that host and device counts cannot drift apart
    if (e == cudaSuccess && LONG && long_claims) *long_claims += n_long_rows + grid * (kTileGeom[V].threads / 32u);
    return e;
}

cudaError_t launch_r1cs_tiled(int field, const DevTileStream& ts, const fr_t* w, uint64_t row_base,
                              const CheckEpilogue& ep, fr_t* Aw, fr_t* Bw, fr_t* Cw, int sm_count, cudaStream_t s,
                              const DevLongRows* long_rows, uint32_t n_long_rows, uint32_t* long_claims) {
    if (ts.n_tiles == 0) return cudaSuccess;
    const bool emit = Aw || Bw || Cw;
    if (long_rows != nullptr) {  // the tiles and, in the same launch, the long rows
        if (!tiled_checks_long_rows((int)ts.variant) || long_claims == nullptr) return cudaErrorInvalidValue;
#define ACG_TILED_LONG(EMITV, VAR)                                                                                   \
    ACG_DISPATCH_FIELD(field, return (launch_tiled_impl<P, EMITV, VAR, true>(ts, w, row_base, ep, Aw, Bw, Cw, sm_count, s, \
                                                                             long_rows, n_long_rows, long_claims)))
        if (ts.variant == 0u) {
            if (emit) ACG_TILED_LONG(true, 0);
            ACG_TILED_LONG(false, 0);
        } else {
            if (emit) ACG_TILED_LONG(true, kDenseTileVariant);
            ACG_TILED_LONG(false, kDenseTileVariant);
        }
#undef ACG_TILED_LONG
    }
#define ACG_TILED(EMITV, VAR)                                                                                     \
    ACG_DISPATCH_FIELD(field, return (launch_tiled_impl<P, EMITV, VAR>(ts, w, row_base, ep, Aw, Bw, Cw, sm_count, s)))
#define ACG_TILED_V(VAR)                \
    do {                                \
        if (emit) ACG_TILED(true, VAR); \
        ACG_TILED(false, VAR);          \
    } while (0)
    switch (ts.variant) {
        case 1: ACG_TILED_V(1); break;
        case 2: ACG_TILED_V(2); break;
        case 3: ACG_TILED_V(3); break;
        case 4: ACG_TILED_V(4); break;
        case 5: ACG_TILED_V(5); break;
        case 6: ACG_TILED_V(6); break;
        case 7: ACG_TILED_V(7); break;
        case 8: ACG_TILED_V(8); break;
        default: ACG_TILED_V(0); break;
    }
#undef ACG_TILED_V
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
