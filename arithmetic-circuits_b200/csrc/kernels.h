// Host-visible launchers of the CUDA kernels (K1..K5 of SURVEY.md section 2).  Each launcher
// dispatches on field_id to the template instantiation and only ENQUEUES on `stream`.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr.cuh"

namespace acg {

// Column words carry a 2-bit coefficient tag in bits 31:30 when the system was preprocessed at upload
// (DevR1cs::tagged): the sparsity pattern and the coefficients are static, so classifying them once is
// free for every later check.
constexpr uint32_t kColMask = 0x3FFFFFFFu;   // witness columns are < 2^30
constexpr uint32_t kTagPlusOne = 0u;   // coefficient == 1
constexpr uint32_t kTagMinusOne = 1u;  // coefficient == r - 1
constexpr uint32_t kTagGeneral = 2u;

struct DevCsr {
    const uint32_t* rowptr;  // shard-local rows + 1 entries, rebased to 0, padded by >= 8
    const uint32_t* col;     // padded by >= 8; tagged when DevR1cs::tagged
    const fr_t* val;         // Montgomery form, one per entry
};
struct DevR1cs {
    DevCsr m[3];             // A, B, C
    uint32_t tagged;
};

// ---- tile stream (tiled kernel) -------------------------------------------------------------------
// The sparsity pattern and the coefficients are static, so upload lays the system out as a stream of
// self-contained, execution-ready tile blobs; the kernel stages one blob per tile with a single TMA
// bulk copy, plus one more bulk copy of the tile's WITNESS WINDOW: the contiguous slice of w that most
// of the tile's references fall into (for circuits built gate by gate: the wires defined just before /
// by the tile's own rows).
//
// Inside the kernel every term of every row lives in ONE shared-memory array of 32-byte slots:
//     [0, window)                         the witness window (TMA)
//     [window, window + 2 * max_far)      "far" witness elements: the distinct columns outside the window,
//                                         gathered once per tile; two buffers, used by even / odd tiles
//     [.., + max_gen)                     products coefficient * operand of the general-coefficient entries
//     last slot                           zero (padding)
// Blob = header | entry words in ELL (slot-major) order: for matrix A, then B, then C, for slot
// j < width[k], for row r < nrows: one 32-bit word = sign<<31 | chunk (16-byte unit from the start of shared
// memory) of the term's low half, see swz16 below (coefficient -1 sets the sign; general entries point at their
// product slot, padding at the zero slot) | u32 far witness columns of the next tile | u16 operand chunks of the
// general entries | their
// coefficient values (Montgomery; 32-byte aligned in shared memory), followed by the coefficient values of
// the general entries ON COLUMN 0.  Column 0 is the constant wire of the reference (w[0] = 1 by construction:
// initialQapSet, src/QAP.hs:591-595; qapSetToMap puts it at index 0, :605-620), so coefficient * w[0] is the
// coefficient itself: the entry word of such an entry points at the coefficient inside the blob and no product
// is computed.  The kernel checks w[0] == 1 at run time and, if a caller passes something else, multiplies
// these coefficients by w[0] in place first -- the result is the same for every witness.
// Every section is 16-byte aligned.
// A CTA walks a CONTIGUOUS run of tiles, so the stream is linear per CTA.  What the kernel must know before
// a blob is in shared memory travels one blob earlier: blob i carries the far witness columns of tile i + 1
// and, in its header, that tile's blob size and window.  With them the CTA issues the bulk copies of tile
// i + 1 and gathers its far witness elements the moment tile i is done, without a dependent global read.
// Only the first tile of a run is described from outside (TileMeta, DevTileStream::far_cols).  Row r of the tile is handled by thread r, so slot-major order makes
// every word read of a warp contiguous and the control flow of the row sums warp-uniform.
struct alignas(16) TileHeader {
    uint32_t row0;          // first (shard-local) row
    uint32_t nrows;
    uint32_t n_general;
    uint32_t bytes;         // blob size (multiple of 16)
    uint32_t width[3];      // ELL widths (max row length in the tile) of A, B, C
    uint32_t off_words;     // byte offsets inside the blob
    uint32_t off_gop;
    uint32_t off_gval;
    uint32_t off_next_far;  // u32 far witness columns of the NEXT tile of the stream
    uint32_t next_n_far;    // .. and what else is needed to start loading it: its far count,
    uint32_t next_bytes;    //    blob size (it starts where this blob ends)
    uint32_t next_win_lo;   //    and witness window [next_win_lo, next_win_lo + next_win_n)
    uint32_t next_win_n;
    uint32_t n_const;       // general entries on column 0 (values follow the n_general product coefficients)
};
static_assert(sizeof(TileHeader) == 64, "TileHeader must be 64 bytes");
// Rows are SORTED BY SHAPE inside a tile (by the lengths of their A, B and C rows), thread t handles the t-th row of
// that order, and the entry words are laid out per WARP: warp q owns the words of its <= 32 rows, slot-major, with
// the ELL widths of ITS rows -- in the synthetic family 56 % of the rows have two-entry A and B rows, and their warps
// no longer execute the third (padding) slot that a tile-wide width would force on them.  The header is followed by
// kMaxTileWarps of these records and by one byte per thread: the row's offset from row0 (for emitted vectors and
// violation reports, which are by original row).
struct TileWarp {
    uint16_t words0;   // first word of the warp's block, in words from off_words
    uint8_t nrows;     // rows of this warp (<= 32)
    uint8_t width[3];  // ELL widths of A, B, C over the warp's rows
    uint8_t pad[2];
};
static_assert(sizeof(TileWarp) == 8, "TileWarp must be 8 bytes");
constexpr uint32_t kMaxTileWarps = 8;
constexpr uint32_t kTileWarpsOffset = 64;                                  // byte offset of the TileWarp records
constexpr uint32_t kTilePermOffset = kTileWarpsOffset + kMaxTileWarps * 8; // .. and of the row-offset bytes
constexpr uint32_t kTermSign = 0x80000000u;

// Tiled kernel geometry (see DESIGN.md "K2"); the variant is bound when the system is uploaded.
struct TileGeometry {
    uint32_t threads;    // == max rows per tile
    uint32_t max_slots;  // sum of the three ELL widths (rows longer than kMaxEllWidth go to the row-wise kernel)
    uint32_t max_gen;    // general-coefficient entries per tile that need a product (column != 0)
    uint32_t window;     // witness elements staged per tile
    uint32_t max_far;    // distinct witness columns outside the window per tile
    uint32_t max_const;  // general-coefficient entries per tile on column 0 (the constant wire)
    uint32_t swizzle;    // 1: 16-byte halves of every term outside the witness window are XOR-swizzled (swz16)
    uint32_t far_bufs;   // 2: far slots double-buffered by tile parity (gathers of tile i + 1 run under tile i);
                         // 1: one far buffer, gathered between two tiles (smaller footprint, one more resident CTA)
    uint32_t prod_in_place;  // 1: a product overwrites its coefficient inside the blob (no separate product slots)
    uint32_t reg_budget; // registers per thread the kernel is compiled for (bounds the resident CTAs per SM)
};
constexpr uint32_t kMaxEllWidth = 8;
constexpr int kNumTileVariants = 9;
// 0: the default (128-row tiles); 1-3: 256 / 64 / 32 rows; 4, 5: 128 / 64 rows with the natural term layout (no swizzle);
// 6: products written over their coefficients; 7: 6 + one far buffer (6 CTAs per SM); 8: the geometry for systems dense
// in general coefficients -- room for 4 products per row (512 per tile, in place), chosen automatically at upload when
// the default would have to cut its tiles down to a third of their rows to stay within 160 products.
constexpr int kDenseTileVariant = 8;
constexpr TileGeometry kTileGeom[kNumTileVariants] = {
    {128, 10, 160, 192, 288, 96, 1, 2, 0, 96}, {256, 10, 320, 320, 576, 192, 1, 2, 0, 96},
    {64, 10, 96, 128, 160, 64, 1, 2, 0, 96},   {32, 10, 64, 96, 96, 32, 1, 2, 0, 96},
    {128, 10, 160, 192, 288, 96, 0, 2, 0, 96}, {64, 10, 96, 128, 160, 64, 0, 2, 0, 96},
    {128, 10, 160, 192, 288, 96, 1, 2, 1, 96}, {128, 10, 160, 192, 288, 96, 1, 1, 1, 85},
    {128, 10, 512, 192, 288, 160, 1, 2, 1, 96}};
// Shared memory is addressed in 16-byte CHUNKS from the start of the CTA's dynamic shared memory.  A term (32 bytes)
// occupies the two chunks of one 32-byte unit; an entry word names the chunk that holds its LOW half and the high
// half is the other chunk of the unit (word ^ 1).  Eight lanes of a 128-bit shared-memory access are served per
// wavefront from 8 bank groups of 16 bytes; with the low half always in the even chunk, the low-half loads of a warp
// can only ever use 4 of them (two-way conflicts even for consecutive terms).  swz16 swaps the halves in every other
// 128-byte row, which spreads both loads over all 8 groups.  The witness window is written by a linear TMA bulk copy
// and stays in natural order; everything the kernel or the upload writes (far terms, products, coefficient values)
// is swizzled.
ACG_HD constexpr uint32_t swz16(uint32_t chunk) {
    return chunk ^ ((chunk >> 3) & 1u);
}
constexpr uint32_t tile_blob_capacity(const TileGeometry& g) {
    return 64u + kMaxTileWarps * 8u + (g.threads + 15u) / 16u * 16u + g.threads * g.max_slots * 4u + g.max_far * 4u +
           ((g.max_gen * 2u + 15u) / 16u) * 16u + 16u + (g.max_gen + g.max_const) * 32u;
}
// shared-memory offset of the term array (the blob sits at offset 0); entry words and operand words address
// 16-byte chunks from the start of shared memory, so that a word can also point INTO the blob (see below)
constexpr uint32_t tile_terms_offset(const TileGeometry& g) {
    return (tile_blob_capacity(g) + 127u) / 128u * 128u;
}
// far gathers per thread (the far witness columns of a tile are gathered tid, tid + threads, ...)
constexpr uint32_t kFarPerThread = 3;
// the far slots are double-buffered by tile parity: tile i uses far buffer (i & 1), so the far witness
// elements of tile i + 1 are gathered (cp.async) while tile i computes
constexpr uint32_t tile_term_slots(const TileGeometry& g) {
    return g.window + g.far_bufs * g.max_far + (g.prod_in_place ? 0u : g.max_gen) + 1u;
}
constexpr uint32_t tile_far_slot0(const TileGeometry& g, uint32_t tile) {
    return g.window + (g.far_bufs == 2u ? (tile & 1u) * g.max_far : 0u);
}
constexpr uint32_t tile_prod_slot0(const TileGeometry& g) {
    return g.window + g.far_bufs * g.max_far;
}

struct alignas(32) TileMeta {
    uint32_t blob_off16;  // blob offset in 16-byte units
    uint32_t blob_bytes;  // multiple of 16
    uint32_t win_lo;      // witness window [win_lo, win_lo + win_n)
    uint32_t win_n;
    uint32_t far_off;     // first entry of the tile in DevTileStream::far_cols
    uint32_t n_far;
    uint32_t pad[2];
};
// resident CTAs of the tiled kernel per SM: bounded by shared memory (227 KB, 1 KB per CTA reserved), by the
// register budget per thread (TileGeometry::reg_budget) and by the hardware limit of 32
constexpr uint32_t tile_smem_bytes(const TileGeometry& g) {
    return tile_terms_offset(g) + tile_term_slots(g) * 32u;
}
constexpr uint32_t tile_ctas_per_sm(const TileGeometry& g) {
    const uint32_t by_smem = (227u * 1024u) / (tile_smem_bytes(g) + 1024u + 64u);
    const uint32_t by_regs = 65536u / (g.threads * g.reg_budget);
    const uint32_t m = by_smem < by_regs ? by_smem : by_regs;
    return m > 32u ? 32u : m;
}
struct CtaRun;
struct DevTileStream {
    const uint8_t* blobs;       // concatenated tile blobs
    const TileMeta* meta;       // n_tiles records
    const uint32_t* far_cols;   // witness columns of the far slots, tile after tile
    uint32_t n_tiles;
    uint32_t variant;
    uint32_t blobs_len16;       // length of `blobs` in 16-byte units
    uint32_t n_cols;            // witness length
    // Static split of the tiles over the CTAs when the grid fills the chip (n_runs == gridDim.x): CTA b walks the tiles
    // [runs[b].t_begin, runs[b].t_end).  The runs are UNEQUAL (see CtaRun); runs == nullptr: equal runs.
    const CtaRun* runs;
    uint32_t n_runs;
    // .. with them, the far witness columns of every run's FIRST tile at a fixed place (max_far words per run, CTA b's at
    // run_far + b * max_far): a CTA reads them alongside its run record instead of after it -- one round trip to memory
    // less before the first tile can be staged
    const uint32_t* run_far;
};
// The warp scheduler does not share an SM evenly between its resident CTAs: the first-launched CTA of an SM retires
// tiles ~30 % faster than the fifth (profiles/r02_cta_timeline_*), and the block scheduler's placement of blocks on SMs
// is a fixed but irregular pattern (blocks 0..5 go to SMs 142..147, some SMs receive their k-th block a hundred blocks
// earlier than others).  Equal runs therefore leave the SMs half empty for the last sixth of the kernel.  At upload the
// library PROBES the placement (k_probe_placement: same launch geometry, every block records its SM), ranks the blocks
// of each SM by arrival, and cuts the tiles into per-rank regions whose sizes follow the measured tile rates, each
// region split evenly over the SMs with the remainders placed cyclically -- every SM ends up with the same number of
// tiles (+-1) and every CTA with a run proportional to the rate it will get.  The record also carries the TileMeta of
// the run's first tile, so the prologue needs one dependent load, not two.
struct alignas(64) CtaRun {
    uint32_t t_begin, t_end;
    uint32_t pad[6];
    TileMeta first;
};
static_assert(sizeof(CtaRun) == 64, "CtaRun must be 64 bytes");

// All-reduce of the check result over peer memory (multi-GPU row shards): base[r] = rank r's exchange buffer as
// mapped into this process, 2 (sequence parity) x kMaxPeers slots of 4 x u64 {count, first bad row, sequence, pad}.
constexpr uint32_t kMaxPeers = 8;
constexpr size_t kPeerBufferBytes = 2 * kMaxPeers * 4 * sizeof(unsigned long long);
// (The table of bases lives in device memory, not in the struct: a kernel parameter indexed by lane would make the
// compiler copy the whole parameter block of the check kernel to local memory at the start of every thread.)
struct PeerSlots {
    unsigned long long* const* base;  // device array of kMaxPeers pointers
    unsigned long long* own;          // == base[rank]
    uint32_t world, rank;
};
// How a check kernel hands over its result.  The kernels accumulate {violation count, first bad row} into the
// context's scratch pair `accum`; the launch that carries `out != nullptr` (the last one of a check) finalises: its
// last CTA to finish (ticket counter) reads the pair, resets scratch and ticket for the next check, all-reduces the
// pair over peer memory when peers.world > 1 (see PeerSlots), and writes the final pair to `out`.  No separate
// initialisation or reduction launch, no collective library call.
// overlap != 0: the launch is a programmatic dependent of the previous check of the SAME system and witness on the
// same stream (cudaLaunchAttributeProgrammaticStreamSerialization): its CTAs take the place of the previous check's
// CTAs as those run out of tiles, and it waits for that check to complete (griddepcontrol.wait) only before it
// touches the shared scratch pair -- consecutive checks overlap their tails, launch latency and cold starts.
// gate != nullptr (tiled kernel, one GPU, no overlap): the DIRECT hand-over, which costs a clean check nothing at its
// end.  Block 0 starts by moving the scratch pair (what earlier launches of the same check accumulated; {0, none}
// otherwise) to `out` and then opens the gate -- a release store of gate_seq, the context's running check number, into
// one word of a ring --; a warp that finds violated rows (rare) waits for that word and updates `out` itself with
// atomics.  Nobody finalises: when the grid has drained, `out` is the result.  The ticket path costs the LAST CTA a
// fence, an atomic round trip, another fence and a load after its last tile -- about 3 us of a 63 us check.
struct CheckEpilogue {
    unsigned long long* accum;
    unsigned int* ticket;
    unsigned long long* out;
    PeerSlots peers;
    unsigned long long seq;
    uint32_t overlap;
    unsigned long long* gate;
    unsigned long long gate_seq;
};
constexpr uint32_t kGateRing = 256;  // checks of one context that may be in flight on different streams at once
// Publishes d_result[0..1] to every peer, waits for every peer's pair of step `seq`, leaves {sum of the counts,
// min of the first bad rows} in d_result (count = ~0 if a peer did not arrive within ~4 s).
cudaError_t launch_peer_allreduce(const PeerSlots& ps, unsigned long long seq, unsigned long long* d_result,
                                  cudaStream_t s);

cudaError_t launch_to_mont(int field, fr_t* v, uint64_t n, int* d_bad_flag, cudaStream_t s);
cudaError_t launch_from_mont(int field, fr_t* v, uint64_t n, cudaStream_t s);
// dst <- src (n elements) unless *d_bad_flag != 0
cudaError_t launch_copy_if_clean(fr_t* dst, const fr_t* src, uint64_t n, const int* d_bad_flag, cudaStream_t s);
cudaError_t launch_fr_binop(int field, int op, const fr_t* a, const fr_t* b, fr_t* o, uint64_t n, cudaStream_t s);
cudaError_t launch_init_result(unsigned long long* d_result, cudaStream_t s);

// thread-per-row check over local rows [row_lo, row_hi); Aw/Bw/Cw may be null
cudaError_t launch_r1cs_rowwise(int field, const DevR1cs& m, const fr_t* w, uint32_t row_lo, uint32_t row_hi,
                                uint64_t row_base, const CheckEpilogue& ep, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                                cudaStream_t s);
// warp-per-row check over a list of (shard-local) rows: every row too long for a tile, in ONE launch
cudaError_t launch_r1cs_longrows(int field, const DevR1cs& m, const fr_t* w, const uint32_t* d_rows, uint32_t n_rows,
                                 uint64_t row_base, const CheckEpilogue& ep, fr_t* Aw, fr_t* Bw, fr_t* Cw,
                                 cudaStream_t s);
// Placement probe (see CtaRun): launches the tiled kernel's grid for geometry `variant` on an empty body; block b writes
// the id of the SM it landed on to d_smid[b] and leaves only when all blocks have arrived (d_arrived: zeroed counter).
// *grid_out = blocks launched (SM count x resident CTAs per SM).
cudaError_t launch_probe_placement(int variant, int sm_count, uint32_t* d_smid, unsigned int* d_arrived,
                                   uint32_t* grid_out, cudaStream_t s);
// resident CTAs per SM of the tiled kernel for geometry `variant`
uint32_t tiled_ctas_per_sm(int variant);
// Rows too long for a tile (kMaxEllWidth), checked by the CTAs of the tiled kernel AFTER their runs of tiles, a warp
// per row on the plain CSR arrays: the same launch, all resident warps of the chip at once.  The warps CLAIM rows from
// a counter as they run out of tiles, so the CTAs that finish their tiles early take more of them.  The counter is
// never reset: every warp of the grid claims until it draws an index past the end, so one check advances it by exactly
// n_rows + (warps of the grid), and the launcher passes each check the value its claims start from (checks of one system
// are ordered on the device, include/acg.h).
struct DevLongRows {
    DevR1cs m;
    const uint32_t* rows;   // shard-local row indices
    uint32_t n_rows;
    unsigned int* counter;  // device word, see above
};
bool tiled_checks_long_rows(int variant);  // whether geometry `variant` has that form of the kernel
// TMA-staged tile kernel over the tile stream.  long_rows: null, or the device copy of the DevLongRows whose rows are to
// be checked in the same launch, with n_long_rows = its n_rows and *long_claims = the host's running count of claims
// (advanced by the call).
cudaError_t launch_r1cs_tiled(int field, const DevTileStream& ts, const fr_t* w, uint64_t row_base,
                              const CheckEpilogue& ep, fr_t* Aw, fr_t* Bw, fr_t* Cw, int sm_count, cudaStream_t s,
                              const DevLongRows* long_rows = nullptr, uint32_t n_long_rows = 0,
                              uint32_t* long_claims = nullptr);
// general values inside the blobs: canonical -> Montgomery in place; offs[i] = byte offset / 16 of value i
cudaError_t launch_to_mont_scattered(int field, uint8_t* blobs, const uint32_t* offs, uint64_t n, int* d_bad_flag,
                                     cudaStream_t s);
// structural validation on the device: rowptr monotone and ending at nnz, col < n_cols.  Sets *d_flag |= 2.
cudaError_t launch_validate_csr(const uint32_t* rowptr, const uint32_t* col, uint32_t n_rows, uint64_t nnz,
                                uint32_t n_cols, int* d_flag, cudaStream_t s);

// Witness generation (K6): gates in dependency-level order (host/circuit.hpp GatePlan has the same record).
struct WitnessGate {
    uint32_t kind;        // 1 Mul, 2 Equal, 3 Split
    uint32_t out;         // Mul / Equal: witness column of the output wire
    uint32_t l0, l1;      // Mul: left terms [l0, l1);  Split: outputs [l0, l1) in split_outs
    uint32_t r0, r1;      // Mul: right terms [r0, r1)
    uint32_t in;          // Equal / Split: witness column of the input wire
    uint32_t magic;       // Equal: witness column of the magic wire
};
// Evaluates the levels in order into w (Montgomery form, zero-initialised except constant and inputs).
cudaError_t launch_witness_levels(int field, const WitnessGate* gates, const uint32_t* level_ptr, uint32_t n_levels,
                                  uint32_t max_width, const uint32_t* term_col, const fr_t* term_coef,
                                  const uint32_t* split_outs, fr_t* w, int sm_count, cudaStream_t s);

// NTT (K3).  A plan owns the twiddle tables of one (field, log_n, direction).
struct NttPlan;
cudaError_t ntt_plan_create(int field, uint32_t log_n, bool inverse, NttPlan** out);
void ntt_plan_destroy(NttPlan* p);
// natural -> natural (DIF passes + bit-reversal permutation); inverse plans also scale by 1/n.
// data: batch * 2^log_n Montgomery elements (batch a power of two), scratch: same size.
// *launches is incremented by the number of kernels enqueued.
cudaError_t ntt_run(const NttPlan* p, fr_t* data, fr_t* scratch, uint32_t batch, cudaStream_t s, uint32_t* launches);
// natural -> bit-reversed, no scaling (building block of the QAP pipeline)
cudaError_t ntt_run_dif(const NttPlan* p, fr_t* data, uint32_t batch, cudaStream_t s, uint32_t* launches);

// QAP pointwise kernels (K4)
// out[i] = base^(i << shift), i < n
cudaError_t launch_fill_powers(int field, fr_t* out, uint32_t n, fr_t base, uint32_t shift, cudaStream_t s);
// v[i] *= pow_hi[e >> lo_bits] * pow_lo[e & mask] (* scale), e = i or bitrev_{log_n}(i)
cudaError_t launch_scale_by_powers(int field, fr_t* v, uint64_t n, const fr_t* d_pow_hi, const fr_t* d_pow_lo,
                                   uint32_t lo_bits, fr_t scale, bool use_scale, bool index_bitrev, uint32_t log_n,
                                   cudaStream_t s);
// h[i] = (a[i]*b[i] - c[i]) * zinv
cudaError_t launch_quotient_pointwise(int field, const fr_t* a, const fr_t* b, const fr_t* c, fr_t* h, uint64_t n,
                                      fr_t zinv, cudaStream_t s);
// out[i] = in[bitrev(i)] within each 2^log_n block
cudaError_t launch_bitrev_permute(const fr_t* in, fr_t* out, uint32_t log_n, uint32_t batch, cudaStream_t s);
// h[i] += d2*a[i] + d1*b[i]  (delta terms of verificationWitnessZk)
cudaError_t launch_axpy2(int field, fr_t* h, const fr_t* a, const fr_t* b, fr_t d2, fr_t d1, uint64_t n,
                         cudaStream_t s);

// out[i] = sum_k weight[k] * polys[k * len + i]
cudaError_t launch_poly_combine(int field, const fr_t* polys, const fr_t* weight, uint32_t n_polys, uint64_t len,
                                fr_t* out, cudaStream_t s);

// K7 (poly_kernels.cu): v[i] += s * t[i];  *d_flag |= 1 if any v[i] != 0;  exact division by the target:
// p (len_p coefficients, destroyed: p[0..n) ends as the remainder) by T (n + 1 coefficients, T[n] != 0,
// lc_inv = 1 / T[n], monic: T[n] == 1) -> h (len_p - n quotient coefficients).  Requires len_p > n.
cudaError_t launch_axpy1(int field, fr_t* v, const fr_t* t, fr_t s, uint64_t n, cudaStream_t st);
cudaError_t launch_any_nonzero(const fr_t* v, uint64_t n, int* d_flag, cudaStream_t s);
cudaError_t launch_poly_divmod(int field, fr_t* p, uint32_t len_p, const fr_t* T, uint32_t n, fr_t lc_inv, bool monic,
                               fr_t* h, cudaStream_t s, uint32_t* launches);

// Linear constraints over any 256-bit prime modulus (linear_kernels.cu): modulus 0 BN254 Fr, 1 BLS12-381 Fr,
// 2 the secp256k1 group order.  x[i] <- x[i] * 2^256 mod n in place (flag: an element >= n); then constraint i holds
// <=> lhs_i . x == rhs_i . v + cst_i (weights and constants canonical limbs, x_mont / v_mont from launch_lin_to_mont).
// d_result = {violated, first violated} accumulates: {0, ~0} on entry.
cudaError_t launch_lin_to_mont(int modulus, uint64_t* x, uint64_t n_el, int* d_bad_flag, cudaStream_t s);
cudaError_t launch_linear_constraints(int modulus, const uint32_t* l_rowptr, const uint32_t* l_col, const uint64_t* l_val,
                                      const uint32_t* r_rowptr, const uint32_t* r_col, const uint64_t* r_val,
                                      const uint64_t* cst, const uint64_t* x_mont, const uint64_t* v_mont,
                                      uint32_t n_constraints, unsigned long long* d_result, int* d_bad_flag,
                                      cudaStream_t s);

// Lagrange (K5): n <= 4096 distinct xs (Montgomery), n_polys value vectors -> coefficient vectors;
// target: n+1 coefficients of prod (X - x_i) (may be null).  d_status: set to 1 if two xs coincide.
cudaError_t launch_lagrange(int field, const fr_t* xs, const fr_t* ys, uint32_t n, uint32_t n_polys, fr_t* coeffs,
                            fr_t* target, fr_t* scratch /* 2(n+1) + n + n_polys*n elements */, int* d_status, cudaStream_t s,
                            uint32_t* launches);

}  // namespace acg
