/* acg.h -- C ABI of the B200-native R1CS / QAP hot path ("arithmetic-circuits on GPU").
 *
 * The reference (sdiehl/arithmetic-circuits @ 18e15de) is pure Haskell and has NO FFI today; the
 * drop-in boundary is the export list of module QAP (src/QAP.hs:11-39).  Each entry point below names
 * the reference function(s) whose numeric work it replaces; the Haskell module that binds them
 * (`foreign import ccall safe`) and keeps the reference's names/types is
 * arithmetic-circuits_b200/hs/QAP/GPU.hs, described in INTEGRATION.md.
 *
 * Conventions
 *   - Field element  = 4 little-endian uint64 limbs (32 bytes), CANONICAL residue in [0, r) on both
 *     sides of the ABI (`fromP` of galois-field's `Prime r`).  A non-canonical input is an error
 *     (ACG_ERR_NON_CANONICAL), never silently reduced.
 *   - Witness vector = dense w in qapSetToMap order (src/QAP.hs:605-620): index 0 the constant 1,
 *     then inputs, intermediates, outputs.
 *   - R1CS           = three CSR matrices A, B, C with one row per root, rows in ascending-root order
 *     (the Map key order of GenQAP (Map k) k, src/QAP.hs:94-99, 530-539), columns = witness indices.
 *   - Caller owns every host buffer; the library copies in/out and owns device memory behind opaque
 *     handles.  No callbacks, no exceptions, no exit() across the ABI.
 *   - Return value: 0 = ACG_OK, negative = error class.  AN INVALID WITNESS IS NOT AN ERROR
 *     (mirrors `Nothing`, src/QAP.hs:310-312): it is reported through the out-parameters.
 *   - One in-flight call per context.  Blocking calls return when the result is on the host.
 *     `_async` calls only enqueue on the given CUDA stream (cudaStream_t passed as void*).
 *     Enqueued checks are NOT replayable from a captured CUDA graph: every check carries per-call sequence numbers
 *     in its kernel arguments (result hand-over, long-row claims, peer exchange).
 *   - There is NO CPU fallback: without a CUDA device every compute call fails with ACG_ERR_NO_DEVICE.
 */
#ifndef ACG_H
#define ACG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACG_ABI_VERSION 1

/* field_id: the reference's type parameter `f` / `k`, monomorphised (FFI cannot be class-polymorphic) */
enum { ACG_FIELD_BN254_FR = 0,       /* Data.Pairing.BN254.Fr   (bench/Circuit.hs:10, test/Test/QAP.hs:12) */
       ACG_FIELD_BLS12_381_FR = 1,   /* BASELINE.json configs[4] */
       ACG_FIELD_SECP256K1_FN = 2 }; /* scalar field of secp256k1 (src/Circuit/Bulletproofs.hs): a modulus id of
                                      * acg_linear_constraints_check ONLY -- contexts exist for the first two */

enum {
    ACG_OK = 0,
    ACG_ERR_BAD_ARG = -1,        /* also the reference's `panic` cases, e.g. src/QAP.hs:445,474 */
    ACG_ERR_NON_CANONICAL = -2,  /* a field element >= r crossed the ABI */
    ACG_ERR_CUDA = -3,
    ACG_ERR_OOM = -4,
    ACG_ERR_NO_DEVICE = -5,
    ACG_ERR_UNSUPPORTED = -6,    /* e.g. log_n above the field's 2-adicity */
    ACG_ERR_INTERNAL = -7
};

typedef struct acg_ctx acg_ctx;          /* one CUDA device + one field */
typedef struct acg_r1cs acg_r1cs;        /* device-resident A, B, C (a row range of them) */
typedef struct acg_vec acg_vec;          /* device-resident vector of field elements (Montgomery form) */

/* CSR matrix on the host: rowptr[n_rows+1], col[nnz], val[4*nnz] canonical limbs. */
typedef struct {
    const uint32_t* rowptr;
    const uint32_t* col;
    const uint64_t* val;
    uint64_t nnz;
} acg_csr;

/* Device time (CUDA events) of the phases of the last call on this context, milliseconds. */
typedef struct {
    float h2d_ms;
    float kernel_ms;
    float d2h_ms;
    uint32_t kernel_launches;   /* launches of this library's own kernels in the last call */
    uint32_t reserved;
} acg_timing;

/* Which K2 kernel acg_r1cs_check uses (tuning / A-B measurement; results are identical). */
enum { ACG_CHECK_AUTO = 0, ACG_CHECK_ROWWISE = 1, ACG_CHECK_TILED = 2 };

/* ---- library / context --------------------------------------------------------------------------- */
int acg_abi_version(void);
const char* acg_strerror(int code);
/* Detail of the last failure on this context (empty string if none).  Valid until the next call. */
const char* acg_last_error(const acg_ctx* ctx);
/* device: CUDA ordinal.  Fails with ACG_ERR_NO_DEVICE when no usable GPU exists (no CPU fallback). */
int acg_ctx_create(int field_id, int device, acg_ctx** out);
void acg_ctx_destroy(acg_ctx* ctx);
int acg_ctx_set_check_kernel(acg_ctx* ctx, int which);
/* Tiled kernel tuning: tile geometry, bound to a system when it is uploaded.  0 = 128-row tiles (default),
 * 1 = 256, 2 = 64, 3 = 32 rows per tile. */
int acg_ctx_set_tiled_variant(acg_ctx* ctx, int variant);
/* Opt-in (default off): back-to-back checks (acg_r1cs_check*, tiled kernel) of the SAME system on the same stream
 * overlap -- the later check is launched as a programmatic dependent of the earlier one, moves onto the SMs as the
 * earlier one's CTAs run out of tiles, and waits for it only before it touches the shared result scratch.  The
 * witnesses may differ (a prover checking a stream of resident witnesses).  The library tracks its own operations: any
 * other call on the context between two checks -- in particular acg_witness_update* -- breaks the chain for that
 * pair.  What it cannot see is a caller rewriting a witness through acg_vec_device_ptr with its own kernels between
 * two checks: such a caller must leave this off. */
int acg_ctx_set_overlap_checks(acg_ctx* ctx, int on);
int acg_last_timing(const acg_ctx* ctx, acg_timing* out);
/* Total launches of this library's kernels on this context since creation. */
uint64_t acg_kernel_launch_count(const acg_ctx* ctx);
/* Per-launch device timing of the dominant check kernel: after acg_profile_begin, each of the next
 * max_launches acg_r1cs_check[_async] calls records a CUDA event pair around its main kernel on the
 * launching stream (no synchronisation).  acg_profile_end waits for them and returns the durations. */
int acg_profile_begin(acg_ctx* ctx, uint32_t max_launches);
int acg_profile_end(acg_ctx* ctx, float* ms_out, uint32_t capacity, uint32_t* n_out);

/* Field constants, host-only (no device needed): modulus, Montgomery R, R^2, -r^-1 mod 2^64, and
 * getRootOfUnity k (pairing-1.0.0; call sites Example.hs:26, bench/Circuit.hs:33). */
int acg_field_constants(int field_id, uint64_t modulus[4], uint64_t mont_r[4], uint64_t mont_r2[4],
                        uint64_t* ninv64, uint32_t* two_adicity);
int acg_root_of_unity(int field_id, uint32_t k, uint64_t out[4]);

/* ---- R1CS check: replaces verifyAssignment / verificationWitness's predicate ---------------------
 * src/QAP.hs:276-327 in its evaluation-domain form: valid <=> for every root g
 * (A.w)_g * (B.w)_g - (C.w)_g == 0, each dot product as src/Circuit/Affine.hs:121-125 (missing = 0). */

/* Upload rows [row_begin, row_end) of A, B, C (pass 0, n_rows for everything; a strict sub-range is a
 * shard for multi-GPU row partitioning).  CSR arrays describe the FULL matrices. */
int acg_r1cs_upload(acg_ctx* ctx, uint32_t n_rows, uint32_t n_cols, const acg_csr* A, const acg_csr* B,
                    const acg_csr* C, uint32_t row_begin, uint32_t row_end, acg_r1cs** out);
void acg_r1cs_free(acg_r1cs* m);
/* Host-only diagnostic (no device needed): builds the tile stream acg_r1cs_upload would build for rows
 * [row_begin, row_end) under tile geometry `variant` on n_threads host threads (0: as the upload chooses:
 * ACG_HOST_THREADS or the hardware concurrency) and returns out4 = {FNV-1a of the blobs, FNV-1a of the tile records +
 * far columns + value offsets + rows left to the long-row path, bytes of the blobs, number of tiles}.  Before hashing,
 * the stream is checked against everything the tiled kernel assumes about it (capacities, section order, every entry
 * and operand word inside its legal places of the CTA's shared memory, warp records, row permutation, hand-over
 * fields, far columns, row coverage): ACG_ERR_INTERNAL and a line on stderr if not.  The build runs
 * on worker threads over contiguous chunks of the tile list; this is how the tests show that its result does not depend
 * on the number of threads. */
int acg_tile_stream_digest(int field_id, int variant, uint32_t n_rows, uint32_t n_cols, const acg_csr* A, const acg_csr* B,
                           const acg_csr* C, uint32_t row_begin, uint32_t row_end, uint32_t n_threads, uint64_t* out4);
/* Algorithmic bytes one check of this (shard of the) system reads: SURVEY.md 8(d) formula
 * sum_M [nnz_M*(32+4) + 4*(rows+1)] + 32*(witness columns the rows reference) + 8.  For a whole system that is
 * 32*n_cols (every wire occurs in some row); a row shard of a larger system is charged only the witness elements its
 * own rows touch, not the whole replicated vector. */
uint64_t acg_r1cs_algorithmic_bytes(const acg_r1cs* m);
/* Bytes of the device-side tile stream the tiled kernel actually reads per check besides the witness (per tile: a
 * header, one 32-bit word per ELL slot, the far witness columns, and the values of the general coefficients only;
 * +-1 coefficients are sign bits of the words). */
uint64_t acg_r1cs_stream_bytes(const acg_r1cs* m);
/* A row block uploaded as a system of its own (CSR arrays of the block's rows only, row_begin = 0): `offset` = global
 * index of its first row, added to every first_bad_row this system reports -- so that the all-reduced minimum over
 * row-sharded ranks is a global row.  (The alternative is to pass the full matrices and a row range to
 * acg_r1cs_upload.) */
int acg_r1cs_set_row_offset(acg_r1cs* m, uint64_t offset);

/* w: n_cols canonical elements in qapSetToMap order (src/QAP.hs:605-620). */
int acg_witness_upload(acg_ctx* ctx, const uint64_t* w, uint32_t n_cols, acg_vec** out);
/* Overwrite an existing device witness from host memory (same length).  Blocking; a rejected update (an element >= r)
 * leaves the vector unchanged. */
int acg_witness_update(acg_ctx* ctx, acg_vec* v, const uint64_t* w, uint32_t n_cols);
/* Overwrite elements [first, first + count) only.  With row shards on several GPUs every rank uploads its own slice
 * of a new witness over its own PCIe link and the slices are then exchanged device to device over NVLink
 * (sharding.upload_witness_sliced), instead of every rank pulling the whole vector from the host. */
int acg_witness_update_range(acg_ctx* ctx, acg_vec* v, const uint64_t* w, uint32_t first, uint32_t count);
/* Enqueue-only variant for pipelining: elements [first, first + count) are copied on the context's COPY stream (H2D
 * straight into the vector, range check + Montgomery conversion in place).  The update waits for the checks of v that
 * were enqueued before it, and checks of v enqueued after it wait for the update -- with two vectors the upload of
 * witness i + 1 overlaps the check of witness i.  `w` must stay valid (pinned memory, or the copy is staged by the
 * driver) until the next blocking call on v.  An element >= r is reported by that call (acg_r1cs_check,
 * acg_vec_status) as ACG_ERR_NON_CANONICAL; the vector's contents are then undefined -- acg_witness_update[_range],
 * by contrast, leave the vector untouched when they reject an update. */
int acg_witness_update_async(acg_ctx* ctx, acg_vec* v, const uint64_t* w, uint32_t first, uint32_t count);
/* Enqueue only: `stream` (a cudaStream_t) waits for the asynchronous updates of v enqueued so far -- for callers that
 * complete the vector with their own device work (the NVLink all-gather of a row-sharded witness) before a check. */
int acg_vec_stream_wait(acg_ctx* ctx, acg_vec* v, void* stream);
/* Blocking: waits for the asynchronous updates of v; ACG_ERR_NON_CANONICAL if one of them met an element >= r. */
int acg_vec_status(acg_ctx* ctx, acg_vec* v);
void acg_vec_free(acg_vec* v);
uint32_t acg_vec_len(const acg_vec* v);
/* Raw device pointer of the vector's storage (Montgomery form), for zero-copy interop. */
void* acg_vec_device_ptr(acg_vec* v);

/* Blocking.  n_violations = number of rows with a non-zero residual; first_bad_row = smallest such
 * GLOBAL row index, or UINT64_MAX when valid.  valid <=> *n_violations == 0  (verifyAssignment). */
int acg_r1cs_check(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, uint64_t* n_violations,
                   uint64_t* first_bad_row);
/* Enqueue only: d_result points to 2 device uint64 {n_violations, first_bad_row}; no initialisation needed, valid once
 * the check has completed on `stream` (its first block writes {0, none}, warps that find violated rows update the
 * pair, nothing runs at the end of a clean check).  Checks of ONE context share its scratch: order them on the device
 * (one stream, or events between streams).  For timing the kernels with CUDA events on `stream`, for back-to-back checks
 * (see acg_ctx_set_overlap_checks) and for callers that reduce the pair across row shards themselves
 * (acg_r1cs_check_async_allreduce does it inside the kernel). */
int acg_r1cs_check_async(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, uint64_t* d_result,
                         void* stream);
/* One-shot from host buffers (upload + check + read-back): the end-to-end call a Haskell wrapper of
 * verifyAssignmentR1CS makes. */
int acg_r1cs_check_host(acg_ctx* ctx, uint32_t n_rows, uint32_t n_cols, const acg_csr* A, const acg_csr* B,
                        const acg_csr* C, const uint64_t* w, uint64_t* n_violations, uint64_t* first_bad_row);
/* A.w, B.w, C.w for the uploaded rows, canonical, to host buffers of 4*rows limbs (any may be NULL). */
int acg_r1cs_eval(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, uint64_t* Aw, uint64_t* Bw,
                  uint64_t* Cw);

/* ---- multi-GPU row shards: one process per GPU, each holding rows [row_begin, row_end) and the whole
 * witness.  The only exchange is the {violation count, first bad row} pair; it travels over peer memory
 * (NVLink / NVSwitch, CUDA IPC) instead of a collective library call:
 *   1. every rank: acg_peer_create -> a 64-byte handle of its exchange buffer;
 *   2. the caller all-gathers the handles (any transport; the Python binding uses torch.distributed);
 *   3. every rank: acg_peer_connect with all `world` handles (rank order);
 *   4. acg_r1cs_check_async_allreduce enqueues the shard's check; the last CTA of the check kernel to finish stores
 *      this rank's pair into every peer's buffer (system-scope stores, release-ordered sequence number) and sums /
 *      mins the world's pairs from its own buffer: d_result (device, 2 x uint64) then holds the GLOBAL count and
 *      first bad row on every rank -- one kernel launch per check, no collective call.  All ranks must make the
 *      call (like a collective).
 *      A peer that does not arrive within ~4 s yields count = UINT64_MAX instead of a hang.
 * world <= 8 (one NVSwitch node). */
#define ACG_PEER_HANDLE_BYTES 64
typedef struct acg_peer acg_peer;
int acg_peer_create(acg_ctx* ctx, uint32_t world, uint32_t rank, acg_peer** out, uint8_t* handle_out);
int acg_peer_connect(acg_ctx* ctx, acg_peer* p, const uint8_t* handles /* world * ACG_PEER_HANDLE_BYTES */);
void acg_peer_free(acg_peer* p);
int acg_r1cs_check_async_allreduce(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, acg_peer* peer,
                                   uint64_t* d_result, void* stream);

/* ---- NTT: replaces FFT.interpolate / the DFT of galois-fft (src/QAP.hs:521-523) -------------------
 * In place on 2^log_n canonical elements, natural order in and out.  inverse=0: out[i] = sum_j
 * in[j] w^(ij); inverse=1: the inverse (scaled by 1/n), w = getRootOfUnity log_n. */
int acg_ntt(acg_ctx* ctx, uint64_t* data, uint32_t log_n, int inverse);
/* Same on a device vector (Montgomery form), enqueue only. */
int acg_ntt_device(acg_ctx* ctx, acg_vec* v, uint32_t log_n, int inverse, void* stream);
/* createPolynomialsFFT's per-wire work (src/QAP.hs:512-525): n_cols_batch columns, each 2^log_n
 * canonical values in ascending-root order (zero-padded by the caller), replaced in place by the
 * coefficients of the interpolating polynomial (little-endian, length 2^log_n, NOT stripped). */
int acg_interpolate_columns(acg_ctx* ctx, uint64_t* cols, uint32_t log_n, uint32_t n_cols_batch);

/* ---- QAP witness polynomials: replaces verificationWitnessZk (src/QAP.hs:300-327) ------------------
 * on the FFT-built QAP of arithCircuitToQAPFFT with N = 2^ceil(log2 n_rows) and T = X^N - 1, using
 * the linearity collapse sum_k w_k * interpolate(col_k) = interpolate(A.w).
 * delta = {delta1, delta2, delta3} (12 limbs; NULL = zeros).  Outputs (each N+1 elements, canonical,
 * little-endian coefficients, not stripped; any may be NULL): a = delta1*T + sum w_k A_k, b, c
 * likewise, h = (a*b - c) / T.  *divisible = 1 iff the remainder is zero (then h is the reference's
 * `Just quotient`); when 0 the reference returns Nothing and h is unspecified.
 * Requires the full system (row_begin = 0, row_end = n_rows). */
int acg_qap_witness(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, const uint64_t* delta,
                    uint64_t* a, uint64_t* b, uint64_t* c, uint64_t* h, int* divisible);

/* ---- Lagrange: replaces createPolynomials.lagrangeInterpolate (src/QAP.hs:495-508) ----------------
 * n distinct canonical xs, n_polys value vectors ys (n_polys*n elements) -> n_polys coefficient
 * vectors (n each, little-endian, not stripped); target (n+1 coefficients, may be NULL) =
 * prod (X - x_i) (src/QAP.hs:492).  O(n^2) per polynomial; n <= 4096. */
int acg_lagrange(acg_ctx* ctx, const uint64_t* xs, const uint64_t* ys, uint32_t n, uint32_t n_polys,
                 uint64_t* coeffs, uint64_t* target);

/* ---- per-wire QAP: the scale-and-sum of verificationWitnessZk (src/QAP.hs:314-324; foldQapSet,
 * combineWithDefaults :163-181) over polynomials held per wire, as createPolynomials[FFT] returns them:
 * out[i] = sum_k weights[k] * polys[k*len + i].  polys: n_polys coefficient vectors of length len (zero-padded),
 * weights: the witness value of each wire.  Canonical limbs in and out; blocking. */
int acg_poly_combine(acg_ctx* ctx, const uint64_t* polys, const uint64_t* weights, uint32_t n_polys, uint32_t len,
                     uint64_t* out);

/* ---- verificationWitnessZk on a per-wire `QAP f` value (src/QAP.hs:74-79, 300-327), whole on the device --------
 * acg_qap_upload keeps the three polynomial sets and the target resident: left / right / out are n_wires coefficient
 * vectors of `len` canonical elements each (little-endian degree, zero padded), wires in qapSetToMap order
 * (src/QAP.hs:605-620; a wire the QAP lacks is the zero polynomial, as combineWithDefaults treats it, :163-181);
 * target: n_target coefficients of qapTarget -- prod (X - root) of createPolynomials (:492), FFT.fftTargetPoly of
 * createPolynomialsFFT (:524), or any other non-zero polynomial (trailing zeros are stripped).
 * acg_qap_verify: a = delta1*T + sum_k w_k L_k, b, c likewise (scale-and-sum, :314-324), p = a*b - c by a product on
 * >= 2*max(len, n_target) - 1 points (NTT), then (h, rem) = p divMod T by long division on the device (:327, K7).
 * w: n_wires canonical witness values (missing wires 0); delta: 12 limbs or NULL.  *divisible = 1 iff rem == 0 -- then
 * h[0 .. *h_len) (canonical, little-endian, NOT stripped; *h_len = acg_qap_quotient_len) is the reference's
 * `Just quotient`; otherwise the reference returns Nothing.  h may be NULL (verifyAssignment only needs the flag). */
typedef struct acg_qap acg_qap;
int acg_qap_upload(acg_ctx* ctx, const uint64_t* left, const uint64_t* right, const uint64_t* out, uint32_t n_wires,
                   uint32_t len, const uint64_t* target, uint32_t n_target, acg_qap** out_qap);
void acg_qap_free(acg_qap* q);
uint32_t acg_qap_quotient_len(const acg_qap* q);
int acg_qap_verify(acg_ctx* ctx, const acg_qap* q, const uint64_t* w, const uint64_t* delta, uint64_t* h,
                   uint32_t h_capacity, uint32_t* h_len, int* divisible);
/* FFT.fftTargetPoly primRoots n (src/QAP.hs:524; galois-fft-0.1.0, not vendored): prod_{i < n_roots} (X - w^i) with w
 * the primitive 2^ceil(log2 n_roots)-th root of unity of the field -- X^N - 1 when n_roots is a power of two.
 * out: n_roots + 1 canonical coefficients.  A partial domain is built on the device (K5 master polynomial) and is
 * limited to 4096 roots. */
int acg_fft_target(acg_ctx* ctx, uint32_t n_roots, uint64_t* out);

/* ---- linear constraints: replaces checkLinearConstraint of the Bulletproofs backend ------------------------------
 * (src/Circuit/Bulletproofs.hs:329-338):  wL.aL + wR.aR + wO.aO == wV.v + c  over the scalar field of secp256k1, each
 * dot product as src/Circuit/Affine.hs:121-125 (missing = 0).  n_constraints constraints at once: `lhs` is the CSR
 * matrix of the weights over the concatenated assignment x = [aL | aR | aO] (n_lhs_vars elements), `rhs` that of wV over
 * v (n_rhs_vars), `constants` the c of every constraint; all canonical limbs.  Constraint i holds <=> lhs_i . x ==
 * rhs_i . v + c_i.  *n_violations / *first_bad as for acg_r1cs_check.  modulus_id: any of the three ids above (the
 * context's own field does not matter: secp256k1's order is >= 2^255, which the main path's Montgomery arithmetic does not
 * cover -- this entry point uses its own textbook 4 x 64-bit-limb arithmetic, linear_kernels.cu).  Blocking. */
int acg_linear_constraints_check(acg_ctx* ctx, int modulus_id, uint32_t n_constraints, uint32_t n_lhs_vars,
                                 uint32_t n_rhs_vars, const acg_csr* lhs, const acg_csr* rhs, const uint64_t* constants,
                                 const uint64_t* x, const uint64_t* v, uint64_t* n_violations, uint64_t* first_bad);

/* ---- field ops on the device (K1 self-test surface) ------------------------------------------------
 * op: 0 add, 1 sub, 2 mul, 3 inverse of a (inv 0 = 0, as evalGate treats it, Arithmetic.hs:130).
 * n canonical elements each; blocking. */
int acg_fr_binop(acg_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t n);

/* ================================================================================================
 * Host side of the path (C++ behind this ABI; no GPU needed): circuit IR, witness generation and
 * R1CS row construction, mirroring the reference functions so a caller can go ArithCircuit ->
 * {QapSet, CSR} -> GPU without Haskell in the loop.
 * ================================================================================================ */
typedef struct acg_circuit acg_circuit;       /* ArithCircuit f,  src/Circuit/Arithmetic.hs:149-150 */
typedef struct acg_assignment acg_assignment; /* QapSet f,        src/QAP.hs:66-71 */
typedef struct acg_r1cs_host acg_r1cs_host;   /* GenQAP (Map k) k as CSR, src/QAP.hs:94-99 */

/* Wire encoding (src/Circuit/Arithmetic.hs:32-36): kind in the top 2 bits, index below. */
#define ACG_WIRE_INPUT 0u
#define ACG_WIRE_INTERMEDIATE 1u
#define ACG_WIRE_OUTPUT 2u
#define ACG_WIRE(kind, ix) ((((uint64_t)(kind)) << 62) | (uint64_t)(ix))

/* Circuit = flat stream of uint64 words (what a Haskell `ArithCircuit Fr -> [Word64]` marshaller emits):
 *   Mul l r out      : 1, out, n_l, <l tokens>, n_r, <r tokens>      (n_* = number of WORDS)
 *   Equal i m out    : 2, i, m, out
 *   Split i outs     : 3, i, n_outs, outs...
 * Affine circuit tokens in POST-ORDER (src/Circuit/Affine.hs:26-31):
 *   Var w : 0, w      ConstGate f : 1, f0..f3      Add : 2 (pops r, l)      ScalarMul f : 3, f0..f3 (pops 1)
 */
int acg_circuit_parse(int field_id, const uint64_t* words, uint64_t n_words, acg_circuit** out);
void acg_circuit_free(acg_circuit* c);
uint64_t acg_circuit_num_gates(const acg_circuit* c);
/* validArithCircuit, src/Circuit/Arithmetic.hs:158-185: returns 1/0. */
int acg_circuit_valid(const acg_circuit* c);
/* Total number of roots generateRoots draws (src/Circuit/Arithmetic.hs:194-216) = number of R1CS rows. */
uint64_t acg_circuit_num_roots(const acg_circuit* c);

/* generateAssignment (src/QAP.hs:597-603 -> evalArithCircuit, Arithmetic.hs:221-235).
 * inputs: n_inputs pairs (index, canonical value). */
int acg_generate_assignment(const acg_circuit* c, const uint32_t* input_ix, const uint64_t* input_vals,
                            uint32_t n_inputs, acg_assignment** out);
void acg_assignment_free(acg_assignment* a);
/* Sizes of the QapSet's three maps as qapSetToMap counts them (maxKey+1) and number of present keys. */
int acg_assignment_dims(const acg_assignment* a, uint32_t* n_in, uint32_t* n_mid, uint32_t* n_out);
/* Look one wire up (lookupAtWire, src/QAP.hs:331-337): returns 1 and fills out when present, else 0. */
int acg_assignment_lookup(const acg_assignment* a, uint64_t wire, uint64_t out[4]);
/* Replace/insert a value (updateAtWire, src/QAP.hs:341-347) -- used to build faulty assignments. */
int acg_assignment_update(acg_assignment* a, uint64_t wire, const uint64_t val[4]);
/* qapSetToMap (src/QAP.hs:605-620) densified: fills w[4*n_cols]; wires not present are 0.  The block
 * sizes are those of the LAYOUT (n_in, n_mid, n_out), which must cover the assignment's keys. */
int acg_assignment_to_vector(const acg_assignment* a, uint32_t n_in, uint32_t n_mid, uint32_t n_out,
                             uint64_t* w);

/* arithCircuitToGenQAP (src/QAP.hs:530-539) without addMissingZeroes' densification: gateToGenQAP per
 * gate (src/QAP.hs:366-474), rows sorted by ascending root.  roots: acg_circuit_num_roots canonical
 * elements in generateRoots order, or NULL for `fromIntegral <$> fresh` starting at root_start
 * (bench/Circuit.hs:31 uses 0, Example.hs:24 uses 1).  Layout block sizes as above; pass zeros to let
 * the library derive them from the circuit's wires. */
int acg_circuit_to_r1cs(const acg_circuit* c, const uint64_t* roots, uint64_t root_start, uint32_t n_in,
                        uint32_t n_mid, uint32_t n_out, acg_r1cs_host** out);
void acg_r1cs_host_free(acg_r1cs_host* m);
int acg_r1cs_host_dims(const acg_r1cs_host* m, uint32_t* n_rows, uint32_t* n_cols, uint32_t* n_in,
                       uint32_t* n_mid, uint32_t* n_out);
/* Borrow the CSR arrays (valid until acg_r1cs_host_free).  which: 0 = A, 1 = B, 2 = C. */
int acg_r1cs_host_csr(const acg_r1cs_host* m, int which, acg_csr* out);
/* Sorted roots, one per row (4*n_rows limbs). */
const uint64_t* acg_r1cs_host_roots(const acg_r1cs_host* m);

/* ---- witness generation on the device (K6): generateAssignment (src/QAP.hs:597-603) = the evalGate fold of
 * evalArithCircuit (src/Circuit/Arithmetic.hs:106-145, 221-235), evaluated by dependency level on the GPU.
 * Needs a context (device); the result is a device-resident witness vector in qapSetToMap order
 * (src/QAP.hs:605-620) with layout (max(n_in, circuit), max(n_mid, circuit), max(n_out, circuit)) -- pass zeros
 * to take the circuit's own sizes -- ready for acg_r1cs_check / acg_qap_witness without a host round trip.
 * The gate list must be in single-assignment, define-before-use form (what validArithCircuit accepts and every
 * circuit of the reference's tests and examples is); otherwise ACG_ERR_UNSUPPORTED and the caller uses the
 * sequential acg_generate_assignment.  *n_levels (optional) = number of dependency levels = barrier count. */
int acg_generate_assignment_device(acg_ctx* ctx, const acg_circuit* c, const uint32_t* input_ix,
                                   const uint64_t* input_vals, uint32_t n_inputs, uint32_t n_in, uint32_t n_mid,
                                   uint32_t n_out, acg_vec** out, uint32_t* n_levels);
/* Host-only: the dependency levels acg_generate_assignment_device would use (number of levels = barriers, widest
 * level).  ACG_ERR_UNSUPPORTED / ACG_ERR_BAD_ARG exactly when the device entry point would refuse the circuit. */
int acg_circuit_plan_stats(const acg_circuit* c, uint32_t* n_levels, uint32_t* max_width);
/* Copy a device vector back as canonical limbs (n must equal acg_vec_len). */
int acg_vec_download(acg_ctx* ctx, const acg_vec* v, uint64_t* out, uint32_t n);

/* Synthetic family S(n, seed, field) of SURVEY.md 8(d) (bench / parity workloads): n Mul gates over
 * 1024 inputs, ~5.5 nnz per row.  Returns the lowered system and its honest witness directly
 * (same result as parse -> generate_assignment -> to_r1cs on the equivalent ArithCircuit, which
 * acg_synth_circuit_words emits for cross-checking at small n).  dense != 0: every coefficient a
 * uniform field element. */
int acg_synth_r1cs(int field_id, uint32_t n, uint64_t seed, int dense, acg_r1cs_host** out_m,
                   uint64_t** out_w /* malloc'ed 4*n_cols limbs; free with acg_free */);
/* Same system, but only rows [row_begin, row_end) are materialised (a shard for multi-GPU runs): the
 * returned matrices have row_end-row_begin rows; the witness is complete. */
int acg_synth_r1cs_rows(int field_id, uint32_t n, uint64_t seed, int dense, uint32_t row_begin, uint32_t row_end,
                        acg_r1cs_host** out_m, uint64_t** out_w);
int acg_synth_circuit_words(int field_id, uint32_t n, uint64_t seed, int dense, uint64_t** out_words,
                            uint64_t* out_n_words, uint32_t** out_input_ix, uint64_t** out_input_vals,
                            uint32_t* out_n_inputs);
void acg_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ACG_H */
