// C ABI of the device path (include/acg.h): contexts, uploads, the R1CS check, NTT and the QAP
// witness pipeline.  Host logic only orchestrates; all arithmetic on bulk data runs in the kernels
// of r1cs_kernels.cu / ntt_kernels.cu / lagrange_kernels.cu.  There is no CPU fallback: every compute
// entry point needs a CUDA device and reports ACG_ERR_NO_DEVICE / ACG_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/acg.h"
#include "host/circuit.hpp"
#include "kernels.h"

using namespace acg;

// --------------------------------------------------------------------------------------------------
// handles
// --------------------------------------------------------------------------------------------------
struct acg_ctx {
    int field = 0;
    int device = 0;
    int sm_count = 148;
    int check_kernel = ACG_CHECK_AUTO;
    int tiled_variant = 0;  // index into kTileGeom, bound to a system at upload
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // acg_witness_update_async: H2D + conversion, overlapping the checks
    fr_t* staging = nullptr;                 // acg_witness_update[_range]: a rejected update must not touch the vector
    size_t staging_cap = 0;                  //   (elements)
    // work buffers of acg_qap_witness, kept between calls (seven vectors of N elements: a fresh cudaMalloc / cudaFree
    // pair per buffer and call cost more than the kernels at N = 2^22)
    // where the block scheduler puts block b of the tiled kernel's full grid (SM id per block), probed once per tile
    // geometry (kernels.h CtaRun); empty after a probe that did not see every SM filled evenly
    std::vector<uint32_t> placement[kNumTileVariants];
    bool placement_probed[kNumTileVariants] = {};
    fr_t* work[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t work_cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long* d_result = nullptr;  // {n_violations, first_bad_row}
    unsigned long long* d_accum = nullptr;   // scratch pair the check kernels accumulate into + CTA ticket (u32)
    unsigned int* d_ticket = nullptr;        //   both reset by the finalising CTA of every check (CheckEpilogue)
    unsigned long long* d_gate = nullptr;    // ring of kGateRing gate words of the direct hand-over (CheckEpilogue::gate)
    unsigned long long gate_seq = 0;         //   .. and the running number of the checks that used it
    bool direct_handover = true;             // ACG_K2_TICKET=1 (a measurement aid): always the ticket path
    int* d_flag = nullptr;
    unsigned long long* h_result = nullptr;  // pinned
    int* h_flag = nullptr;                   // pinned
    std::string err;
    acg_timing timing{};
    uint64_t launches = 0;
    std::map<std::pair<uint32_t, int>, NttPlan*> plans;
    struct CosetTables {
        fr_t *hi = nullptr, *lo = nullptr, *ihi = nullptr, *ilo = nullptr;
        uint32_t lo_bits = 0;
    };
    std::map<uint32_t, CosetTables> coset;
    // optional per-launch device timing of the main check kernel (acg_profile_*)
    std::vector<cudaEvent_t> prof_ev;  // pairs
    uint32_t prof_used = 0;
    // overlap of consecutive checks (CheckEpilogue::overlap): allowed when the previous operation of this context
    // was a single-launch tiled check of the same system and witness on the same stream
    int overlap_checks = 0;
    const void* last_m = nullptr;
    cudaStream_t last_stream = nullptr;
    uint64_t last_check_op = 0;  // value of `ops` right after that check
    uint64_t ops = 0;            // bumped by every entry point that enqueues device work
};

struct acg_r1cs {
    acg_ctx* ctx = nullptr;
    uint32_t n_rows_total = 0, n_cols = 0, row_begin = 0, row_end = 0;
    uint64_t row_offset = 0;  // added to reported rows: a shard uploaded as a system of its own (acg_r1cs_set_row_offset)
    uint64_t nnz[3] = {0, 0, 0};
    uint64_t distinct_cols = 0;  // witness columns referenced by the rows of this shard
    uint32_t* d_rowptr[3] = {nullptr, nullptr, nullptr};
    uint32_t* d_col[3] = {nullptr, nullptr, nullptr};
    fr_t* d_val[3] = {nullptr, nullptr, nullptr};
    DevR1cs dev{};
    // execution-ready tile stream of the tiled kernel (kernels.h), built for geometry `variant`
    uint8_t* d_stream = nullptr;
    TileMeta* d_meta = nullptr;
    uint32_t* d_far_cols = nullptr;
    uint64_t stream_bytes = 0;
    uint64_t blob_bytes = 0;  // of d_stream
    uint32_t n_tiles = 0;
    int variant = 0;
    std::vector<std::pair<uint32_t, uint32_t>> long_ranges;  // local row ranges too wide for a tile
    CtaRun* d_runs = nullptr;         // weighted runs of the tiled kernel's CTAs (kernels.h CtaRun), or null
    uint32_t n_runs = 0;
    // the plan behind d_runs: the blocks in the order their runs follow each other in the stream, the run length of
    // every block, and the tile records (a run carries its first tile's)
    std::vector<uint32_t> run_order, run_cnt;
    std::vector<TileMeta> h_meta;
    std::vector<uint32_t> h_far_cols;  // host copy of d_far_cols
    uint32_t* d_run_far = nullptr;     // DevTileStream::run_far
    uint32_t* d_long_rows = nullptr;  // .. flattened: the rows the warp-per-row kernel handles in one launch
    uint32_t n_long_rows = 0;
    DevLongRows* d_long_desc = nullptr;  // kernels.h DevLongRows on the device: the tiled kernel checks them itself
    unsigned int* d_long_counter = nullptr;  //   .. claiming them from this counter,
    mutable uint32_t long_claims = 0;        //   .. which has been advanced this far by the checks enqueued so far
};

struct acg_vec {
    acg_ctx* ctx = nullptr;
    fr_t* d = nullptr;
    uint32_t n = 0;
    // asynchronous updates (acg_witness_update_async): `ready` is recorded on the copy stream behind the update, every
    // check of the vector waits for it and records `used` behind itself, which the next update waits for
    cudaEvent_t ready = nullptr, used = nullptr;
    int* d_bad = nullptr;          // set by the conversion kernel of an asynchronous update: an element was >= r
    bool async_pending = false;    // an asynchronous update was enqueued and its verdict not yet read back
    bool has_used = false;
};

struct acg_peer {  // exchange buffers of a group of row-shard ranks (one process per GPU, CUDA IPC)
    acg_ctx* ctx = nullptr;
    uint32_t world = 0, rank = 0;
    unsigned long long* local = nullptr;  // this rank's buffer (cudaMalloc)
    void* mapped[kMaxPeers] = {};         // peers' buffers as opened here (null for self / not connected)
    PeerSlots slots{};
    unsigned long long** d_table = nullptr;  // PeerSlots::base
    unsigned long long seq = 0;
    bool connected = false;
};

namespace {

int fail(acg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
int acg_guard_fail(acg_ctx* ctx, int code, const char* msg) noexcept {
    try {
        if (ctx) ctx->err = msg;
    } catch (...) {
    }
    return code;
}
int fail_cuda(acg_ctx* ctx, cudaError_t e, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    if (ctx) ctx->err = buf;
    if (e == cudaErrorMemoryAllocation) return ACG_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return ACG_ERR_NO_DEVICE;
    return ACG_ERR_CUDA;
}

// No C++ exception crosses the C ABI (include/acg.h): every int-returning entry point runs inside ACG_TRY / ACG_CATCH,
// which map std::bad_alloc to ACG_ERR_OOM and anything else to ACG_ERR_INTERNAL.
#define ACG_TRY try {
#define ACG_CATCH_IMPL(ctxp)                                                                            \
    }                                                                                                   \
    catch (const std::bad_alloc&) { return acg_guard_fail((ctxp), ACG_ERR_OOM, "out of host memory"); } \
    catch (const std::exception& e__) { return acg_guard_fail((ctxp), ACG_ERR_INTERNAL, e__.what()); }  \
    catch (...) { return acg_guard_fail((ctxp), ACG_ERR_INTERNAL, "unknown C++ exception"); }
#define ACG_CATCH(ctxp) ACG_CATCH_IMPL(const_cast<acg_ctx*>(static_cast<const acg_ctx*>(ctxp)))
#define ACG_CATCH_NOCTX() ACG_CATCH_IMPL(nullptr)
#define CU(ctx, expr)                                              \
    do {                                                           \
        cudaError_t e__ = (expr);                                  \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #expr); \
    } while (0)

struct DevBuf {  // scoped device allocation
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T>
    T* as() const {
        return static_cast<T*>(p);
    }
};

template <class F>
int with_field(int field, F&& f) {
    if (field == ACG_FIELD_BN254_FR) return f(Bn254Fr{});
    if (field == ACG_FIELD_BLS12_381_FR) return f(Bls12381Fr{});
    return ACG_ERR_BAD_ARG;
}

template <class P>
fr_t host_pow(fr_t base, uint64_t e) {
    fr_t acc = fr_one<P>();
    while (e) {
        if (e & 1ull) acc = fr_mul<P>(acc, base);
        base = fr_sqr<P>(base);
        e >>= 1;
    }
    return acc;
}
template <class P>
fr_t host_const(uint32_t (*f)(int)) {
    fr_t r;
    for (int i = 0; i < 8; ++i) r.l[i] = f(i);
    return r;
}
template <class P>
fr_t host_root_of_unity(uint32_t k) {  // Montgomery form
    fr_t w = host_const<P>(&P::two_adic_root);
    for (uint32_t i = k; i < (uint32_t)P::TWO_ADICITY; ++i) w = fr_sqr<P>(w);
    return w;
}
void limbs_from_fr(uint64_t out[4], const fr_t& v) {
    for (int i = 0; i < 4; ++i) out[i] = (uint64_t)v.l[2 * i] | ((uint64_t)v.l[2 * i + 1] << 32);
}
fr_t fr_from_limbs(const uint64_t in[4]) {
    fr_t v;
    for (int i = 0; i < 4; ++i) {
        v.l[2 * i] = (uint32_t)in[i];
        v.l[2 * i + 1] = (uint32_t)(in[i] >> 32);
    }
    return v;
}

int two_adicity(int field) { return field == 0 ? Bn254Fr::TWO_ADICITY : Bls12381Fr::TWO_ADICITY; }

int activate(acg_ctx* ctx) {
    if (!ctx) return ACG_ERR_BAD_ARG;
    ctx->err.clear();
    ++ctx->ops;
    CU(ctx, cudaSetDevice(ctx->device));
    return ACG_OK;
}

int get_plan(acg_ctx* ctx, uint32_t log_n, bool inverse, NttPlan** out) {
    auto key = std::make_pair(log_n, inverse ? 1 : 0);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) {
        *out = it->second;
        return ACG_OK;
    }
    if ((int)log_n > two_adicity(ctx->field))
        return fail(ctx, ACG_ERR_UNSUPPORTED, "log_n exceeds the 2-adicity of the field");
    NttPlan* p = nullptr;
    CU(ctx, ntt_plan_create(ctx->field, log_n, inverse, &p));
    ctx->launches += 3;
    ctx->plans[key] = p;
    *out = p;
    return ACG_OK;
}

// convert n Montgomery elements at d (in place), copy them to host
int download_canonical(acg_ctx* ctx, fr_t* d, uint64_t n, uint64_t* host) {
    CU(ctx, launch_from_mont(ctx->field, d, n, ctx->stream));
    ctx->launches += 1;
    CU(ctx, cudaMemcpyAsync(host, d, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    return ACG_OK;
}
// copy n canonical elements to d, validate, convert to Montgomery.  Synchronises.
int upload_canonical(acg_ctx* ctx, fr_t* d, const uint64_t* host, uint64_t n) {
    if (n == 0) return ACG_OK;
    CU(ctx, cudaMemcpyAsync(d, host, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    CU(ctx, launch_to_mont(ctx->field, d, n, ctx->d_flag, ctx->stream));
    ctx->launches += 1;
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag) return fail(ctx, ACG_ERR_NON_CANONICAL, "field element >= modulus");
    return ACG_OK;
}

// Same, but the destination is only written when every element is canonical: H2D into the context's staging buffer,
// validation + conversion there, then a device-side commit (copy unless the flag is set).  One synchronisation.
int upload_canonical_staged(acg_ctx* ctx, fr_t* d, const uint64_t* host, uint64_t n) {
    if (n == 0) return ACG_OK;
    if (ctx->staging_cap < n) {
        if (ctx->staging) cudaFree(ctx->staging);
        ctx->staging = nullptr;
        ctx->staging_cap = 0;
        CU(ctx, cudaMalloc(&ctx->staging, n * sizeof(fr_t)));
        ctx->staging_cap = n;
    }
    CU(ctx, cudaMemcpyAsync(ctx->staging, host, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    CU(ctx, launch_to_mont(ctx->field, ctx->staging, n, ctx->d_flag, ctx->stream));
    CU(ctx, launch_copy_if_clean(d, ctx->staging, n, ctx->d_flag, ctx->stream));
    ctx->launches += 2;
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_flag) return fail(ctx, ACG_ERR_NON_CANONICAL, "field element >= modulus (the vector was left unchanged)");
    return ACG_OK;
}

// A check is about to read vector v on stream s / has been enqueued there (see acg_vec)
int vec_acquire(acg_ctx* ctx, const acg_vec* v, cudaStream_t s) {
    if (v->ready && v->async_pending) CU(ctx, cudaStreamWaitEvent(s, v->ready, 0));
    return ACG_OK;
}
int vec_release(acg_ctx* ctx, const acg_vec* v, cudaStream_t s) {
    if (v->used) {
        CU(ctx, cudaEventRecord(v->used, s));
        const_cast<acg_vec*>(v)->has_used = true;
    }
    return ACG_OK;
}

struct WorkBuf {  // a view of one of the context's work buffers (same accessors as DevBuf, no ownership)
    fr_t* p = nullptr;
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};
// work buffer `slot` of the context with room for n elements (contents undefined)
int work_buffer(acg_ctx* ctx, int slot, size_t n, fr_t** out) {
    if (ctx->work_cap[slot] < n) {
        if (ctx->work[slot]) cudaFree(ctx->work[slot]);
        ctx->work[slot] = nullptr;
        ctx->work_cap[slot] = 0;
        CU(ctx, cudaMalloc(&ctx->work[slot], std::max<size_t>(n, 1) * sizeof(fr_t)));
        ctx->work_cap[slot] = n;
    }
    *out = ctx->work[slot];
    return ACG_OK;
}

// Relative tile rates of the 1st .. n-th CTA to arrive on an SM (kernels.h CtaRun), measured with the per-CTA timeline
// of the instrumented build on B200 for the 128-row geometries (profiles/r02_cta_timeline_*.txt).
// ACG_K2_WAVE_SHARES="a,b,c,.." overrides them (one relative weight per resident CTA; "0": equal runs); geometries
// without a measurement get equal runs.  Returns the number of weights, 0 for equal runs.
unsigned tile_rate_weights(int variant, unsigned ctas, double* wgt) {
    unsigned n = 0;
    if (const char* env = getenv("ACG_K2_WAVE_SHARES")) {
        const char* p = env;
        while (*p && n < 8) {
            char* end = nullptr;
            const double v = strtod(p, &end);
            if (end == p) break;
            wgt[n++] = v;
            p = (*end == ',') ? end + 1 : end;
        }
        if (n != ctas) return 0;
    } else if ((variant == 0 || variant == 4 || variant == 6) && ctas == 5) {
        static const double dflt[5] = {1.0 / 3.99, 1.0 / 4.12, 1.0 / 4.45, 1.0 / 5.10, 1.0 / 5.95};
        for (n = 0; n < 5; ++n) wgt[n] = dflt[n];
    } else {
        return 0;
    }
    for (unsigned i = 0; i < n; ++i)
        if (!(wgt[i] > 0.0)) return 0;
    return n;
}

// Block placement of the tiled kernel's full grid for this geometry (probed once): smid per block, or empty
int get_placement(acg_ctx* ctx, int variant, const std::vector<uint32_t>** out) {
    *out = &ctx->placement[variant];
    if (ctx->placement_probed[variant]) return ACG_OK;
    ctx->placement_probed[variant] = true;
    const uint32_t ctas = tiled_ctas_per_sm(variant);
    const uint32_t grid = (uint32_t)ctx->sm_count * ctas;
    DevBuf d_smid, d_arrived;
    CU(ctx, d_smid.alloc((size_t)grid * sizeof(uint32_t)));
    CU(ctx, d_arrived.alloc(sizeof(unsigned int)));
    CU(ctx, cudaMemsetAsync(d_arrived.p, 0, sizeof(unsigned int), ctx->stream));
    CU(ctx, cudaMemsetAsync(d_smid.p, 0xFF, (size_t)grid * sizeof(uint32_t), ctx->stream));
    uint32_t launched = 0;
    CU(ctx, launch_probe_placement(variant, ctx->sm_count, d_smid.as<uint32_t>(), d_arrived.as<unsigned int>(), &launched,
                                   ctx->stream));
    ++ctx->launches;
    std::vector<uint32_t> smid(grid);
    CU(ctx, cudaMemcpyAsync(smid.data(), d_smid.p, (size_t)grid * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    // usable only if every SM received exactly `ctas` blocks (nothing else was resident, the grid filled the chip)
    std::map<uint32_t, uint32_t> per_sm;
    for (uint32_t v : smid) ++per_sm[v];
    bool ok = launched == grid && per_sm.size() == (size_t)ctx->sm_count;
    for (const auto& kv : per_sm) ok = ok && kv.second == ctas;
    if (ok) ctx->placement[variant] = std::move(smid);
    return ACG_OK;
}

// (Re)builds the CtaRun table of a system from its plan (run_order, run_cnt, h_meta) and puts it on the device.
int upload_runs(acg_ctx* ctx, acg_r1cs* m) {
    const uint32_t grid = (uint32_t)m->run_cnt.size();
    std::vector<CtaRun> runs(grid);
    uint32_t t = 0;
    for (uint32_t b : m->run_order) {
        CtaRun& run = runs[b];
        run = CtaRun{};
        run.t_begin = t;
        run.t_end = t + m->run_cnt[b];
        if (run.t_begin < run.t_end) run.first = m->h_meta[run.t_begin];
        t = run.t_end;
    }
    if (t != m->n_tiles) return fail(ctx, ACG_ERR_INTERNAL, "upload_runs: the runs do not cover the tiles");
    const uint32_t max_far = kTileGeom[m->variant].max_far;
    std::vector<uint32_t> run_far((size_t)grid * max_far, 0u);
    for (uint32_t b = 0; b < grid; ++b) {
        const TileMeta& f = runs[b].first;
        if (runs[b].t_begin >= runs[b].t_end || f.n_far == 0) continue;
        if (f.n_far > max_far || (size_t)f.far_off + f.n_far > m->h_far_cols.size())
            return fail(ctx, ACG_ERR_INTERNAL, "upload_runs: far columns out of range");
        std::copy(m->h_far_cols.begin() + f.far_off, m->h_far_cols.begin() + f.far_off + f.n_far,
                  run_far.begin() + (size_t)b * max_far);
    }
    if (!m->d_runs) CU(ctx, cudaMalloc(&m->d_runs, runs.size() * sizeof(CtaRun)));
    if (!m->d_run_far) CU(ctx, cudaMalloc(&m->d_run_far, run_far.size() * sizeof(uint32_t)));
    CU(ctx, cudaMemcpyAsync(m->d_runs, runs.data(), runs.size() * sizeof(CtaRun), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(m->d_run_far, run_far.data(), run_far.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));  // (`runs` is about to go out of scope)
    m->n_runs = grid;
    return ACG_OK;
}

struct HostTile {
    uint32_t row0, nrows, e0[3], ne[3], width[3];
};
// Greedy tiling in groups of up to 4 rows: a tile holds at most geom.threads rows and geom.max_gen
// general-coefficient entries, and the sum of its three ELL widths (max row length per matrix) stays within
// geom.max_slots.  Rows longer than kMaxEllWidth in any matrix (Split gates, src/QAP.hs:443-473) are left to
// the row-wise kernel.  gcum[k][r] = number of general entries of matrix k in local rows < r.
void build_tiles(const TileGeometry& geom, const uint32_t* rp[3], const uint32_t* gcum[3],
                 const uint32_t* ccum[3], uint32_t n_local, std::vector<HostTile>& tiles,
                 std::vector<std::pair<uint32_t, uint32_t>>& long_ranges) {
    auto add_long = [&](uint32_t a, uint32_t b) {
        if (!long_ranges.empty() && long_ranges.back().second == a)
            long_ranges.back().second = b;
        else
            long_ranges.emplace_back(a, b);
    };
    uint32_t r = 0;
    while (r < n_local) {
        HostTile t{};
        t.row0 = r;
        uint32_t g_base = 0, c_base = 0;
        for (int k = 0; k < 3; ++k) {
            t.e0[k] = rp[k][r];
            g_base += gcum[k][r];
            c_base += ccum[k][r];
        }
        uint32_t end = r;
        const uint32_t row_cap = geom.threads;
        uint32_t width[3] = {0, 0, 0};
        bool hit_long = false;
        uint32_t step = 4u;
        while (end < n_local && end - r < row_cap) {
            uint32_t g_end = std::min(std::min(end + step, n_local), r + row_cap);
            // a row too long for a tile ends the group before it (and the tile, if it is the group's first row): only
            // that row goes to the warp-per-row kernel, its neighbours stay in tiles
            for (uint32_t q = end; q < g_end; ++q)
                if (rp[0][q + 1] - rp[0][q] > kMaxEllWidth || rp[1][q + 1] - rp[1][q] > kMaxEllWidth ||
                    rp[2][q + 1] - rp[2][q] > kMaxEllWidth) {
                    g_end = q;
                    hit_long = true;
                    break;
                }
            if (g_end == end) break;
            uint32_t nw[3] = {width[0], width[1], width[2]};
            uint64_t gen = 0, cst = 0;
            for (int k = 0; k < 3; ++k) {
                for (uint32_t q = end; q < g_end; ++q) nw[k] = std::max(nw[k], rp[k][q + 1] - rp[k][q]);
                gen += gcum[k][g_end];
                cst += ccum[k][g_end];
            }
            // products needed: general entries off column 0.  One full round of the product phase (one entry
            // per thread) beats one round and a bit: past half a tile of rows, stop at `threads` products.
            const uint64_t prods = (gen - g_base) - (cst - c_base);
            if (nw[0] + nw[1] + nw[2] > geom.max_slots || prods > (uint64_t)geom.max_gen ||
                cst - c_base > (uint64_t)geom.max_const) {
                if (end == r && g_end - end > 1u) {  // not even the first group fits: try its first row alone
                    step = 1u;
                    continue;
                }
                break;
            }
            if (geom.max_gen <= 2u * geom.threads && prods > (uint64_t)geom.threads && end - r >= row_cap / 2u) break;
            for (int k = 0; k < 3; ++k) width[k] = nw[k];
            end = g_end;
        }
        if (end == r) {  // row r cannot open a tile (too long, or alone over a cap): the warp-per-row kernel handles it
            (void)hit_long;
            add_long(r, r + 1u);
            r += 1u;
            continue;
        }
        t.nrows = end - r;
        for (int k = 0; k < 3; ++k) {
            t.ne[k] = rp[k][end] - t.e0[k];
            t.width[k] = width[k];
        }
        tiles.push_back(t);
        r = end;
    }
}

// Host pass over the uploaded slice of a system: structural validation, shard-local row pointers, column words tagged
// with the coefficient class (+1 / -1 / general), running counts of general entries (all / on column 0) per row.
struct HostRows {
    std::vector<uint32_t> local_rp[3], tagged_col[3], gcum[3], ccum[3];
};
inline int host_rows_fail(const char** why, int code, const char* msg) {
    if (why) *why = msg;
    return code;
}
int host_rows(int field, uint32_t n_rows, uint32_t n_cols, const acg_csr* const (&src)[3], uint32_t row_begin,
              uint32_t row_end, HostRows& out, const char** why) {
    const uint32_t n_local = row_end - row_begin;
    // canonical 1 and r-1: the two coefficient values with a multiplication-free fast path
    uint64_t modulus[4], minus_one[4];
    acg_field_constants(field, modulus, nullptr, nullptr, nullptr, nullptr);
    std::memcpy(minus_one, modulus, 32);
    minus_one[0] -= 1;  // r is odd
    // host pass over the uploaded slice: structural validation, coefficient tags, per-row general counts
    std::vector<uint32_t> (&local_rp)[3] = out.local_rp, (&tagged_col)[3] = out.tagged_col, (&gcum)[3] = out.gcum, (&ccum)[3] = out.ccum;
    for (int k = 0; k < 3; ++k) {
        const acg_csr* M = src[k];
        if (!M->rowptr || (M->nnz && (!M->col || !M->val)) || M->nnz > 0xFFFFFFF0ull)
            return host_rows_fail(why, ACG_ERR_BAD_ARG, "null array or nnz too large");
        if (M->rowptr[n_rows] != M->nnz) return host_rows_fail(why, ACG_ERR_BAD_ARG, "rowptr[n_rows] != nnz");
        for (uint32_t r = row_begin; r < row_end; ++r)
            if (M->rowptr[r] > M->rowptr[r + 1])
                return host_rows_fail(why, ACG_ERR_BAD_ARG, "rowptr not monotone");
        const uint32_t e0 = M->rowptr[row_begin], e1 = M->rowptr[row_end];
        if (e1 > M->nnz) return host_rows_fail(why, ACG_ERR_BAD_ARG, "rowptr exceeds nnz");
        local_rp[k].resize((size_t)n_local + 1);
        gcum[k].resize((size_t)n_local + 1);
        ccum[k].resize((size_t)n_local + 1);
        tagged_col[k].resize((size_t)(e1 - e0));
        uint32_t bad = 0, gen = 0, cst = 0;
        for (uint32_t r = 0; r < n_local; ++r) {
            local_rp[k][r] = M->rowptr[row_begin + r] - e0;
            gcum[k][r] = gen;
            ccum[k][r] = cst;
            for (uint32_t e = M->rowptr[row_begin + r]; e < M->rowptr[row_begin + r + 1]; ++e) {
                const uint32_t c = M->col[e];
                bad |= (c >= n_cols);
                const uint64_t* v = M->val + 4ull * e;
                uint32_t tag = kTagGeneral;
                if (v[0] == 1 && (v[1] | v[2] | v[3]) == 0)
                    tag = kTagPlusOne;
                else if (k != 2 && v[0] == minus_one[0] && v[1] == minus_one[1] && v[2] == minus_one[2] &&
                         v[3] == minus_one[3])
                    tag = kTagMinusOne;  // (a -1 in C stays general: C.w is compared, so its terms are never negated)
                gen += (tag == kTagGeneral);
                cst += (tag == kTagGeneral && c == 0u);
                tagged_col[k][e - e0] = (c & kColMask) | (tag << 30);
            }
        }
        local_rp[k][n_local] = e1 - e0;
        gcum[k][n_local] = gen;
        ccum[k][n_local] = cst;
        if (bad) return host_rows_fail(why, ACG_ERR_BAD_ARG, "column index >= n_cols");
    }
    return ACG_OK;
}

// The geometry a system is bound to: the requested one, except that the default gives way to the roomier one for
// systems dense in general coefficients (more than 1.6 products per row).
int pick_variant(int requested, const HostRows& hr, uint32_t n_local) {
    if (requested != 0 || n_local == 0) return requested;
    uint64_t prods = 0;
    for (int k = 0; k < 3; ++k) prods += (uint64_t)hr.gcum[k][n_local] - hr.ccum[k][n_local];
    return prods * 10u > (uint64_t)n_local * 16u ? kDenseTileVariant : 0;
}

// Worker threads of the host-side tile-stream build: ACG_HOST_THREADS, else the hardware concurrency (at most 32).
unsigned upload_threads() {
    if (const char* e = getenv("ACG_HOST_THREADS")) {
        const long v = strtol(e, nullptr, 10);
        if (v >= 1) return (unsigned)std::min<long>(v, 256);
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hc ? hc : 1u, 32u));
}
// f(t, begin, end) over T contiguous chunks of [0, n) on T threads (T = 1: on the caller's).  Nothing escapes a worker;
// if the system refuses a thread, the chunks that did not get one run on the caller's thread afterwards.
template <class F>
int run_chunks(size_t n, size_t T, F&& f) {
    if (T <= 1 || n == 0) return f((size_t)0, (size_t)0, n);
    std::vector<int> rcs(T, ACG_OK);
    std::vector<uint8_t> started(T, 0);
    std::vector<std::thread> workers;
    workers.reserve(T);
    auto guarded = [&](size_t t) {
        try {
            rcs[t] = f(t, n * t / T, n * (t + 1) / T);
        } catch (const std::bad_alloc&) {
            rcs[t] = ACG_ERR_OOM;
        } catch (...) {
            rcs[t] = ACG_ERR_INTERNAL;
        }
    };
    for (size_t t = 0; t < T; ++t) {
        try {
            workers.emplace_back(guarded, t);
            started[t] = 1;
        } catch (...) {
            break;
        }
    }
    for (auto& w : workers) w.join();
    for (size_t t = 0; t < T; ++t)
        if (!started[t]) guarded(t);
    for (int rc : rcs)
        if (rc != ACG_OK) return rc;
    return ACG_OK;
}

// The execution-ready form of a system for the tiled kernel (kernels.h): blobs, tile records, far column lists, where
// the general coefficient values sit (for the Montgomery conversion on the device), and the rows that fit no tile.
struct TileStream {
    std::vector<uint8_t> stream;
    std::vector<uint32_t> gval_offs;  // 16-byte units; bit 31: the high half sits BEFORE the low one
    std::vector<TileMeta> metas;
    std::vector<uint32_t> far_all;
    std::vector<std::pair<uint32_t, uint32_t>> extra_long;  // tiles dropped for their far columns
};
struct FinalTile {
    HostTile t;
    uint32_t win_lo, win_n, far_off, n_far;
};
// Pure host code, no CUDA calls.  The tiles are independent of each other apart from what blob i says about tile i + 1
// (window, far columns, size), so both passes run over contiguous chunks of the tile list on worker threads and the
// pieces are concatenated in order: the result does not depend on the number of threads (acg_tile_stream_digest, tested
// on CPU).  rp / tagged_col: the shard-local CSR structure with tagged column words; val0[k]: value of local entry 0.
int build_tile_stream(const TileGeometry& geom, uint32_t n_cols, const std::vector<HostTile>& tiles,
                      const std::vector<uint32_t> (&rp)[3], const std::vector<uint32_t> (&tagged_col)[3],
                      const uint64_t* const (&val0)[3], unsigned threads, TileStream& out) {
    const uint32_t kProd0 = tile_prod_slot0(geom), kZero = tile_term_slots(geom) - 1u;
    static const bool timing = getenv("ACG_UPLOAD_TIMING") != nullptr;  // (a measurement aid: phase times on stderr)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t_start = now();
    // ---- pass 1: window and far columns of every tile; a tile whose distinct far references exceed the far slots is
    //      split in two (rows stay multiples of 4)
    const size_t T1 = std::min<size_t>(threads, tiles.size() / 64 + 1);
    struct Part1 {
        std::vector<FinalTile> tiles;
        std::vector<uint32_t> far;
        std::vector<std::pair<uint32_t, uint32_t>> dropped;
    };
    std::vector<Part1> p1(T1);
    int rc = run_chunks(tiles.size(), T1, [&](size_t t, size_t begin, size_t end) -> int {
        Part1& part = p1[t];
        std::vector<uint32_t> ref_cols, far_cols;
        std::vector<HostTile> work(std::make_reverse_iterator(tiles.begin() + (ptrdiff_t)end),
                                   std::make_reverse_iterator(tiles.begin() + (ptrdiff_t)begin));  // pop from the back
        while (!work.empty()) {
            HostTile tl = work.back();
            work.pop_back();
            // witness window: the contiguous slice of geom.window elements that covers most references of the tile
            ref_cols.clear();
            for (int k = 0; k < 3; ++k)
                for (uint32_t e = 0; e < tl.ne[k]; ++e) ref_cols.push_back(tagged_col[k][tl.e0[k] + e] & kColMask);
            std::sort(ref_cols.begin(), ref_cols.end());
            const uint32_t win_n = std::min<uint32_t>(geom.window, n_cols);
            uint32_t win_lo = 0;
            {
                size_t best = 0, lo_i = 0;
                for (size_t hi_i = 0; hi_i < ref_cols.size(); ++hi_i) {
                    while (ref_cols[hi_i] - ref_cols[lo_i] >= win_n) ++lo_i;
                    if (hi_i - lo_i + 1 > best) {
                        best = hi_i - lo_i + 1;
                        win_lo = ref_cols[lo_i];
                    }
                }
                if (win_lo + win_n > n_cols) win_lo = n_cols - win_n;
            }
            far_cols.clear();
            for (uint32_t c : ref_cols)
                if ((c < win_lo || c - win_lo >= win_n) && (far_cols.empty() || far_cols.back() != c)) far_cols.push_back(c);
            if (far_cols.size() > geom.max_far && tl.nrows > 4) {
                const uint32_t half = ((tl.nrows / 2 + 3) / 4) * 4;
                HostTile lo_t{}, hi_t{};
                lo_t.row0 = tl.row0;
                lo_t.nrows = half;
                hi_t.row0 = tl.row0 + half;
                hi_t.nrows = tl.nrows - half;
                for (HostTile* q : {&lo_t, &hi_t})
                    for (int k = 0; k < 3; ++k) {
                        q->e0[k] = rp[k][q->row0];
                        q->ne[k] = rp[k][q->row0 + q->nrows] - q->e0[k];
                        q->width[k] = 0;
                        for (uint32_t r = q->row0; r < q->row0 + q->nrows; ++r)
                            q->width[k] = std::max(q->width[k], rp[k][r + 1] - rp[k][r]);
                    }
                work.push_back(hi_t);
                work.push_back(lo_t);
                continue;
            }
            if (far_cols.size() > geom.max_far) {  // 4 rows with more distinct far columns than slots: row-wise kernel
                part.dropped.emplace_back(tl.row0, tl.row0 + tl.nrows);
                continue;
            }
            FinalTile ft{};
            ft.t = tl;
            ft.win_lo = win_lo;
            ft.win_n = win_n;
            ft.far_off = (uint32_t)part.far.size();  // (relative to the part; rebased below)
            ft.n_far = (uint32_t)far_cols.size();
            part.far.insert(part.far.end(), far_cols.begin(), far_cols.end());
            part.tiles.push_back(ft);
        }
        return ACG_OK;
    });
    if (rc) return rc;
    std::vector<FinalTile> final_tiles;
    {
        size_t n_final = 0, n_far = 0;
        for (const Part1& part : p1) {
            n_final += part.tiles.size();
            n_far += part.far.size();
        }
        if (n_far > 0xFFFFFFF0ull) return ACG_ERR_UNSUPPORTED;
        final_tiles.reserve(n_final);
        out.far_all.reserve(n_far);
        for (Part1& part : p1) {
            const uint32_t base = (uint32_t)out.far_all.size();
            for (FinalTile ft : part.tiles) {
                ft.far_off += base;
                final_tiles.push_back(ft);
            }
            out.far_all.insert(out.far_all.end(), part.far.begin(), part.far.end());
            out.extra_long.insert(out.extra_long.end(), part.dropped.begin(), part.dropped.end());
            std::vector<FinalTile>().swap(part.tiles);
            std::vector<uint32_t>().swap(part.far);
        }
    }
    const std::vector<uint32_t>& far_all = out.far_all;
    const auto t_pass1 = now();
    // ---- pass 2: emit the blobs.  Blob i also carries what the kernel needs to start tile i + 1 while blob i is
    //      still the only one in shared memory: its far witness columns and, in the header, its size and window.
    const size_t T2 = std::min<size_t>(threads, final_tiles.size() / 64 + 1);
    struct Part2 {
        std::vector<uint8_t> stream;
        std::vector<uint32_t> gval_offs;  // relative to the part's stream
        std::vector<TileMeta> metas;      // blob_off16 relative to the part's stream
    };
    std::vector<Part2> p2(T2);
    rc = run_chunks(final_tiles.size(), T2, [&](size_t t, size_t begin, size_t end) -> int {
        Part2& part = p2[t];
        std::vector<uint8_t>& stream = part.stream;
        stream.reserve((end - begin) * (size_t)(tile_blob_capacity(geom) / 2u + 256u));
        auto align16 = [&]() { stream.resize((stream.size() + 15) & ~(size_t)15, 0); };
        auto put16 = [&](uint16_t v) {
            stream.push_back((uint8_t)(v & 0xFF));
            stream.push_back((uint8_t)(v >> 8));
        };
        auto put32 = [&](uint32_t v) {
            for (int i = 0; i < 4; ++i) stream.push_back((uint8_t)(v >> (8 * i)));
        };
        std::vector<int32_t> gid[3];
        std::vector<uint32_t> order;
        for (size_t ti = begin; ti < end; ++ti) {
            const FinalTile& ft = final_tiles[ti];
            const HostTile& tl = ft.t;
            const uint32_t win_lo = ft.win_lo, win_n = ft.win_n;
            const uint32_t* far_b = far_all.data() + ft.far_off;
            const uint32_t* far_e = far_b + ft.n_far;
            // chunk (16-byte unit from the start of shared memory) of the low half of a term slot / of the 32-byte
            // value j of a blob section; the window is written by a linear bulk copy and stays in natural order
            const uint32_t term_base16 = tile_terms_offset(geom) / 16u;
            auto term_chunk = [&](uint32_t slot, bool in_window) -> uint32_t {
                const uint32_t c = term_base16 + 2u * slot;
                return (geom.swizzle && !in_window) ? swz16(c) : c;
            };
            auto blob_chunk = [&](uint32_t off, uint32_t j) -> uint32_t {
                const uint32_t c = off / 16u + 2u * j;
                return geom.swizzle ? swz16(c) : c;
            };
            auto chunk_of = [&](uint32_t c) -> uint32_t {  // witness column -> chunk of its term
                if (c >= win_lo && c - win_lo < win_n) return term_chunk(c - win_lo, true);
                return term_chunk(tile_far_slot0(geom, (uint32_t)ti) + (uint32_t)(std::lower_bound(far_b, far_e, c) - far_b),
                                  false);
            };
            align16();
            const size_t base = stream.size();
            TileMeta tm{};
            tm.blob_off16 = (uint32_t)(base / 16);
            tm.win_lo = win_lo;
            tm.win_n = win_n;
            tm.far_off = ft.far_off;
            tm.n_far = ft.n_far;
            stream.resize(base + sizeof(TileHeader), 0);
            TileHeader h{};
            h.row0 = tl.row0;
            h.nrows = tl.nrows;
            for (int k = 0; k < 3; ++k) h.width[k] = tl.width[k];
            if (ti + 1 < final_tiles.size()) {
                const FinalTile& nx = final_tiles[ti + 1];
                h.next_win_lo = nx.win_lo;
                h.next_win_n = nx.win_n;
                h.next_n_far = nx.n_far;
            }
            // general entries, numbered A rows, then B rows, then C rows, entry order -- separately for the ones that
            // need a product (gid >= 0: product index) and the ones on column 0 (gid < 0: ~index among those)
            uint32_t n_prod = 0, n_const = 0;
            for (int k = 0; k < 3; ++k) {
                gid[k].assign(tl.ne[k], 0);
                for (uint32_t e = 0; e < tl.ne[k]; ++e) {
                    const uint32_t word = tagged_col[k][tl.e0[k] + e];
                    if ((word >> 30) != kTagGeneral) continue;
                    gid[k][e] = (word & kColMask) == 0u ? ~(int32_t)(n_const++) : (int32_t)(n_prod++);
                }
            }
            h.n_general = n_prod;
            h.n_const = n_const;
            // rows sorted by shape (lengths of their A, B, C rows: rows of one shape end up in the same warps), and the
            // ELL widths per warp of that order (kernels.h TileWarp)
            order.resize(tl.nrows);
            for (uint32_t r = 0; r < tl.nrows; ++r) order[r] = r;
            auto row_len = [&](int k, uint32_t r) { return rp[k][tl.row0 + r + 1] - rp[k][tl.row0 + r]; };
            std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
                const uint32_t kx = (row_len(1, x) << 16) | (row_len(0, x) << 8) | row_len(2, x);
                const uint32_t ky = (row_len(1, y) << 16) | (row_len(0, y) << 8) | row_len(2, y);
                return kx < ky;
            });
            TileWarp warps[kMaxTileWarps] = {};
            const uint32_t n_warps = (tl.nrows + 31u) / 32u;
            uint32_t n_words = 0;
            for (uint32_t q = 0; q < n_warps; ++q) {
                TileWarp& tw = warps[q];
                tw.words0 = (uint16_t)n_words;
                tw.nrows = (uint8_t)std::min(32u, tl.nrows - 32u * q);
                for (uint32_t l = 0; l < tw.nrows; ++l)
                    for (int k = 0; k < 3; ++k)
                        tw.width[k] = std::max<uint8_t>(tw.width[k], (uint8_t)row_len(k, order[32u * q + l]));
                n_words += (uint32_t)(tw.width[0] + tw.width[1] + tw.width[2]) * tw.nrows;
            }
            // layout (offsets from the blob start; the blob lands at shared-memory offset 0)
            auto up = [](uint32_t x, uint32_t a) { return (x + a - 1) / a * a; };
            h.off_words = up(kTilePermOffset + tl.nrows, 16);
            h.off_next_far = up(h.off_words + n_words * 4u, 16);
            h.off_gop = up(h.off_next_far + h.next_n_far * 4u, 16);
            h.off_gval = up(h.off_gop + n_prod * 2u, 32);
            // warp records, row offsets, then the entry words: per warp, slot-major over the warp's rows
            stream.resize(base + kTileWarpsOffset, 0);
            for (uint32_t q = 0; q < kMaxTileWarps; ++q) {
                put16(warps[q].words0);
                stream.push_back(warps[q].nrows);
                for (int k = 0; k < 3; ++k) stream.push_back(warps[q].width[k]);
                put16(0);
            }
            for (uint32_t r = 0; r < tl.nrows; ++r) stream.push_back((uint8_t)order[r]);
            stream.resize(base + h.off_words, 0);
            for (uint32_t q = 0; q < n_warps; ++q)
                for (int k = 0; k < 3; ++k)
                    for (uint32_t j = 0; j < warps[q].width[k]; ++j)
                        for (uint32_t l = 0; l < warps[q].nrows; ++l) {
                            const uint32_t r = order[32u * q + l];
                            const uint32_t s0 = rp[k][tl.row0 + r], s1 = rp[k][tl.row0 + r + 1];
                            if (s0 + j >= s1) {
                                put32(term_chunk(kZero, false));
                                continue;
                            }
                            const uint32_t word = tagged_col[k][s0 + j];
                            const uint32_t tag = word >> 30;
                            if (tag == kTagGeneral) {
                                const int32_t g = gid[k][s0 + j - tl.e0[k]];
                                put32(g >= 0 ? (geom.prod_in_place ? blob_chunk(h.off_gval, (uint32_t)g)
                                                                   : term_chunk(kProd0 + (uint32_t)g, false))
                                             : blob_chunk(h.off_gval, n_prod + (uint32_t)(~g)));
                            } else {
                                put32((tag == kTagMinusOne ? kTermSign : 0u) | chunk_of(word & kColMask));
                            }
                        }
            stream.resize(base + h.off_next_far, 0);
            for (uint32_t f = 0; f < h.next_n_far; ++f) put32(far_all[final_tiles[ti + 1].far_off + f]);
            stream.resize(base + h.off_gop, 0);
            for (int k = 0; k < 3; ++k)
                for (uint32_t e = 0; e < tl.ne[k]; ++e) {
                    const uint32_t word = tagged_col[k][tl.e0[k] + e];
                    if ((word >> 30) == kTagGeneral && gid[k][e] >= 0) put16((uint16_t)chunk_of(word & kColMask));
                }
            stream.resize(base + h.off_gval + (size_t)(n_prod + n_const) * 32u, 0);
            for (int k = 0; k < 3; ++k)
                for (uint32_t e = 0; e < tl.ne[k]; ++e)
                    if ((tagged_col[k][tl.e0[k] + e] >> 30) == kTagGeneral) {
                        const int32_t g = gid[k][e];
                        const uint32_t lo = blob_chunk(h.off_gval, g >= 0 ? (uint32_t)g : n_prod + (uint32_t)(~g));
                        const uint64_t* v = val0[k] + 4ull * (tl.e0[k] + e);
                        std::memcpy(stream.data() + base + (size_t)lo * 16u, v, 16);
                        std::memcpy(stream.data() + base + (size_t)(lo ^ 1u) * 16u, v + 2, 16);
                        // (the conversion kernel finds the high half before / after the low one: bit 31)
                        part.gval_offs.push_back((uint32_t)(base / 16 + lo) | ((lo & 1u) ? 0x80000000u : 0u));
                    }
            align16();
            h.bytes = (uint32_t)(stream.size() - base);
            std::memcpy(stream.data() + base, &h, sizeof h);
            tm.blob_bytes = h.bytes;
            part.metas.push_back(tm);
        }
        align16();
        return ACG_OK;
    });
    if (rc) return rc;
    const auto t_pass2 = now();
    {   // concatenate the parts (every part is a multiple of 16 bytes) and rebase what points into the stream
        size_t bytes = 0, n_g = 0;
        for (const Part2& part : p2) {
            bytes += part.stream.size();
            n_g += part.gval_offs.size();
        }
        if (bytes / 16 > 0x7FFFFFF0ull) return ACG_ERR_UNSUPPORTED;
        out.stream.resize(bytes);
        out.gval_offs.reserve(n_g);
        out.metas.reserve(final_tiles.size());
        size_t at = 0;
        for (Part2& part : p2) {
            if (!part.stream.empty()) std::memcpy(out.stream.data() + at, part.stream.data(), part.stream.size());
            const uint32_t base16 = (uint32_t)(at / 16);
            for (uint32_t g : part.gval_offs) out.gval_offs.push_back(((g & 0x7FFFFFFFu) + base16) | (g & 0x80000000u));
            for (TileMeta tm : part.metas) {
                tm.blob_off16 += base16;
                out.metas.push_back(tm);
            }
            at += part.stream.size();
            std::vector<uint8_t>().swap(part.stream);
        }
    }
    for (size_t ti = 0; ti + 1 < out.metas.size(); ++ti) {  // next_bytes: known once the next blob is laid out
        const uint32_t nb = out.metas[ti + 1].blob_bytes;
        std::memcpy(out.stream.data() + (size_t)out.metas[ti].blob_off16 * 16u + offsetof(TileHeader, next_bytes), &nb, sizeof nb);
    }
    if (timing)
        fprintf(stderr, "[tile stream] %zu tiles, %u threads: windows / far columns %.1f ms, blobs %.1f ms, concatenation %.1f ms\n",
                out.metas.size(), threads, ms(t_start, t_pass1), ms(t_pass1, t_pass2), ms(t_pass2, now()));
    return ACG_OK;
}

// Structural check of a tile stream against everything the tiled kernel assumes about it (host-only; run by
// acg_tile_stream_digest, i.e. by the CPU test suite): sizes within the geometry's capacities, sections inside the blob
// and in order, every entry / operand word a chunk inside the CTA's shared memory (and inside the blob, the witness
// window, the far buffer of the tile's parity, the product slots or the zero slot), warp records consistent with the
// row count, the row offsets a permutation, the hand-over fields of blob i equal to tile i + 1's record, far columns
// sorted, distinct and outside the window, and tiles plus long ranges covering every row exactly once.
// Returns nullptr or what is wrong.
const char* validate_tile_stream(const TileGeometry& geom, uint32_t n_local, uint32_t n_cols, const TileStream& ts,
                                 const std::vector<std::pair<uint32_t, uint32_t>>& long_ranges) {
    const uint32_t smem_chunks = tile_smem_bytes(geom) / 16u, terms16 = tile_terms_offset(geom) / 16u;
    const uint32_t kZero = tile_term_slots(geom) - 1u, kProd0 = tile_prod_slot0(geom);
    std::vector<uint8_t> covered(n_local, 0);
    for (const auto& lr : long_ranges) {
        if (lr.first > lr.second || lr.second > n_local) return "long range out of bounds";
        for (uint32_t r = lr.first; r < lr.second; ++r) {
            if (covered[r]) return "row covered twice";
            covered[r] = 1;
        }
    }
    auto unswz = [&](uint32_t chunk) { return geom.swizzle ? swz16(chunk) : chunk; };  // (swz16 is an involution)
    for (size_t ti = 0; ti < ts.metas.size(); ++ti) {
        const TileMeta& tm = ts.metas[ti];
        if ((size_t)tm.blob_off16 * 16u + tm.blob_bytes > ts.stream.size()) return "blob outside the stream";
        if (tm.blob_bytes % 16u || tm.blob_bytes > tile_blob_capacity(geom)) return "blob size";
        if (tm.win_n > geom.window || (uint64_t)tm.win_lo + tm.win_n > n_cols) return "window";
        if (tm.n_far > geom.max_far || (size_t)tm.far_off + tm.n_far > ts.far_all.size()) return "far columns";
        const uint32_t* far = ts.far_all.data() + tm.far_off;
        for (uint32_t f = 0; f < tm.n_far; ++f) {
            if (far[f] >= n_cols) return "far column out of range";
            if (f && far[f] <= far[f - 1]) return "far columns not strictly ascending";
            if (far[f] >= tm.win_lo && far[f] - tm.win_lo < tm.win_n) return "far column inside the window";
        }
        const uint8_t* blob = ts.stream.data() + (size_t)tm.blob_off16 * 16u;
        TileHeader h;
        std::memcpy(&h, blob, sizeof h);
        if (h.bytes != tm.blob_bytes) return "header size != record size";
        if (h.nrows == 0 || h.nrows > geom.threads || (uint64_t)h.row0 + h.nrows > n_local) return "rows of a tile";
        if (h.n_general > geom.max_gen || h.n_const > geom.max_const) return "general entries beyond the caps";
        if (h.width[0] + h.width[1] + h.width[2] > geom.max_slots) return "ELL widths beyond the slots";
        for (uint32_t r = h.row0; r < h.row0 + h.nrows; ++r) {
            if (covered[r]) return "row covered twice";
            covered[r] = 1;
        }
        if (!(kTilePermOffset + h.nrows <= h.off_words && h.off_words <= h.off_next_far && h.off_next_far <= h.off_gop &&
              h.off_gop <= h.off_gval && h.off_gval % 32u == 0 &&
              h.off_gval + (uint64_t)(h.n_general + h.n_const) * 32u <= h.bytes))
            return "blob sections out of order";
        if (ti + 1 < ts.metas.size()) {
            const TileMeta& nx = ts.metas[ti + 1];
            if (h.next_bytes != nx.blob_bytes || h.next_win_lo != nx.win_lo || h.next_win_n != nx.win_n ||
                h.next_n_far != nx.n_far || nx.blob_off16 != tm.blob_off16 + tm.blob_bytes / 16u)
                return "hand-over fields != the next tile's record";
            if (h.off_next_far + (uint64_t)h.next_n_far * 4u > h.off_gop) return "next far columns overflow their section";
            if (h.next_n_far && std::memcmp(blob + h.off_next_far, ts.far_all.data() + nx.far_off, (size_t)nx.n_far * 4u))
                return "next far columns != the next tile's";
        } else if (h.next_bytes || h.next_n_far) {
            return "last tile hands over to nothing";
        }
        // warp records and the row permutation
        const uint32_t n_warps = (h.nrows + 31u) / 32u;
        uint32_t rows_seen = 0, n_words = 0;
        for (uint32_t q = 0; q < kMaxTileWarps; ++q) {
            TileWarp tw;
            std::memcpy(&tw, blob + kTileWarpsOffset + q * sizeof(TileWarp), sizeof tw);
            if (q >= n_warps) {
                if (tw.nrows) return "warp record beyond the tile's warps";
                continue;
            }
            if (tw.nrows == 0 || tw.nrows > 32u || tw.words0 != n_words) return "warp record";
            if (tw.width[0] > h.width[0] || tw.width[1] > h.width[1] || tw.width[2] > h.width[2]) return "warp wider than its tile";
            rows_seen += tw.nrows;
            n_words += (uint32_t)(tw.width[0] + tw.width[1] + tw.width[2]) * tw.nrows;
        }
        if (rows_seen != h.nrows || h.off_words + (uint64_t)n_words * 4u > h.off_next_far) return "warp records != rows / words";
        std::vector<uint8_t> seen(h.nrows, 0);
        for (uint32_t r = 0; r < h.nrows; ++r) {
            const uint8_t o = blob[kTilePermOffset + r];
            if (o >= h.nrows || seen[o]) return "row offsets are not a permutation";
            seen[o] = 1;
        }
        // where a term may live: window, this tile's far buffer, a product slot / an in-place value, the zero slot
        const uint32_t far0 = tile_far_slot0(geom, (uint32_t)ti);
        auto term_ok = [&](uint32_t chunk, bool allow_values) {
            if (chunk >= smem_chunks) return false;
            if (chunk >= terms16) {  // the term array: the window is linear, the rest swizzled
                const uint32_t lin = chunk - terms16;
                if (lin < 2u * geom.window) return (lin & 1u) == 0u && lin / 2u < tm.win_n;
                const uint32_t c = unswz(chunk);
                if (c < terms16 || ((c - terms16) & 1u)) return false;
                const uint32_t slot = (c - terms16) / 2u;
                if (slot >= far0 && slot < far0 + tm.n_far) return true;
                if (allow_values && !geom.prod_in_place && slot >= kProd0 && slot < kProd0 + h.n_general) return true;
                return allow_values && slot == kZero;
            }
            if (!allow_values) return false;  // inside the blob: a value of the general section
            const uint32_t c = unswz(chunk);
            if ((uint64_t)c * 16u < h.off_gval || ((c * 16u - h.off_gval) % 32u)) return false;
            const uint32_t j = (c * 16u - h.off_gval) / 32u;
            return geom.prod_in_place ? j < h.n_general + h.n_const : (j >= h.n_general && j < h.n_general + h.n_const);
        };
        for (uint32_t k = 0; k < n_words; ++k) {
            uint32_t word;
            std::memcpy(&word, blob + h.off_words + 4u * k, 4);
            if (!term_ok(word & ~kTermSign, true)) return "entry word outside its legal places";
        }
        if (h.off_gop + (uint64_t)h.n_general * 2u > h.off_gval) return "operand words overflow their section";
        for (uint32_t j = 0; j < h.n_general; ++j) {
            uint16_t op;
            std::memcpy(&op, blob + h.off_gop + 2u * j, 2);
            if (!term_ok(op, false)) return "operand word is not a witness term";
        }
    }
    for (uint32_t r = 0; r < n_local; ++r)
        if (!covered[r]) return "row not covered";
    // value offsets: one per general entry, inside the stream
    uint64_t n_values = 0;
    for (const TileMeta& tm : ts.metas) {
        TileHeader h;
        std::memcpy(&h, ts.stream.data() + (size_t)tm.blob_off16 * 16u, sizeof h);
        n_values += h.n_general + h.n_const;
    }
    if (n_values != ts.gval_offs.size()) return "value offsets != general entries";
    for (uint32_t g : ts.gval_offs)
        if (((size_t)(g & 0x7FFFFFFFu) + 1u) * 16u > ts.stream.size()) return "value offset outside the stream";
    return nullptr;
}

// Enqueues one check of the shard: the kernels accumulate into the context's scratch pair and the LAST launch
// finalises into d_result (kernels.h CheckEpilogue) -- including, when `peer` is given, the all-reduce over peer
// memory.  No initialisation launch; one kernel for a system without over-long rows.
int enqueue_check(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* wv, unsigned long long* d_result, fr_t* Aw, fr_t* Bw,
                  fr_t* Cw, cudaStream_t s, uint32_t* launches, acg_peer* peer = nullptr) {
    const uint32_t n_local = m->row_end - m->row_begin;
    const fr_t* w = wv->d;
    {   // an asynchronous update of the witness (copy stream) comes first
        int rc = vec_acquire(ctx, wv, s);
        if (rc) return rc;
    }
    CheckEpilogue acc{ctx->d_accum, ctx->d_ticket, nullptr, PeerSlots{}, 0ull, 0u};  // accumulate only
    CheckEpilogue fin = acc;                                                     // .. and finalise
    fin.out = d_result;
    if (peer) {
        fin.peers = peer->slots;
        fin.seq = ++peer->seq;
    }
    const int which = ctx->check_kernel == ACG_CHECK_AUTO ? ACG_CHECK_TILED : ctx->check_kernel;
    const bool prof = 2 * (ctx->prof_used + 1) <= ctx->prof_ev.size();
    bool finalised = false;
    if (which == ACG_CHECK_ROWWISE) {
        if (prof) CU(ctx, cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used], s));
        if (n_local) {
            CU(ctx, launch_r1cs_rowwise(ctx->field, m->dev, w, 0, n_local, m->row_begin + m->row_offset, fin, Aw, Bw, Cw, s));
            ++*launches;
            finalised = true;
        }
        if (prof) {
            CU(ctx, cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used + 1], s));
            ++ctx->prof_used;
        }
    } else {
        // rows too long for a tile first -- all of them in one warp-per-row launch (accumulate only) --, the tiled
        // kernel last: it finalises the check.  At most two launches, however many Split gates the circuit has.
        // .. unless the tiled kernel of this geometry checks them itself after its tiles: then one launch
        static const bool separate_long = getenv("ACG_K2_SEPARATE_LONGROWS") != nullptr;  // (a measurement aid)
        const bool fused_long = m->n_long_rows && m->n_tiles && m->d_long_desc && !separate_long &&
                                tiled_checks_long_rows(m->variant);
        if (m->n_long_rows && !fused_long) {
            const bool last = m->n_tiles == 0;
            CU(ctx, launch_r1cs_longrows(ctx->field, m->dev, w, m->d_long_rows, m->n_long_rows, m->row_begin + m->row_offset,
                                         last ? fin : acc, Aw, Bw, Cw, s));
            ++*launches;
            finalised = finalised || last;
        }
        if (prof) CU(ctx, cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used], s));
        if (m->n_tiles) {
            DevTileStream ts{};
            ts.blobs = m->d_stream;
            ts.meta = m->d_meta;
            ts.far_cols = m->d_far_cols;
            ts.n_tiles = m->n_tiles;
            ts.variant = (uint32_t)m->variant;
            ts.blobs_len16 = (uint32_t)(m->blob_bytes / 16);
            ts.n_cols = m->n_cols;
            ts.runs = m->d_runs;
            ts.n_runs = m->n_runs;
            ts.run_far = m->d_run_far;
            // back-to-back checks of the same system may overlap (see CheckEpilogue): every entry point bumps ctx->ops,
            // so "the previous operation was that check" is ops == last_check_op + 1 -- nothing, in particular no
            // witness update, was enqueued through this context in between (the witnesses of the two checks may be
            // different resident vectors)
            const bool emit = Aw || Bw || Cw;
            const bool single = m->long_ranges.empty();
            const bool chain = single && !emit && !prof && ctx->overlap_checks && ctx->last_m == m &&
                               ctx->last_stream == s && ctx->ops == ctx->last_check_op + 1;
            if (chain) fin.overlap = 1u;
            if (!chain && !peer && ctx->direct_handover) {  // one GPU, no overlap: nothing to do at the end of a clean check
                fin.gate_seq = ++ctx->gate_seq;
                fin.gate = ctx->d_gate + fin.gate_seq % kGateRing;
            }
            CU(ctx, launch_r1cs_tiled(ctx->field, ts, w, m->row_begin + m->row_offset, fin, Aw, Bw, Cw, ctx->sm_count, s,
                                      fused_long ? m->d_long_desc : nullptr, m->n_long_rows, &m->long_claims));
            ++*launches;
            finalised = true;
            if (single && !emit && !prof) {
                ctx->last_m = m;
                ctx->last_stream = s;
                ctx->last_check_op = ctx->ops;
            }
        }
        if (prof) {
            CU(ctx, cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used + 1], s));
            ++ctx->prof_used;
        }
    }
    if (!finalised) {  // an empty shard: nothing ran, the result is {0, none} (and the peers still expect this rank)
        CU(ctx, launch_init_result(d_result, s));
        ++*launches;
        if (peer) {
            CU(ctx, launch_peer_allreduce(fin.peers, fin.seq, d_result, s));
            ++*launches;
        }
    }
    return vec_release(ctx, wv, s);
}

float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

}  // namespace

// --------------------------------------------------------------------------------------------------
// library / context
// --------------------------------------------------------------------------------------------------
extern "C" {

int acg_abi_version(void) { return ACG_ABI_VERSION; }

const char* acg_strerror(int code) {
    switch (code) {
        case ACG_OK: return "ok";
        case ACG_ERR_BAD_ARG: return "bad argument";
        case ACG_ERR_NON_CANONICAL: return "non-canonical field element (>= modulus)";
        case ACG_ERR_CUDA: return "CUDA error";
        case ACG_ERR_OOM: return "out of device memory";
        case ACG_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case ACG_ERR_UNSUPPORTED: return "unsupported size or configuration";
        case ACG_ERR_INTERNAL: return "internal error";
        default: return "unknown error code";
    }
}

const char* acg_last_error(const acg_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

int acg_ctx_create(int field_id, int device, acg_ctx** out) {
    ACG_TRY
    if (!out || (field_id != ACG_FIELD_BN254_FR && field_id != ACG_FIELD_BLS12_381_FR)) return ACG_ERR_BAD_ARG;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev <= 0) return ACG_ERR_NO_DEVICE;
    if (device < 0 || device >= n_dev) return ACG_ERR_BAD_ARG;
    acg_ctx* ctx = new (std::nothrow) acg_ctx();
    if (!ctx) return ACG_ERR_OOM;
    ctx->field = field_id;
    ctx->device = device;
    auto bail = [&](cudaError_t ce, const char* what) {
        int rc = fail_cuda(nullptr, ce, what);
        acg_ctx_destroy(ctx);
        return rc;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) {
        acg_ctx_destroy(ctx);
        return ACG_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    }
    ctx->sm_count = prop.multiProcessorCount;
    {   // keep stream-ordered scratch (one-shot path) cached in the pool instead of returning it at every sync
        cudaMemPool_t mp;
        if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMalloc(&ctx->d_result, 2 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc(&ctx->d_accum, (4 + kGateRing) * sizeof(unsigned long long))) != cudaSuccess)
        return bail(e, "cudaMalloc");
    ctx->d_ticket = reinterpret_cast<unsigned int*>(ctx->d_accum + 2);
    ctx->d_gate = ctx->d_accum + 4;
    {
        if ((e = cudaMemset(ctx->d_accum, 0, (4 + kGateRing) * sizeof(unsigned long long))) != cudaSuccess)
            return bail(e, "cudaMemset");
        const unsigned long long init[4] = {0ull, ~0ull, 0ull, 0ull};
        if ((e = cudaMemcpy(ctx->d_accum, init, sizeof init, cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(e, "cudaMemcpy");
    }
    if (const char* t = getenv("ACG_K2_TICKET")) ctx->direct_handover = atoi(t) == 0;
    if ((e = cudaMalloc(&ctx->d_flag, sizeof(int))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMallocHost(&ctx->h_result, 2 * sizeof(unsigned long long))) != cudaSuccess)
        return bail(e, "cudaMallocHost");
    if ((e = cudaMallocHost(&ctx->h_flag, sizeof(int))) != cudaSuccess) return bail(e, "cudaMallocHost");
    *out = ctx;
    return ACG_OK;
    ACG_CATCH_NOCTX()
}

void acg_ctx_destroy(acg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto& kv : ctx->plans) ntt_plan_destroy(kv.second);
    for (auto& kv : ctx->coset) {
        cudaFree(kv.second.hi);
        cudaFree(kv.second.lo);
        cudaFree(kv.second.ihi);
        cudaFree(kv.second.ilo);
    }
    if (ctx->d_result) cudaFree(ctx->d_result);
    if (ctx->d_accum) cudaFree(ctx->d_accum);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto e : ctx->prof_ev) cudaEventDestroy(e);
    if (ctx->staging) cudaFree(ctx->staging);
    for (auto& p : ctx->work)
        if (p) cudaFree(p);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int acg_ctx_set_check_kernel(acg_ctx* ctx, int which) {
    ACG_TRY
    if (!ctx || which < ACG_CHECK_AUTO || which > ACG_CHECK_TILED) return ACG_ERR_BAD_ARG;
    ctx->check_kernel = which;
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_ctx_set_overlap_checks(acg_ctx* ctx, int on) {
    ACG_TRY
    if (!ctx) return ACG_ERR_BAD_ARG;
    ctx->overlap_checks = on ? 1 : 0;
    return ACG_OK;
    ACG_CATCH(ctx)
}
int acg_ctx_set_tiled_variant(acg_ctx* ctx, int variant) {
    ACG_TRY
    if (!ctx || variant < 0 || variant >= kNumTileVariants) return ACG_ERR_BAD_ARG;
    ctx->tiled_variant = variant;
    return ACG_OK;
    ACG_CATCH(ctx)
}


int acg_last_timing(const acg_ctx* ctx, acg_timing* out) {
    ACG_TRY
    if (!ctx || !out) return ACG_ERR_BAD_ARG;
    *out = ctx->timing;
    return ACG_OK;
    ACG_CATCH(ctx)
}

uint64_t acg_kernel_launch_count(const acg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int acg_profile_begin(acg_ctx* ctx, uint32_t max_launches) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    while (ctx->prof_ev.size() < 2ull * max_launches) {
        cudaEvent_t e;
        CU(ctx, cudaEventCreate(&e));
        ctx->prof_ev.push_back(e);
    }
    while (ctx->prof_ev.size() > 2ull * max_launches) {
        cudaEventDestroy(ctx->prof_ev.back());
        ctx->prof_ev.pop_back();
    }
    ctx->prof_used = 0;
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_profile_end(acg_ctx* ctx, float* ms_out, uint32_t capacity, uint32_t* n_out) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    uint32_t n = ctx->prof_used < capacity ? ctx->prof_used : capacity;
    for (uint32_t i = 0; i < n; ++i) {
        CU(ctx, cudaEventSynchronize(ctx->prof_ev[2 * i + 1]));
        CU(ctx, cudaEventElapsedTime(&ms_out[i], ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
    }
    if (n_out) *n_out = n;
    for (auto e : ctx->prof_ev) cudaEventDestroy(e);
    ctx->prof_ev.clear();
    ctx->prof_used = 0;
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_field_constants(int field_id, uint64_t modulus[4], uint64_t mont_r[4], uint64_t mont_r2[4], uint64_t* ninv64,
                        uint32_t* two_adic) {
    ACG_TRY
    return with_field(field_id, [&](auto p) {
        using P = decltype(p);
        if (modulus) limbs_from_fr(modulus, host_const<P>(&P::p));
        if (mont_r) limbs_from_fr(mont_r, host_const<P>(&P::one));
        if (mont_r2) limbs_from_fr(mont_r2, host_const<P>(&P::r2));
        if (ninv64) *ninv64 = P::NINV64;
        if (two_adic) *two_adic = (uint32_t)P::TWO_ADICITY;
        return (int)ACG_OK;
    });
    ACG_CATCH_NOCTX()
}

int acg_root_of_unity(int field_id, uint32_t k, uint64_t out[4]) {
    ACG_TRY
    if (!out) return ACG_ERR_BAD_ARG;
    return with_field(field_id, [&](auto p) {
        using P = decltype(p);
        if (k > (uint32_t)P::TWO_ADICITY) return (int)ACG_ERR_UNSUPPORTED;
        limbs_from_fr(out, fr_from_mont<P>(host_root_of_unity<P>(k)));
        return (int)ACG_OK;
    });
    ACG_CATCH_NOCTX()
}

// --------------------------------------------------------------------------------------------------
// R1CS upload / check
// --------------------------------------------------------------------------------------------------
void acg_r1cs_free(acg_r1cs* m) {
    if (!m) return;
    if (m->ctx) cudaSetDevice(m->ctx->device);
    for (int k = 0; k < 3; ++k) {
        cudaFree(m->d_rowptr[k]);
        cudaFree(m->d_col[k]);
        cudaFree(m->d_val[k]);
    }
    cudaFree(m->d_stream);
    cudaFree(m->d_meta);
    cudaFree(m->d_far_cols);
    cudaFree(m->d_long_rows);
    cudaFree(m->d_long_desc);
    cudaFree(m->d_long_counter);
    cudaFree(m->d_runs);
    cudaFree(m->d_run_far);
    delete m;
}

int acg_r1cs_upload(acg_ctx* ctx, uint32_t n_rows, uint32_t n_cols, const acg_csr* A, const acg_csr* B,
                    const acg_csr* C, uint32_t row_begin, uint32_t row_end, acg_r1cs** out) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!out || !A || !B || !C || row_begin > row_end || row_end > n_rows || n_cols == 0)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_upload: bad argument");
    if (n_cols > kColMask) return fail(ctx, ACG_ERR_UNSUPPORTED, "acg_r1cs_upload: more than 2^30 witness columns");
    *out = nullptr;
    const acg_csr* src[3] = {A, B, C};
    const uint32_t n_local = row_end - row_begin;
    HostRows hr;
    {
        const char* why = nullptr;
        int prc = host_rows(ctx->field, n_rows, n_cols, src, row_begin, row_end, hr, &why);
        if (prc) return fail(ctx, prc, std::string("acg_r1cs_upload: ") + (why ? why : "bad argument"));
    }
    std::vector<uint32_t> (&local_rp)[3] = hr.local_rp, (&tagged_col)[3] = hr.tagged_col, (&gcum)[3] = hr.gcum, (&ccum)[3] = hr.ccum;
    acg_r1cs* m = new (std::nothrow) acg_r1cs();
    if (!m) return ACG_ERR_OOM;
    m->ctx = ctx;
    {
        std::vector<bool> seen(n_cols, false);
        uint64_t distinct = 0;
        for (int k = 0; k < 3; ++k)
            for (uint32_t c : tagged_col[k]) {
                const uint32_t col = c & kColMask;
                if (!seen[col]) {
                    seen[col] = true;
                    ++distinct;
                }
            }
        m->distinct_cols = distinct;
    }
    m->n_rows_total = n_rows;
    m->n_cols = n_cols;
    m->row_begin = row_begin;
    m->row_end = row_end;
    struct Guard {
        acg_r1cs* p;
        ~Guard() {
            if (p) acg_r1cs_free(p);
        }
    } guard{m};

    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    uint32_t launches = 0;
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    for (int k = 0; k < 3; ++k) {
        const acg_csr* M = src[k];
        const uint32_t e0 = M->rowptr[row_begin];
        const uint64_t cnt = tagged_col[k].size();
        m->nnz[k] = cnt;
        // pads: TMA slices are rounded up to 16 bytes and may read a few elements past the end
        CU(ctx, cudaMalloc(&m->d_rowptr[k], ((size_t)n_local + 1 + 8) * sizeof(uint32_t)));
        CU(ctx, cudaMalloc(&m->d_col[k], (size_t)(cnt + 8) * sizeof(uint32_t)));
        CU(ctx, cudaMalloc(&m->d_val[k], (size_t)(cnt + 1) * sizeof(fr_t)));
        CU(ctx, cudaMemsetAsync(m->d_rowptr[k] + n_local + 1, 0, 8 * sizeof(uint32_t), ctx->stream));
        CU(ctx, cudaMemsetAsync(m->d_col[k] + cnt, 0, 8 * sizeof(uint32_t), ctx->stream));
        CU(ctx, cudaMemcpyAsync(m->d_rowptr[k], local_rp[k].data(), ((size_t)n_local + 1) * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, ctx->stream));
        if (cnt) {
            CU(ctx, cudaMemcpyAsync(m->d_col[k], tagged_col[k].data(), cnt * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                    ctx->stream));
            CU(ctx, cudaMemcpyAsync(m->d_val[k], M->val + 4ull * e0, cnt * sizeof(fr_t), cudaMemcpyHostToDevice,
                                    ctx->stream));
            CU(ctx, launch_to_mont(ctx->field, m->d_val[k], cnt, ctx->d_flag, ctx->stream));
            ++launches;
        }
        m->dev.m[k].rowptr = m->d_rowptr[k];
        m->dev.m[k].col = m->d_col[k];
        m->dev.m[k].val = m->d_val[k];
    }
    // tiles and their static lists of general entries (pool indices, uint16)
    const uint32_t* rp[3] = {local_rp[0].data(), local_rp[1].data(), local_rp[2].data()};
    const uint32_t* gc[3] = {gcum[0].data(), gcum[1].data(), gcum[2].data()};
    const uint32_t* cc[3] = {ccum[0].data(), ccum[1].data(), ccum[2].data()};
    // ---- tile stream (see kernels.h): one self-contained blob per tile, entries in ELL order
    m->variant = pick_variant(ctx->tiled_variant, hr, n_local);
    const TileGeometry geom = kTileGeom[m->variant];
    std::vector<HostTile> tiles;
    build_tiles(geom, rp, gc, cc, n_local, tiles, m->long_ranges);
    TileStream tstream;
    {
        const uint64_t* const val0[3] = {src[0]->val + 4ull * src[0]->rowptr[row_begin], src[1]->val + 4ull * src[1]->rowptr[row_begin],
                                         src[2]->val + 4ull * src[2]->rowptr[row_begin]};
        int brc = build_tile_stream(geom, n_cols, tiles, local_rp, tagged_col, val0, upload_threads(), tstream);
        if (brc == ACG_ERR_UNSUPPORTED) return fail(ctx, brc, "acg_r1cs_upload: tile stream too large");
        if (brc) return fail(ctx, brc, "acg_r1cs_upload: tile-stream build failed");
    }
    std::vector<uint8_t>& stream = tstream.stream;
    std::vector<uint32_t>& gval_offs = tstream.gval_offs;
    std::vector<TileMeta>& metas = tstream.metas;
    std::vector<uint32_t>& far_all = tstream.far_all;
    m->long_ranges.insert(m->long_ranges.end(), tstream.extra_long.begin(), tstream.extra_long.end());
    const uint32_t n_tiles_out = (uint32_t)metas.size();
    std::sort(m->long_ranges.begin(), m->long_ranges.end());
    {   // the rows the tiles leave out, as one list
        std::vector<uint32_t> long_rows;
        for (const auto& lr : m->long_ranges)
            for (uint32_t r = lr.first; r < lr.second; ++r) long_rows.push_back(r);
        m->n_long_rows = (uint32_t)long_rows.size();
        if (m->n_long_rows) {
            CU(ctx, cudaMalloc(&m->d_long_rows, long_rows.size() * sizeof(uint32_t)));
            CU(ctx, cudaMemcpy(m->d_long_rows, long_rows.data(), long_rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    m->n_tiles = n_tiles_out;
    // what one check streams from HBM besides the witness: blobs, tile records, far column lists
    m->blob_bytes = stream.size();
    m->stream_bytes = stream.size() + metas.size() * sizeof(TileMeta) + far_all.size() * sizeof(uint32_t);
    DevBuf d_goffs;
    CU(ctx, cudaMalloc(&m->d_stream, stream.size() + 64));
    CU(ctx, cudaMalloc(&m->d_meta, (metas.size() + 1) * sizeof(TileMeta)));
    CU(ctx, cudaMalloc(&m->d_far_cols, (far_all.size() + 4) * sizeof(uint32_t)));
    if (!metas.empty())
        CU(ctx, cudaMemcpyAsync(m->d_meta, metas.data(), metas.size() * sizeof(TileMeta), cudaMemcpyHostToDevice,
                                ctx->stream));
    if (!far_all.empty())
        CU(ctx, cudaMemcpyAsync(m->d_far_cols, far_all.data(), far_all.size() * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, d_goffs.alloc(gval_offs.size() * sizeof(uint32_t)));
    CU(ctx, cudaMemcpyAsync(m->d_stream, stream.data(), stream.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!gval_offs.empty()) {
        CU(ctx, cudaMemcpyAsync(d_goffs.p, gval_offs.data(), gval_offs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                ctx->stream));
        CU(ctx, launch_to_mont_scattered(ctx->field, m->d_stream, d_goffs.as<uint32_t>(), gval_offs.size(), ctx->d_flag,
                                         ctx->stream));
        ++launches;
    }
    // weighted runs of the tiled kernel's CTAs (kernels.h CtaRun) when the system fills the chip
    {
        const uint32_t ctas = tiled_ctas_per_sm(m->variant);
        const uint32_t grid = (uint32_t)ctx->sm_count * ctas;
        double wgt[8];
        const unsigned n_w = n_tiles_out >= grid ? tile_rate_weights(m->variant, ctas, wgt) : 0u;
        const std::vector<uint32_t>* place = nullptr;
        if (n_w) {
            int prc = get_placement(ctx, m->variant, &place);
            if (prc) return prc;
        }
        if (n_w && place && place->size() == grid) {
            const uint32_t n_sm = (uint32_t)ctx->sm_count;
            // rank of every block among the blocks of its SM (arrival order = block order), SMs numbered by first use
            std::map<uint32_t, uint32_t> sm_pos, sm_seen;
            std::vector<uint32_t> rank(grid), pos(grid);
            for (uint32_t b = 0; b < grid; ++b) {
                const uint32_t sm = (*place)[b];
                if (!sm_pos.count(sm)) {
                    const uint32_t p = (uint32_t)sm_pos.size();
                    sm_pos[sm] = p;
                }
                pos[b] = sm_pos[sm];
                rank[b] = sm_seen[sm]++;
            }
            // every SM gets the same number of tiles (+-1); inside an SM they are dealt to its CTAs one at a time, always
            // to the CTA that would finish its share first (list scheduling with the per-rank tile times 1 / weight);
            // the tiles of rank k over all SMs form one contiguous region of the stream, so that every SM works on every
            // part of the system (later rows gather from a wider part of the witness and cost more)
            std::vector<std::vector<uint32_t>> cnt(n_sm, std::vector<uint32_t>(n_w, 0));
            for (uint32_t p = 0; p < n_sm; ++p) {
                const uint32_t total = n_tiles_out / n_sm + (p < n_tiles_out % n_sm ? 1u : 0u);
                for (uint32_t i = 0; i < total; ++i) {
                    unsigned best = 0;
                    double best_t = 1e300;
                    for (unsigned k = 0; k < n_w; ++k) {
                        const double t = (double)(cnt[p][k] + 1u) / wgt[k];
                        if (t < best_t) {
                            best_t = t;
                            best = k;
                        }
                    }
                    ++cnt[p][best];
                }
            }
            // stream order of the blocks: rank by rank, inside a rank SM position by SM position
            std::vector<uint32_t> block_at((size_t)n_w * n_sm);
            for (uint32_t b = 0; b < grid; ++b) block_at[(size_t)rank[b] * n_sm + pos[b]] = b;
            m->run_order.assign(block_at.begin(), block_at.end());
            m->run_cnt.assign(grid, 0u);
            for (uint32_t b = 0; b < grid; ++b) m->run_cnt[b] = cnt[pos[b]][rank[b]];
            m->h_meta = metas;
            m->h_far_cols = far_all;
            int urc = upload_runs(ctx, m);
            if (urc) return urc;
        }
    }
    m->dev.tagged = 1;
    if (m->n_long_rows) {  // the description of the long rows the tiled kernel reads (kernels.h DevLongRows)
        CU(ctx, cudaMalloc(&m->d_long_counter, sizeof(unsigned int)));
        CU(ctx, cudaMemset(m->d_long_counter, 0, sizeof(unsigned int)));
        const DevLongRows desc{m->dev, m->d_long_rows, m->n_long_rows, m->d_long_counter};
        CU(ctx, cudaMalloc(&m->d_long_desc, sizeof desc));
        CU(ctx, cudaMemcpy(m->d_long_desc, &desc, sizeof desc, cudaMemcpyHostToDevice));
    }
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += launches;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), 0.f, 0.f, launches, 0};
    if (*ctx->h_flag) return fail(ctx, ACG_ERR_NON_CANONICAL, "acg_r1cs_upload: matrix coefficient >= modulus");
    guard.p = nullptr;
    *out = m;
    return ACG_OK;
    ACG_CATCH(ctx)
}

// Host-only (no device needed): builds the tile stream acg_r1cs_upload would build for this slice and geometry on
// n_threads worker threads (0: as the upload chooses) and returns digests of everything it produced -- the test that
// the multi-threaded build is the single-threaded one, byte for byte.
int acg_tile_stream_digest(int field_id, int variant, uint32_t n_rows, uint32_t n_cols, const acg_csr* A, const acg_csr* B,
                           const acg_csr* C, uint32_t row_begin, uint32_t row_end, uint32_t n_threads, uint64_t* out4) {
    try {
        if (!A || !B || !C || !out4 || row_begin > row_end || row_end > n_rows || n_cols == 0 || n_cols > kColMask ||
            variant < 0 || variant >= kNumTileVariants || (field_id != ACG_FIELD_BN254_FR && field_id != ACG_FIELD_BLS12_381_FR))
            return ACG_ERR_BAD_ARG;
        const acg_csr* src[3] = {A, B, C};
        HostRows hr;
        int rc = host_rows(field_id, n_rows, n_cols, src, row_begin, row_end, hr, nullptr);
        if (rc) return rc;
        const uint32_t n_local = row_end - row_begin;
        const int v = pick_variant(variant, hr, n_local);
        const uint32_t* rp[3] = {hr.local_rp[0].data(), hr.local_rp[1].data(), hr.local_rp[2].data()};
        const uint32_t* gc[3] = {hr.gcum[0].data(), hr.gcum[1].data(), hr.gcum[2].data()};
        const uint32_t* cc[3] = {hr.ccum[0].data(), hr.ccum[1].data(), hr.ccum[2].data()};
        std::vector<HostTile> tiles;
        std::vector<std::pair<uint32_t, uint32_t>> long_ranges;
        build_tiles(kTileGeom[v], rp, gc, cc, n_local, tiles, long_ranges);
        const uint64_t* const val0[3] = {A->val + 4ull * A->rowptr[row_begin], B->val + 4ull * B->rowptr[row_begin],
                                         C->val + 4ull * C->rowptr[row_begin]};
        TileStream ts;
        rc = build_tile_stream(kTileGeom[v], n_cols, tiles, hr.local_rp, hr.tagged_col, val0,
                               n_threads ? n_threads : upload_threads(), ts);
        if (rc) return rc;
        long_ranges.insert(long_ranges.end(), ts.extra_long.begin(), ts.extra_long.end());
        std::sort(long_ranges.begin(), long_ranges.end());
        if (const char* wrong = validate_tile_stream(kTileGeom[v], n_local, n_cols, ts, long_ranges)) {
            fprintf(stderr, "acg_tile_stream_digest: invalid tile stream: %s\n", wrong);
            return ACG_ERR_INTERNAL;
        }
        auto fnv = [](uint64_t h, const void* p, size_t n) {
            const uint8_t* b = static_cast<const uint8_t*>(p);
            for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
            return h;
        };
        uint64_t h1 = fnv(0xcbf29ce484222325ull, ts.stream.data(), ts.stream.size());
        uint64_t h2 = 0xcbf29ce484222325ull;
        for (const TileMeta& tm : ts.metas) {  // field by field: the struct may have padding
            const uint32_t f[6] = {tm.blob_off16, tm.blob_bytes, tm.win_lo, tm.win_n, tm.far_off, tm.n_far};
            h2 = fnv(h2, f, sizeof f);
        }
        h2 = fnv(h2, ts.far_all.data(), ts.far_all.size() * sizeof(uint32_t));
        h2 = fnv(h2, ts.gval_offs.data(), ts.gval_offs.size() * sizeof(uint32_t));
        for (const auto& lr : long_ranges) {
            const uint32_t f[2] = {lr.first, lr.second};
            h2 = fnv(h2, f, sizeof f);
        }
        out4[0] = h1;
        out4[1] = h2;
        out4[2] = ts.stream.size();
        out4[3] = ts.metas.size();
        return ACG_OK;
    } catch (const std::bad_alloc&) {
        return ACG_ERR_OOM;
    } catch (...) {
        return ACG_ERR_INTERNAL;
    }
}

uint64_t acg_r1cs_algorithmic_bytes(const acg_r1cs* m) {
    if (!m) return 0;
    const uint64_t rows = m->row_end - m->row_begin;
    uint64_t b = 32ull * m->distinct_cols + 8ull;
    for (int k = 0; k < 3; ++k) b += m->nnz[k] * 36ull + 4ull * (rows + 1);
    return b;
}

uint64_t acg_r1cs_stream_bytes(const acg_r1cs* m) { return m ? m->stream_bytes : 0; }

int acg_r1cs_set_row_offset(acg_r1cs* m, uint64_t offset) {
    if (!m) return ACG_ERR_BAD_ARG;
    m->row_offset = offset;
    return ACG_OK;
}

void acg_vec_free(acg_vec* v) {
    if (!v) return;
    if (v->ctx) cudaSetDevice(v->ctx->device);
    if (v->async_pending && v->ctx) cudaStreamSynchronize(v->ctx->copy_stream);
    cudaFree(v->d);
    cudaFree(v->d_bad);
    if (v->ready) cudaEventDestroy(v->ready);
    if (v->used) cudaEventDestroy(v->used);
    delete v;
}
uint32_t acg_vec_len(const acg_vec* v) { return v ? v->n : 0; }
void* acg_vec_device_ptr(acg_vec* v) { return v ? v->d : nullptr; }

// an asynchronous update of v is still in flight: let it finish before the vector is touched from the main stream
static int vec_settle(acg_ctx* ctx, acg_vec* v) {
    if (v->async_pending) {
        CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
        v->async_pending = false;
    }
    return ACG_OK;
}

int acg_witness_update(acg_ctx* ctx, acg_vec* v, const uint64_t* w, uint32_t n_cols) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!v || v->ctx != ctx || !w || v->n != n_cols) return fail(ctx, ACG_ERR_BAD_ARG, "acg_witness_update: bad argument");
    if ((rc = vec_settle(ctx, v))) return rc;
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = upload_canonical_staged(ctx, v->d, w, n_cols);
    if (rc) return rc;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), 0.f, 0.f, 2, 0};
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_witness_update_range(acg_ctx* ctx, acg_vec* v, const uint64_t* w, uint32_t first, uint32_t count) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!v || v->ctx != ctx || (count && !w) || first > v->n || count > v->n - first)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_witness_update_range: bad argument");
    if (count == 0) return ACG_OK;
    if ((rc = vec_settle(ctx, v))) return rc;
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = upload_canonical_staged(ctx, v->d + first, w, count);
    if (rc) return rc;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), 0.f, 0.f, 2, 0};
    return ACG_OK;
    ACG_CATCH(ctx)
}

// Enqueue only, on the context's copy stream: H2D straight into the vector, validation + conversion in place.  The
// update waits for the checks of v enqueued so far, and later checks of v wait for it -- so with two vectors the
// upload of witness i + 1 overlaps the check of witness i.  A non-canonical element is reported by the next blocking
// call that reads v (acg_r1cs_check, acg_vec_status) as ACG_ERR_NON_CANONICAL; the vector's contents are then
// undefined.  `w` must stay valid (and should be pinned) until that call.
int acg_witness_update_async(acg_ctx* ctx, acg_vec* v, const uint64_t* w, uint32_t first, uint32_t count) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!v || v->ctx != ctx || (count && !w) || first > v->n || count > v->n - first)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_witness_update_async: bad argument");
    if (count == 0) return ACG_OK;
    if (!v->ready) {
        CU(ctx, cudaEventCreateWithFlags(&v->ready, cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&v->used, cudaEventDisableTiming));
        CU(ctx, cudaMalloc(&v->d_bad, sizeof(int)));
        CU(ctx, cudaMemsetAsync(v->d_bad, 0, sizeof(int), ctx->copy_stream));
    }
    cudaStream_t cs = ctx->copy_stream;
    if (v->has_used) CU(ctx, cudaStreamWaitEvent(cs, v->used, 0));  // the checks of v enqueued so far
    CU(ctx, cudaMemcpyAsync(v->d + first, w, (size_t)count * sizeof(fr_t), cudaMemcpyHostToDevice, cs));
    CU(ctx, launch_to_mont(ctx->field, v->d + first, count, v->d_bad, cs));
    CU(ctx, cudaEventRecord(v->ready, cs));
    ++ctx->launches;
    v->async_pending = true;
    return ACG_OK;
    ACG_CATCH(ctx)
}

// Enqueue only: `stream` waits for the asynchronous updates of v enqueued so far (for callers that read or complete the
// vector with their own device work -- e.g. the NVLink all-gather of the slices of a row-sharded witness -- before a
// check), and the next asynchronous update of v will wait for what `stream` holds when this vector is next checked.
int acg_vec_stream_wait(acg_ctx* ctx, acg_vec* v, void* stream) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!v || v->ctx != ctx) return fail(ctx, ACG_ERR_BAD_ARG, "acg_vec_stream_wait: bad argument");
    return vec_acquire(ctx, v, static_cast<cudaStream_t>(stream));
    ACG_CATCH(ctx)
}

// Blocking: waits for the asynchronous updates of v and returns ACG_ERR_NON_CANONICAL if one of them met an element
// >= r (the flag is cleared), ACG_OK otherwise.
int acg_vec_status(acg_ctx* ctx, acg_vec* v) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!v || v->ctx != ctx) return fail(ctx, ACG_ERR_BAD_ARG, "acg_vec_status: bad argument");
    if (!v->d_bad) return ACG_OK;
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, v->d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
    v->async_pending = false;
    if (*ctx->h_flag) {
        CU(ctx, cudaMemsetAsync(v->d_bad, 0, sizeof(int), ctx->copy_stream));
        return fail(ctx, ACG_ERR_NON_CANONICAL, "an asynchronous witness update met a field element >= modulus");
    }
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_witness_upload(acg_ctx* ctx, const uint64_t* w, uint32_t n_cols, acg_vec** out) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!out || !w || n_cols == 0) return fail(ctx, ACG_ERR_BAD_ARG, "acg_witness_upload: bad argument");
    *out = nullptr;
    acg_vec* v = new (std::nothrow) acg_vec();
    if (!v) return ACG_ERR_OOM;
    v->ctx = ctx;
    v->n = n_cols;
    cudaError_t e = cudaMalloc(&v->d, (size_t)n_cols * sizeof(fr_t));
    if (e != cudaSuccess) {
        delete v;
        return fail_cuda(ctx, e, "cudaMalloc(witness)");
    }
    rc = acg_witness_update(ctx, v, w, n_cols);
    if (rc) {
        acg_vec_free(v);
        return rc;
    }
    *out = v;
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_vec_download(acg_ctx* ctx, const acg_vec* v, uint64_t* out, uint32_t n) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!v || !out || v->ctx != ctx || n != v->n) return fail(ctx, ACG_ERR_BAD_ARG, "acg_vec_download: bad argument");
    if ((rc = vec_settle(ctx, const_cast<acg_vec*>(v)))) return rc;
    DevBuf tmp;
    CU(ctx, tmp.alloc((size_t)n * sizeof(fr_t)));
    CU(ctx, cudaMemcpyAsync(tmp.p, v->d, (size_t)n * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, launch_from_mont(ctx->field, tmp.as<fr_t>(), n, ctx->stream));
    ++ctx->launches;
    CU(ctx, cudaMemcpyAsync(out, tmp.p, (size_t)n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_generate_assignment_device(acg_ctx* ctx, const acg_circuit* c, const uint32_t* input_ix,
                                   const uint64_t* input_vals, uint32_t n_inputs, uint32_t n_in, uint32_t n_mid,
                                   uint32_t n_out, acg_vec** out, uint32_t* n_levels_out) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!c || !out || (n_inputs && (!input_ix || !input_vals)) || c->field != ctx->field)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_generate_assignment_device: bad argument");
    *out = nullptr;
    for (uint32_t i = 0; i < n_inputs; ++i) {
        if (input_ix[i] >= 0x7FFFFFFFu) return fail(ctx, ACG_ERR_BAD_ARG, "acg_generate_assignment_device: input index");
        n_in = std::max(n_in, input_ix[i] + 1u);
    }
    host::GatePlan plan;
    rc = host::build_gate_plan(c, n_in, n_mid, n_out, plan);
    if (rc == ACG_ERR_UNSUPPORTED)
        return fail(ctx, rc, "acg_generate_assignment_device: gate list is not in single-assignment, define-before-use "
                             "form; use acg_generate_assignment (sequential fold)");
    if (rc) return fail(ctx, rc, "acg_generate_assignment_device: cannot plan the circuit");
    if (!plan.required_inputs.empty()) {  // an Equal / Split gate on an input the caller did not supply: the reference panics
        std::vector<uint8_t> supplied((size_t)plan.n_in + 1, 0);
        for (uint32_t i = 0; i < n_inputs; ++i) supplied[(size_t)input_ix[i] + 1] = 1;
        for (uint32_t col : plan.required_inputs)
            if (!supplied[col])
                return fail(ctx, ACG_ERR_BAD_ARG, "acg_generate_assignment_device: an Equal or Split gate reads an input wire "
                                                  "that was not supplied (src/QAP.hs:445,474)");
    }
    static_assert(sizeof(host::GateRec) == sizeof(WitnessGate), "gate record layouts differ");
    const uint64_t n_cols64 = 1ull + plan.n_in + plan.n_mid + plan.n_out;
    if (n_cols64 > kColMask) return fail(ctx, ACG_ERR_UNSUPPORTED, "acg_generate_assignment_device: too many wires");
    const uint32_t n_cols = (uint32_t)n_cols64;
    // initial witness: constant 1 (initialQapSet, src/QAP.hs:591-595), the inputs, zero elsewhere (missing = 0)
    std::vector<uint64_t> init((size_t)4 * (1 + plan.n_in), 0);
    init[0] = 1;
    for (uint32_t i = 0; i < n_inputs; ++i) std::memcpy(&init[4 * (1 + (size_t)input_ix[i])], input_vals + 4ull * i, 32);
    acg_vec* v = new (std::nothrow) acg_vec();
    if (!v) return ACG_ERR_OOM;
    v->ctx = ctx;
    v->n = n_cols;
    struct Guard {
        acg_vec* p;
        ~Guard() {
            if (p) acg_vec_free(p);
        }
    } guard{v};
    CU(ctx, cudaMalloc(&v->d, (size_t)n_cols * sizeof(fr_t)));
    CU(ctx, cudaMemsetAsync(v->d, 0, (size_t)n_cols * sizeof(fr_t), ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = upload_canonical(ctx, v->d, init.data(), 1 + plan.n_in);  // rejects inputs >= r
    if (rc) return rc;
    DevBuf d_gates, d_lvl, d_tcol, d_tcoef, d_souts;
    const size_t n_terms = plan.term_col.size();
    CU(ctx, d_gates.alloc(plan.gates.size() * sizeof(WitnessGate)));
    CU(ctx, d_lvl.alloc(plan.level_ptr.size() * sizeof(uint32_t)));
    CU(ctx, d_tcol.alloc(n_terms * sizeof(uint32_t)));
    CU(ctx, d_tcoef.alloc(n_terms * sizeof(fr_t)));
    CU(ctx, d_souts.alloc(plan.split_outs.size() * sizeof(uint32_t)));
    auto h2d = [&](DevBuf& d, const void* src, size_t bytes) {
        return bytes ? cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
    };
    CU(ctx, h2d(d_gates, plan.gates.data(), plan.gates.size() * sizeof(WitnessGate)));
    CU(ctx, h2d(d_lvl, plan.level_ptr.data(), plan.level_ptr.size() * sizeof(uint32_t)));
    CU(ctx, h2d(d_tcol, plan.term_col.data(), n_terms * sizeof(uint32_t)));
    CU(ctx, h2d(d_tcoef, plan.term_coef.data(), n_terms * sizeof(fr_t)));  // already Montgomery (host mirror)
    CU(ctx, h2d(d_souts, plan.split_outs.data(), plan.split_outs.size() * sizeof(uint32_t)));
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    const uint32_t n_levels = plan.level_ptr.empty() ? 0u : (uint32_t)plan.level_ptr.size() - 1u;
    CU(ctx, launch_witness_levels(ctx->field, d_gates.as<WitnessGate>(), d_lvl.as<uint32_t>(), n_levels, plan.max_width,
                                  d_tcol.as<uint32_t>(), d_tcoef.as<fr_t>(), d_souts.as<uint32_t>(), v->d,
                                  ctx->sm_count, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += n_levels ? 2 : 1;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), elapsed(ctx->ev[1], ctx->ev[2]), 0.f, n_levels ? 2u : 1u, 0};
    if (n_levels_out) *n_levels_out = n_levels;
    guard.p = nullptr;
    *out = v;
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_r1cs_check_async(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, uint64_t* d_result, void* stream) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!m || !w || !d_result || m->ctx != ctx || w->ctx != ctx || w->n != m->n_cols)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_check_async: bad argument");
    uint32_t launches = 0;
    rc = enqueue_check(ctx, m, w, reinterpret_cast<unsigned long long*>(d_result), nullptr, nullptr, nullptr,
                       static_cast<cudaStream_t>(stream), &launches);
    ctx->launches += launches;
    ctx->timing.kernel_launches = launches;
    return rc;
    ACG_CATCH(ctx)
}

int acg_r1cs_check(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, uint64_t* n_violations,
                   uint64_t* first_bad_row) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!m || !w || m->ctx != ctx || w->ctx != ctx || w->n != m->n_cols)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_check: bad argument");
    uint32_t launches = 0;
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    rc = enqueue_check(ctx, m, w, ctx->d_result, nullptr, nullptr, nullptr, ctx->stream, &launches);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            ctx->stream));
    const bool async_verdict = w->async_pending && w->d_bad;  // (the check waited for that update: its flag is final)
    if (async_verdict)
        CU(ctx, cudaMemcpyAsync(ctx->h_flag, w->d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += launches;
    ctx->timing = acg_timing{0.f, elapsed(ctx->ev[1], ctx->ev[2]), elapsed(ctx->ev[2], ctx->ev[3]), launches, 0};
    if (async_verdict) {
        const_cast<acg_vec*>(w)->async_pending = false;
        if (*ctx->h_flag) {
            CU(ctx, cudaMemsetAsync(w->d_bad, 0, sizeof(int), ctx->stream));
            return fail(ctx, ACG_ERR_NON_CANONICAL, "acg_r1cs_check: the witness update met a field element >= modulus");
        }
    }
    if (n_violations) *n_violations = ctx->h_result[0];
    if (first_bad_row) *first_bad_row = ctx->h_result[1];
    return ACG_OK;
    ACG_CATCH(ctx)
}

// One-shot path: no host-side preprocessing (a single check cannot amortise it).  Plain copies from the
// caller's buffers on the context stream, validation + Montgomery conversion + row-wise check on the
// device, stream-ordered allocations.
int acg_r1cs_check_host(acg_ctx* ctx, uint32_t n_rows, uint32_t n_cols, const acg_csr* A, const acg_csr* B,
                        const acg_csr* C, const uint64_t* w, uint64_t* n_violations, uint64_t* first_bad_row) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!A || !B || !C || !w || n_cols == 0 || n_cols > kColMask)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_check_host: bad argument");
    const acg_csr* src[3] = {A, B, C};
    for (int k = 0; k < 3; ++k)
        if (!src[k]->rowptr || (src[k]->nnz && (!src[k]->col || !src[k]->val)) || src[k]->nnz > 0xFFFFFFF0ull)
            return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_check_host: null array or nnz too large");
    cudaStream_t s = ctx->stream;
    struct Pool {  // stream-ordered scratch, released in reverse order on every exit path
        cudaStream_t s;
        std::vector<void*> ptrs;
        ~Pool() {
            for (auto it = ptrs.rbegin(); it != ptrs.rend(); ++it) cudaFreeAsync(*it, s);
        }
        cudaError_t get(void** p, size_t bytes) {
            cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, s);
            if (e == cudaSuccess) ptrs.push_back(*p);
            return e;
        }
    } pool{s, {}};
    DevR1cs dev{};
    uint32_t launches = 0;
    fr_t* d_w = nullptr;
    CU(ctx, cudaEventRecord(ctx->ev[0], s));
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), s));
    for (int k = 0; k < 3; ++k) {
        const acg_csr* M = src[k];
        uint32_t *d_rp = nullptr, *d_col = nullptr;
        fr_t* d_val = nullptr;
        CU(ctx, pool.get((void**)&d_rp, ((size_t)n_rows + 1) * sizeof(uint32_t)));
        CU(ctx, pool.get((void**)&d_col, (size_t)M->nnz * sizeof(uint32_t)));
        CU(ctx, pool.get((void**)&d_val, (size_t)M->nnz * sizeof(fr_t)));
        CU(ctx, cudaMemcpyAsync(d_rp, M->rowptr, ((size_t)n_rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        if (M->nnz) {
            CU(ctx, cudaMemcpyAsync(d_col, M->col, (size_t)M->nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            CU(ctx, cudaMemcpyAsync(d_val, M->val, (size_t)M->nnz * sizeof(fr_t), cudaMemcpyHostToDevice, s));
            CU(ctx, launch_to_mont(ctx->field, d_val, M->nnz, ctx->d_flag, s));
            ++launches;
        }
        CU(ctx, launch_validate_csr(d_rp, d_col, n_rows, M->nnz, n_cols, ctx->d_flag, s));
        ++launches;
        dev.m[k].rowptr = d_rp;
        dev.m[k].col = d_col;
        dev.m[k].val = d_val;
    }
    CU(ctx, pool.get((void**)&d_w, (size_t)n_cols * sizeof(fr_t)));
    CU(ctx, cudaMemcpyAsync(d_w, w, (size_t)n_cols * sizeof(fr_t), cudaMemcpyHostToDevice, s));
    CU(ctx, launch_to_mont(ctx->field, d_w, n_cols, ctx->d_flag, s));
    ++launches;
    // the check must not run on unvalidated indices: read the flag back first
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaEventRecord(ctx->ev[1], s));
    CU(ctx, cudaStreamSynchronize(s));
    if (*ctx->h_flag & 2) return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_check_host: malformed CSR (rowptr / column index)");
    if (*ctx->h_flag & 1) return fail(ctx, ACG_ERR_NON_CANONICAL, "acg_r1cs_check_host: field element >= modulus");
    dev.tagged = 0;
    const CheckEpilogue fin{ctx->d_accum, ctx->d_ticket, ctx->d_result, PeerSlots{}, 0ull, 0u};
    if (n_rows == 0) {
        CU(ctx, launch_init_result(ctx->d_result, s));
        ++launches;
    }
    if (n_rows) {
        CU(ctx, launch_r1cs_rowwise(ctx->field, dev, d_w, 0, n_rows, 0, fin, nullptr, nullptr, nullptr, s));
        ++launches;
    }
    CU(ctx, cudaEventRecord(ctx->ev[2], s));
    CU(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaEventRecord(ctx->ev[3], s));
    CU(ctx, cudaStreamSynchronize(s));
    ctx->launches += launches;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), elapsed(ctx->ev[1], ctx->ev[2]),
                             elapsed(ctx->ev[2], ctx->ev[3]), launches, 0};
    if (n_violations) *n_violations = ctx->h_result[0];
    if (first_bad_row) *first_bad_row = ctx->h_result[1];
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_r1cs_eval(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, uint64_t* Aw, uint64_t* Bw, uint64_t* Cw) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!m || !w || m->ctx != ctx || w->ctx != ctx || w->n != m->n_cols)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_eval: bad argument");
    const uint32_t n_local = m->row_end - m->row_begin;
    if (n_local == 0) return ACG_OK;
    DevBuf buf[3];
    uint64_t* host[3] = {Aw, Bw, Cw};
    for (int k = 0; k < 3; ++k) CU(ctx, buf[k].alloc((size_t)n_local * sizeof(fr_t)));
    uint32_t launches = 0;
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    rc = enqueue_check(ctx, m, w, ctx->d_result, buf[0].as<fr_t>(), buf[1].as<fr_t>(), buf[2].as<fr_t>(),
                       ctx->stream, &launches);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    for (int k = 0; k < 3; ++k) {
        if (!host[k]) continue;
        rc = download_canonical(ctx, buf[k].as<fr_t>(), n_local, host[k]);
        if (rc) return rc;
        ++launches;
    }
    CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += launches;
    ctx->timing = acg_timing{0.f, elapsed(ctx->ev[1], ctx->ev[2]), elapsed(ctx->ev[2], ctx->ev[3]), launches, 0};
    return ACG_OK;
    ACG_CATCH(ctx)
}

// --------------------------------------------------------------------------------------------------
// NTT
// --------------------------------------------------------------------------------------------------
int acg_ntt_device(acg_ctx* ctx, acg_vec* v, uint32_t log_n, int inverse, void* stream) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!v || v->ctx != ctx || log_n > 31 || v->n != (1u << log_n))
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_ntt_device: bad argument");
    if (log_n == 0) return ACG_OK;
    NttPlan* plan = nullptr;
    if ((rc = get_plan(ctx, log_n, inverse != 0, &plan))) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    fr_t* scratch = nullptr;
    CU(ctx, cudaMallocAsync(&scratch, (size_t)v->n * sizeof(fr_t), s));
    uint32_t launches = 0;
    cudaError_t e = ntt_run(plan, v->d, scratch, 1, s, &launches);
    cudaFreeAsync(scratch, s);
    ctx->launches += launches;
    ctx->timing.kernel_launches = launches;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "ntt_run");
    return ACG_OK;
    ACG_CATCH(ctx)
}

static int ntt_host_batched(acg_ctx* ctx, uint64_t* data, uint32_t log_n, uint32_t n_batch, bool inverse) {
    if (log_n > 31) return fail(ctx, ACG_ERR_BAD_ARG, "log_n too large");
    if ((int)log_n > two_adicity(ctx->field))
        return fail(ctx, ACG_ERR_UNSUPPORTED, "log_n exceeds the 2-adicity of the field");
    if (n_batch == 0) return ACG_OK;
    const uint64_t n = 1ull << log_n;
    NttPlan* plan = nullptr;
    int rc = ACG_OK;
    if (log_n > 0 && (rc = get_plan(ctx, log_n, inverse, &plan))) return rc;
    // chunk: a power-of-two number of transforms, at most ~256 MiB of elements at a time
    uint32_t chunk = 1;
    while ((uint64_t)chunk * 2 * n * sizeof(fr_t) <= (256ull << 20) && chunk * 2 <= n_batch) chunk *= 2;
    DevBuf d, scratch;
    CU(ctx, d.alloc((size_t)chunk * n * sizeof(fr_t)));
    CU(ctx, scratch.alloc((size_t)chunk * n * sizeof(fr_t)));
    uint32_t launches = 0;
    float h2d = 0.f, ker = 0.f, d2h = 0.f;
    for (uint32_t done = 0; done < n_batch;) {
        uint32_t cur = chunk;
        while (cur > n_batch - done) cur >>= 1;
        uint64_t* hp = data + 4ull * n * done;
        CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        rc = upload_canonical(ctx, d.as<fr_t>(), hp, (uint64_t)cur * n);
        if (rc) return rc;
        ++launches;
        CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        if (log_n > 0) CU(ctx, ntt_run(plan, d.as<fr_t>(), scratch.as<fr_t>(), cur, ctx->stream, &launches));
        CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
        rc = download_canonical(ctx, d.as<fr_t>(), (uint64_t)cur * n, hp);
        if (rc) return rc;
        ++launches;
        CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        h2d += elapsed(ctx->ev[0], ctx->ev[1]);
        ker += elapsed(ctx->ev[1], ctx->ev[2]);
        d2h += elapsed(ctx->ev[2], ctx->ev[3]);
        done += cur;
    }
    ctx->launches += launches;
    ctx->timing = acg_timing{h2d, ker, d2h, launches, 0};
    return ACG_OK;
}

int acg_ntt(acg_ctx* ctx, uint64_t* data, uint32_t log_n, int inverse) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!data) return fail(ctx, ACG_ERR_BAD_ARG, "acg_ntt: null data");
    return ntt_host_batched(ctx, data, log_n, 1, inverse != 0);
    ACG_CATCH(ctx)
}

int acg_interpolate_columns(acg_ctx* ctx, uint64_t* cols, uint32_t log_n, uint32_t n_cols_batch) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!cols && n_cols_batch) return fail(ctx, ACG_ERR_BAD_ARG, "acg_interpolate_columns: null data");
    return ntt_host_batched(ctx, cols, log_n, n_cols_batch, true);
    ACG_CATCH(ctx)
}

// --------------------------------------------------------------------------------------------------
// QAP witness polynomials
// --------------------------------------------------------------------------------------------------
}  // extern "C"

template <class P>
static int coset_tables(acg_ctx* ctx, uint32_t log_n, acg_ctx::CosetTables** out) {
    auto it = ctx->coset.find(log_n);
    if (it != ctx->coset.end()) {
        *out = &it->second;
        return ACG_OK;
    }
    acg_ctx::CosetTables t;
    t.lo_bits = (log_n + 1) / 2;
    const uint32_t lo_n = 1u << t.lo_bits, hi_n = 1u << (log_n - t.lo_bits);
    const fr_t g = host_const<P>(&P::gen_mont), gi = host_const<P>(&P::gen_inv_mont);
    CU(ctx, cudaMalloc(&t.lo, (size_t)lo_n * sizeof(fr_t)));
    CU(ctx, cudaMalloc(&t.hi, (size_t)hi_n * sizeof(fr_t)));
    CU(ctx, cudaMalloc(&t.ilo, (size_t)lo_n * sizeof(fr_t)));
    CU(ctx, cudaMalloc(&t.ihi, (size_t)hi_n * sizeof(fr_t)));
    CU(ctx, launch_fill_powers(ctx->field, t.lo, lo_n, g, 0, ctx->stream));
    CU(ctx, launch_fill_powers(ctx->field, t.hi, hi_n, g, t.lo_bits, ctx->stream));
    CU(ctx, launch_fill_powers(ctx->field, t.ilo, lo_n, gi, 0, ctx->stream));
    CU(ctx, launch_fill_powers(ctx->field, t.ihi, hi_n, gi, t.lo_bits, ctx->stream));
    ctx->launches += 4;
    ctx->coset[log_n] = t;
    *out = &ctx->coset[log_n];
    return ACG_OK;
}

template <class P>
static int qap_witness_impl(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, const uint64_t* delta, uint64_t* a_out,
                            uint64_t* b_out, uint64_t* c_out, uint64_t* h_out, int* divisible) {
    const uint32_t n_rows = m->n_rows_total;
    uint32_t log_n = 0;
    while ((1ull << log_n) < n_rows) ++log_n;
    if ((int)log_n > P::TWO_ADICITY) return fail(ctx, ACG_ERR_UNSUPPORTED, "too many rows for the field's 2-adicity");
    const uint64_t N = 1ull << log_n;
    fr_t d[3] = {fr_zero<P>(), fr_zero<P>(), fr_zero<P>()};
    bool any_delta = false;
    if (delta) {
        for (int k = 0; k < 3; ++k) {
            fr_t x = fr_from_limbs(delta + 4 * k);
            if (!fr_is_canonical<P>(x)) return fail(ctx, ACG_ERR_NON_CANONICAL, "delta >= modulus");
            d[k] = fr_to_mont<P>(x);
            any_delta = any_delta || !fr_is_zero(x);
        }
    }
    cudaStream_t s = ctx->stream;
    uint32_t launches = 0;
    WorkBuf ev[3], coef[2], hbuf, scratch;  // views of the context's work buffers
    int rc = ACG_OK;
    for (int k = 0; k < 3; ++k) {
        if ((rc = work_buffer(ctx, k, N, &ev[k].p))) return rc;
        // rows beyond the system (padding of the domain) evaluate to zero; the check kernel writes the others
        if (N > n_rows) CU(ctx, cudaMemsetAsync(ev[k].p + n_rows, 0, (N - n_rows) * sizeof(fr_t), s));
    }
    if ((rc = work_buffer(ctx, 3, N, &hbuf.p))) return rc;
    if ((rc = work_buffer(ctx, 4, N, &scratch.p))) return rc;
    CU(ctx, cudaEventRecord(ctx->ev[1], s));
    rc = enqueue_check(ctx, m, w, ctx->d_result, ev[0].as<fr_t>(), ev[1].as<fr_t>(), ev[2].as<fr_t>(), s,
                       &launches);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));

    uint64_t* outs[3] = {a_out, b_out, c_out};
    NttPlan *inv = nullptr, *fwd = nullptr;
    acg_ctx::CosetTables* ct = nullptr;
    if (log_n == 0) {
        // N = 1: a, b, c are constants, T = X - 1, h = 0 when divisible
        CU(ctx, cudaMemsetAsync(hbuf.p, 0, sizeof(fr_t), s));
    } else {
        if ((rc = get_plan(ctx, log_n, true, &inv))) return rc;
        if ((rc = get_plan(ctx, log_n, false, &fwd))) return rc;
        if ((rc = coset_tables<P>(ctx, log_n, &ct))) return rc;
        // values on the domain -> coefficients
        for (int k = 0; k < 3; ++k)
            CU(ctx, ntt_run(inv, ev[k].as<fr_t>(), scratch.as<fr_t>(), 1, s, &launches));
    }
    // hand the coefficients out (canonical) before they are overwritten by the coset evaluation
    for (int k = 0; k < 3; ++k) {
        if (!outs[k]) continue;
        CU(ctx, cudaMemcpyAsync(scratch.p, ev[k].p, N * sizeof(fr_t), cudaMemcpyDeviceToDevice, s));
        if ((rc = download_canonical(ctx, scratch.as<fr_t>(), N, outs[k]))) return rc;
        ++launches;
    }
    if (any_delta) {  // keep a and b coefficients for h += d1*b + d2*a
        for (int k = 0; k < 2; ++k) {
            if ((rc = work_buffer(ctx, 5 + k, N, &coef[k].p))) return rc;
            CU(ctx, cudaMemcpyAsync(coef[k].p, ev[k].p, N * sizeof(fr_t), cudaMemcpyDeviceToDevice, s));
        }
    }
    if (log_n > 0) {
        // coset evaluation g * w^i: scale coefficient i by g^i, forward DIF (bit-reversed order out)
        for (int k = 0; k < 3; ++k) {
            CU(ctx, launch_scale_by_powers(ctx->field, ev[k].as<fr_t>(), N, ct->hi, ct->lo, ct->lo_bits, fr_one<P>(),
                                           false, false, log_n, s));
            ++launches;
            CU(ctx, ntt_run_dif(fwd, ev[k].as<fr_t>(), 1, s, &launches));
        }
        // h on the coset = (a*b - c) / (g^N - 1), any consistent order
        const fr_t g = host_const<P>(&P::gen_mont);
        const fr_t zinv = fr_inv<P>(fr_sub<P>(host_pow<P>(g, N), fr_one<P>()));
        CU(ctx, launch_quotient_pointwise(ctx->field, ev[0].as<fr_t>(), ev[1].as<fr_t>(), ev[2].as<fr_t>(),
                                          scratch.as<fr_t>(), N, zinv, s));
        ++launches;
        CU(ctx, launch_bitrev_permute(scratch.as<fr_t>(), hbuf.as<fr_t>(), log_n, 1, s));
        ++launches;
        CU(ctx, ntt_run(inv, hbuf.as<fr_t>(), scratch.as<fr_t>(), 1, s, &launches));
        CU(ctx, launch_scale_by_powers(ctx->field, hbuf.as<fr_t>(), N, ct->ihi, ct->ilo, ct->lo_bits, fr_one<P>(),
                                       false, false, log_n, s));
        ++launches;
    }
    if (any_delta) {
        CU(ctx, launch_axpy2(ctx->field, hbuf.as<fr_t>(), coef[0].as<fr_t>(), coef[1].as<fr_t>(), d[1], d[0], N, s));
        ++launches;
    }
    CU(ctx, cudaEventRecord(ctx->ev[2], s));
    if (h_out) {
        if ((rc = download_canonical(ctx, hbuf.as<fr_t>(), N, h_out))) return rc;
        ++launches;
    }
    CU(ctx, cudaEventRecord(ctx->ev[3], s));
    CU(ctx, cudaStreamSynchronize(s));
    ctx->launches += launches;
    ctx->timing = acg_timing{0.f, elapsed(ctx->ev[1], ctx->ev[2]), elapsed(ctx->ev[2], ctx->ev[3]), launches, 0};
    if (divisible) *divisible = ctx->h_result[0] == 0 ? 1 : 0;

    // constant-size delta fix-ups on the host: T = X^N - 1
    //   x' = x + dk*T      -> x'[0] -= dk, x'[N] = dk
    //   h' = h + d1*b + d2*a + d1*d2*T - d3 -> h'[0] -= d1*d2 + d3, h'[N] = d1*d2   (axpy part done above)
    auto sub_at0 = [&](uint64_t* poly, const fr_t& v_mont) {
        fr_t x0 = fr_to_mont<P>(fr_from_limbs(poly));
        limbs_from_fr(poly, fr_from_mont<P>(fr_sub<P>(x0, v_mont)));
    };
    for (int k = 0; k < 3; ++k) {
        if (!outs[k]) continue;
        sub_at0(outs[k], d[k]);
        limbs_from_fr(outs[k] + 4 * N, fr_from_mont<P>(d[k]));
    }
    if (h_out) {
        const fr_t d12 = fr_mul<P>(d[0], d[1]);
        sub_at0(h_out, fr_add<P>(d12, d[2]));
        limbs_from_fr(h_out + 4 * N, fr_from_mont<P>(d12));
    }
    return ACG_OK;
}

extern "C" {

int acg_qap_witness(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, const uint64_t* delta, uint64_t* a,
                    uint64_t* b, uint64_t* c, uint64_t* h, int* divisible) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!m || !w || m->ctx != ctx || w->ctx != ctx || w->n != m->n_cols)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_qap_witness: bad argument");
    if (m->row_begin != 0 || m->row_end != m->n_rows_total || m->n_rows_total == 0 || m->row_offset != 0)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_qap_witness: needs the full (non-empty) system, not a row shard");
    return with_field(ctx->field, [&](auto p) {
        using P = decltype(p);
        return qap_witness_impl<P>(ctx, m, w, delta, a, b, c, h, divisible);
    });
    ACG_CATCH(ctx)
}

// --------------------------------------------------------------------------------------------------
// Lagrange
// --------------------------------------------------------------------------------------------------
int acg_lagrange(acg_ctx* ctx, const uint64_t* xs, const uint64_t* ys, uint32_t n, uint32_t n_polys,
                 uint64_t* coeffs, uint64_t* target) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!xs || n == 0 || n > 4096 || (n_polys && (!ys || !coeffs)))
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_lagrange: bad argument (1 <= n <= 4096)");
    DevBuf dx, dy, dc, dt, ds;
    const size_t np = (size_t)n_polys * n;
    CU(ctx, dx.alloc((size_t)n * sizeof(fr_t)));
    CU(ctx, dy.alloc(np * sizeof(fr_t)));
    CU(ctx, dc.alloc(np * sizeof(fr_t)));
    CU(ctx, dt.alloc((size_t)(n + 1) * sizeof(fr_t)));
    CU(ctx, ds.alloc((2 * (size_t)(n + 1) + n + np) * sizeof(fr_t)));
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    if ((rc = upload_canonical(ctx, dx.as<fr_t>(), xs, n))) return rc;
    if ((rc = upload_canonical(ctx, dy.as<fr_t>(), ys, np))) return rc;
    uint32_t launches = 2;
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(ctx, launch_lagrange(ctx->field, dx.as<fr_t>(), dy.as<fr_t>(), n, n_polys, dc.as<fr_t>(), dt.as<fr_t>(),
                            ds.as<fr_t>(), ctx->d_flag, ctx->stream, &launches));
    CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (np) {
        if ((rc = download_canonical(ctx, dc.as<fr_t>(), np, coeffs))) return rc;
        ++launches;
    }
    if (target) {
        if ((rc = download_canonical(ctx, dt.as<fr_t>(), n + 1, target))) return rc;
        ++launches;
    }
    CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += launches;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), elapsed(ctx->ev[1], ctx->ev[2]),
                             elapsed(ctx->ev[2], ctx->ev[3]), launches, 0};
    if (*ctx->h_flag) return fail(ctx, ACG_ERR_BAD_ARG, "acg_lagrange: interpolation nodes are not distinct");
    return ACG_OK;
    ACG_CATCH(ctx)
}

// --------------------------------------------------------------------------------------------------
// field ops self-test surface
// --------------------------------------------------------------------------------------------------
int acg_fr_binop(acg_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t n) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (op < 0 || op > 3 || !a || !out || (op != 3 && !b)) return fail(ctx, ACG_ERR_BAD_ARG, "acg_fr_binop: bad argument");
    if (n == 0) return ACG_OK;
    DevBuf da, db, dout;
    CU(ctx, da.alloc(n * sizeof(fr_t)));
    CU(ctx, db.alloc(n * sizeof(fr_t)));
    CU(ctx, dout.alloc(n * sizeof(fr_t)));
    if ((rc = upload_canonical(ctx, da.as<fr_t>(), a, n))) return rc;
    if ((rc = upload_canonical(ctx, db.as<fr_t>(), op == 3 ? a : b, n))) return rc;
    CU(ctx, launch_fr_binop(ctx->field, op, da.as<fr_t>(), db.as<fr_t>(), dout.as<fr_t>(), n, ctx->stream));
    if ((rc = download_canonical(ctx, dout.as<fr_t>(), n, out))) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += 4;
    ctx->timing = acg_timing{0.f, 0.f, 0.f, 4, 0};
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_poly_combine(acg_ctx* ctx, const uint64_t* polys, const uint64_t* weights, uint32_t n_polys, uint32_t len,
                     uint64_t* out) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!out || (n_polys && len && (!polys || !weights))) return fail(ctx, ACG_ERR_BAD_ARG, "acg_poly_combine: bad argument");
    if (len == 0) return ACG_OK;
    DevBuf dp, dw, dout;
    CU(ctx, dp.alloc((size_t)n_polys * len * sizeof(fr_t)));
    CU(ctx, dw.alloc((size_t)n_polys * sizeof(fr_t)));
    CU(ctx, dout.alloc((size_t)len * sizeof(fr_t)));
    if (n_polys) {
        if ((rc = upload_canonical(ctx, dp.as<fr_t>(), polys, (uint64_t)n_polys * len))) return rc;
        if ((rc = upload_canonical(ctx, dw.as<fr_t>(), weights, n_polys))) return rc;
    }
    CU(ctx, launch_poly_combine(ctx->field, dp.as<fr_t>(), dw.as<fr_t>(), n_polys, len, dout.as<fr_t>(), ctx->stream));
    if ((rc = download_canonical(ctx, dout.as<fr_t>(), len, out))) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches += 4;
    ctx->timing = acg_timing{0.f, 0.f, 0.f, 4, 0};
    return ACG_OK;
    ACG_CATCH(ctx)
}

// --------------------------------------------------------------------------------------------------
// per-wire QAP value: verificationWitnessZk on a `QAP f` (src/QAP.hs:300-327), whole on the device
// --------------------------------------------------------------------------------------------------
}  // extern "C"

struct acg_qap {
    acg_ctx* ctx = nullptr;
    uint32_t n_wires = 0, len = 0, n_target = 0;  // n_target: coefficients of the target after stripping zeros
    fr_t* d_polys[3] = {nullptr, nullptr, nullptr};  // left, right, out: n_wires * len coefficients each
    fr_t* d_target = nullptr;
    fr_t lc_inv{};       // 1 / leading coefficient of the target (Montgomery)
    bool monic = false;
};

template <class P>
static int qap_verify_impl(acg_ctx* ctx, const acg_qap* q, const uint64_t* w, const uint64_t* delta, uint64_t* h_out,
                           uint32_t h_cap, uint32_t* h_len_out, int* divisible) {
    fr_t d[3] = {fr_zero<P>(), fr_zero<P>(), fr_zero<P>()};
    if (delta) {
        for (int k = 0; k < 3; ++k) {
            const fr_t x = fr_from_limbs(delta + 4 * k);
            if (!fr_is_canonical<P>(x)) return fail(ctx, ACG_ERR_NON_CANONICAL, "delta >= modulus");
            d[k] = fr_to_mont<P>(x);
        }
    }
    // a = delta1 * T + sum_k w_k L_k and b, c likewise (src/QAP.hs:314-324): la coefficients each
    const uint32_t n = q->n_target - 1u;  // degree of the target
    const uint32_t la = std::max(q->len, q->n_target);
    const uint64_t len_p = 2ull * la - 1ull;  // formal length of p = a * b - c
    uint32_t log_m = 0;
    while ((1ull << log_m) < len_p) ++log_m;
    if ((int)log_m > P::TWO_ADICITY) return fail(ctx, ACG_ERR_UNSUPPORTED, "acg_qap_verify: polynomials too long for the field's 2-adicity");
    const uint64_t M = 1ull << log_m;
    const uint32_t h_len = len_p > n ? (uint32_t)(len_p - n) : 0u;
    if (h_len_out) *h_len_out = h_len;
    if (h_out && h_cap < h_len) return fail(ctx, ACG_ERR_BAD_ARG, "acg_qap_verify: h buffer too small (acg_qap_quotient_len)");
    cudaStream_t s = ctx->stream;
    uint32_t launches = 0;
    DevBuf ev[3], pbuf, hbuf, scratch, dw;
    for (int k = 0; k < 3; ++k) {
        CU(ctx, ev[k].alloc(M * sizeof(fr_t)));
        CU(ctx, cudaMemsetAsync(ev[k].p, 0, M * sizeof(fr_t), s));
    }
    CU(ctx, pbuf.alloc(M * sizeof(fr_t)));
    CU(ctx, scratch.alloc(M * sizeof(fr_t)));
    CU(ctx, hbuf.alloc((size_t)std::max(h_len, 1u) * sizeof(fr_t)));
    CU(ctx, dw.alloc((size_t)std::max(q->n_wires, 1u) * sizeof(fr_t)));
    CU(ctx, cudaEventRecord(ctx->ev[0], s));
    int rc = upload_canonical(ctx, dw.as<fr_t>(), w, q->n_wires);  // rejects witness values >= r
    if (rc) return rc;
    ++launches;
    CU(ctx, cudaEventRecord(ctx->ev[1], s));
    for (int k = 0; k < 3; ++k) {
        CU(ctx, launch_poly_combine(ctx->field, q->d_polys[k], dw.as<fr_t>(), q->n_wires, q->len, ev[k].as<fr_t>(), s));
        ++launches;
        if (!fr_is_zero(d[k])) {
            CU(ctx, launch_axpy1(ctx->field, ev[k].as<fr_t>(), q->d_target, d[k], q->n_target, s));
            ++launches;
        }
    }
    // p = a * b - c through a product on M >= 2 la - 1 points
    if (log_m > 0) {
        NttPlan *fwd = nullptr, *inv = nullptr;
        if ((rc = get_plan(ctx, log_m, false, &fwd))) return rc;
        if ((rc = get_plan(ctx, log_m, true, &inv))) return rc;
        for (int k = 0; k < 3; ++k) CU(ctx, ntt_run(fwd, ev[k].as<fr_t>(), scratch.as<fr_t>(), 1, s, &launches));
        CU(ctx, launch_quotient_pointwise(ctx->field, ev[0].as<fr_t>(), ev[1].as<fr_t>(), ev[2].as<fr_t>(), pbuf.as<fr_t>(), M,
                                          fr_one<P>(), s));
        ++launches;
        CU(ctx, ntt_run(inv, pbuf.as<fr_t>(), scratch.as<fr_t>(), 1, s, &launches));
    } else {
        CU(ctx, launch_quotient_pointwise(ctx->field, ev[0].as<fr_t>(), ev[1].as<fr_t>(), ev[2].as<fr_t>(), pbuf.as<fr_t>(), 1,
                                          fr_one<P>(), s));
        ++launches;
    }
    // (h, rem) = p divMod T (src/QAP.hs:327); valid <=> rem == 0
    uint64_t rem_len = len_p;
    if (h_len) {
        CU(ctx, launch_poly_divmod(ctx->field, pbuf.as<fr_t>(), (uint32_t)len_p, q->d_target, n, q->lc_inv, q->monic,
                                   hbuf.as<fr_t>(), s, &launches));
        rem_len = n;
    }
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), s));
    CU(ctx, launch_any_nonzero(pbuf.as<fr_t>(), rem_len, ctx->d_flag, s));
    if (rem_len) ++launches;
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaEventRecord(ctx->ev[2], s));
    if (h_out && h_len) {
        if ((rc = download_canonical(ctx, hbuf.as<fr_t>(), h_len, h_out))) return rc;
        ++launches;
    }
    CU(ctx, cudaEventRecord(ctx->ev[3], s));
    CU(ctx, cudaStreamSynchronize(s));
    ctx->launches += launches;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), elapsed(ctx->ev[1], ctx->ev[2]), elapsed(ctx->ev[2], ctx->ev[3]),
                             launches, 0};
    if (divisible) *divisible = *ctx->h_flag ? 0 : 1;
    return ACG_OK;
}

extern "C" {

void acg_qap_free(acg_qap* q) {
    if (!q) return;
    if (q->ctx) cudaSetDevice(q->ctx->device);
    for (auto& p : q->d_polys) cudaFree(p);
    cudaFree(q->d_target);
    delete q;
}

int acg_qap_upload(acg_ctx* ctx, const uint64_t* left, const uint64_t* right, const uint64_t* out, uint32_t n_wires,
                   uint32_t len, const uint64_t* target, uint32_t n_target, acg_qap** out_q) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!out_q || !target || n_target == 0 || (n_wires && len && (!left || !right || !out)))
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_qap_upload: bad argument");
    *out_q = nullptr;
    while (n_target && (target[4ull * (n_target - 1)] | target[4ull * (n_target - 1) + 1] | target[4ull * (n_target - 1) + 2] |
                        target[4ull * (n_target - 1) + 3]) == 0)
        --n_target;  // VPoly normal form: no trailing zero coefficients
    if (n_target == 0) return fail(ctx, ACG_ERR_BAD_ARG, "acg_qap_upload: the target polynomial is zero");
    if (len == 0 || n_wires == 0) len = n_wires = 0;
    acg_qap* q = new (std::nothrow) acg_qap();
    if (!q) return ACG_ERR_OOM;
    q->ctx = ctx;
    q->n_wires = n_wires;
    q->len = len;
    q->n_target = n_target;
    struct Guard {
        acg_qap* p;
        ~Guard() {
            if (p) acg_qap_free(p);
        }
    } guard{q};
    const uint64_t* src[3] = {left, right, out};
    const uint64_t cnt = (uint64_t)n_wires * len;
    for (int k = 0; k < 3; ++k) {
        CU(ctx, cudaMalloc(&q->d_polys[k], std::max<uint64_t>(cnt, 1) * sizeof(fr_t)));
        if ((rc = upload_canonical(ctx, q->d_polys[k], src[k], cnt))) return rc;
    }
    CU(ctx, cudaMalloc(&q->d_target, (size_t)n_target * sizeof(fr_t)));
    if ((rc = upload_canonical(ctx, q->d_target, target, n_target))) return rc;
    ctx->launches += 4;
    rc = with_field(ctx->field, [&](auto p) {
        using P = decltype(p);
        const fr_t lc = fr_to_mont<P>(fr_from_limbs(target + 4ull * (n_target - 1)));
        q->monic = fr_is_one<P>(lc);
        q->lc_inv = fr_inv<P>(lc);
        return (int)ACG_OK;
    });
    if (rc) return rc;
    guard.p = nullptr;
    *out_q = q;
    return ACG_OK;
    ACG_CATCH(ctx)
}

uint32_t acg_qap_quotient_len(const acg_qap* q) {
    if (!q) return 0;
    const uint64_t la = std::max(q->len, q->n_target);
    const uint64_t len_p = 2 * la - 1, n = q->n_target - 1u;
    return len_p > n ? (uint32_t)(len_p - n) : 0u;
}

int acg_qap_verify(acg_ctx* ctx, const acg_qap* q, const uint64_t* w, const uint64_t* delta, uint64_t* h,
                   uint32_t h_capacity, uint32_t* h_len, int* divisible) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!q || q->ctx != ctx || (q->n_wires && !w)) return fail(ctx, ACG_ERR_BAD_ARG, "acg_qap_verify: bad argument");
    return with_field(ctx->field, [&](auto p) {
        using P = decltype(p);
        return qap_verify_impl<P>(ctx, q, w, delta, h, h_capacity, h_len, divisible);
    });
    ACG_CATCH(ctx)
}

int acg_fft_target(acg_ctx* ctx, uint32_t n_roots, uint64_t* out) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!out) return fail(ctx, ACG_ERR_BAD_ARG, "acg_fft_target: null output");
    uint32_t log_n = 0;
    while ((1ull << log_n) < n_roots) ++log_n;
    if ((int)log_n > two_adicity(ctx->field)) return fail(ctx, ACG_ERR_UNSUPPORTED, "acg_fft_target: too many roots");
    std::memset(out, 0, ((size_t)n_roots + 1) * 32);
    if (n_roots == 0) {  // the empty product
        out[0] = 1;
        return ACG_OK;
    }
    if (n_roots == (1u << log_n)) {  // the whole domain: X^N - 1 (N = 1: X - 1)
        uint64_t modulus[4];
        acg_field_constants(ctx->field, modulus, nullptr, nullptr, nullptr, nullptr);
        std::memcpy(out, modulus, 32);
        out[0] -= 1;  // r - 1 (r is odd)
        out[4ull * n_roots] = 1;
        return ACG_OK;
    }
    if (n_roots > 4096) return fail(ctx, ACG_ERR_UNSUPPORTED, "acg_fft_target: a partial domain is limited to 4096 roots");
    return with_field(ctx->field, [&](auto p) -> int {
        using P = decltype(p);
        DevBuf xs, tg, scratch;
        CU(ctx, xs.alloc((size_t)n_roots * sizeof(fr_t)));
        CU(ctx, tg.alloc(((size_t)n_roots + 1) * sizeof(fr_t)));
        CU(ctx, scratch.alloc((3 * (size_t)n_roots + 2) * sizeof(fr_t)));
        uint32_t launches = 1;
        CU(ctx, launch_fill_powers(ctx->field, xs.as<fr_t>(), n_roots, host_root_of_unity<P>(log_n), 0, ctx->stream));
        CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
        CU(ctx, launch_lagrange(ctx->field, xs.as<fr_t>(), nullptr, n_roots, 0, nullptr, tg.as<fr_t>(), scratch.as<fr_t>(),
                                ctx->d_flag, ctx->stream, &launches));
        int rc2 = download_canonical(ctx, tg.as<fr_t>(), (uint64_t)n_roots + 1, out);
        if (rc2) return rc2;
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->launches += launches + 1;
        return ACG_OK;
    });
    ACG_CATCH(ctx)
}

// --------------------------------------------------------------------------------------------------
// linear constraints (Bulletproofs backend): checkLinearConstraint, src/Circuit/Bulletproofs.hs:329-338
// --------------------------------------------------------------------------------------------------
int acg_linear_constraints_check(acg_ctx* ctx, int modulus_id, uint32_t n_constraints, uint32_t n_lhs_vars,
                                 uint32_t n_rhs_vars, const acg_csr* lhs, const acg_csr* rhs, const uint64_t* constants,
                                 const uint64_t* x, const uint64_t* v, uint64_t* n_violations, uint64_t* first_bad) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (modulus_id < 0 || modulus_id > ACG_FIELD_SECP256K1_FN || !lhs || !rhs || (n_constraints && !constants) ||
        (n_lhs_vars && !x) || (n_rhs_vars && !v))
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_linear_constraints_check: bad argument");
    const acg_csr* src[2] = {lhs, rhs};
    const uint32_t n_vars[2] = {n_lhs_vars, n_rhs_vars};
    for (int k = 0; k < 2; ++k) {
        const acg_csr* M = src[k];
        if (!M->rowptr || (M->nnz && (!M->col || !M->val)) || M->nnz > 0xFFFFFFF0ull || M->rowptr[0] != 0 ||
            M->rowptr[n_constraints] != M->nnz)
            return fail(ctx, ACG_ERR_BAD_ARG, "acg_linear_constraints_check: malformed CSR");
        for (uint32_t r = 0; r < n_constraints; ++r)
            if (M->rowptr[r] > M->rowptr[r + 1]) return fail(ctx, ACG_ERR_BAD_ARG, "acg_linear_constraints_check: rowptr not monotone");
        for (uint64_t e = 0; e < M->nnz; ++e)
            if (M->col[e] >= n_vars[k]) return fail(ctx, ACG_ERR_BAD_ARG, "acg_linear_constraints_check: column index out of range");
    }
    cudaStream_t s = ctx->stream;
    DevBuf d_rp[2], d_col[2], d_val[2], d_x[2], d_c, d_res;
    const uint64_t* host_x[2] = {x, v};
    auto h2d = [&](DevBuf& d, const void* p, size_t bytes) -> cudaError_t {
        cudaError_t e = d.alloc(bytes);
        if (e != cudaSuccess || bytes == 0) return e;
        return cudaMemcpyAsync(d.p, p, bytes, cudaMemcpyHostToDevice, s);
    };
    CU(ctx, cudaEventRecord(ctx->ev[0], s));
    for (int k = 0; k < 2; ++k) {
        CU(ctx, h2d(d_rp[k], src[k]->rowptr, ((size_t)n_constraints + 1) * sizeof(uint32_t)));
        CU(ctx, h2d(d_col[k], src[k]->col, (size_t)src[k]->nnz * sizeof(uint32_t)));
        CU(ctx, h2d(d_val[k], src[k]->val, (size_t)src[k]->nnz * 32));
        CU(ctx, h2d(d_x[k], host_x[k], (size_t)n_vars[k] * 32));
    }
    CU(ctx, h2d(d_c, constants, (size_t)n_constraints * 32));
    const unsigned long long init[2] = {0ull, ~0ull};
    CU(ctx, h2d(d_res, init, sizeof init));
    CU(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), s));
    CU(ctx, cudaEventRecord(ctx->ev[1], s));
    uint32_t launches = 0;
    for (int k = 0; k < 2; ++k) {
        CU(ctx, launch_lin_to_mont(modulus_id, d_x[k].as<uint64_t>(), n_vars[k], ctx->d_flag, s));
        launches += n_vars[k] ? 1 : 0;
    }
    CU(ctx, launch_linear_constraints(modulus_id, d_rp[0].as<uint32_t>(), d_col[0].as<uint32_t>(), d_val[0].as<uint64_t>(),
                                      d_rp[1].as<uint32_t>(), d_col[1].as<uint32_t>(), d_val[1].as<uint64_t>(),
                                      d_c.as<uint64_t>(), d_x[0].as<uint64_t>(), d_x[1].as<uint64_t>(), n_constraints,
                                      d_res.as<unsigned long long>(), ctx->d_flag, s));
    launches += n_constraints ? 1 : 0;
    CU(ctx, cudaEventRecord(ctx->ev[2], s));
    CU(ctx, cudaMemcpyAsync(ctx->h_result, d_res.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaEventRecord(ctx->ev[3], s));
    CU(ctx, cudaStreamSynchronize(s));
    ctx->launches += launches;
    ctx->timing = acg_timing{elapsed(ctx->ev[0], ctx->ev[1]), elapsed(ctx->ev[1], ctx->ev[2]), elapsed(ctx->ev[2], ctx->ev[3]),
                             launches, 0};
    if (*ctx->h_flag) return fail(ctx, ACG_ERR_NON_CANONICAL, "acg_linear_constraints_check: element >= modulus");
    if (n_violations) *n_violations = ctx->h_result[0];
    if (first_bad) *first_bad = ctx->h_result[1];
    return ACG_OK;
    ACG_CATCH(ctx)
}

// --------------------------------------------------------------------------------------------------
// multi-GPU: result all-reduce over peer memory
// --------------------------------------------------------------------------------------------------
int acg_peer_create(acg_ctx* ctx, uint32_t world, uint32_t rank, acg_peer** out, uint8_t* handle_out) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!out || !handle_out || world == 0 || world > kMaxPeers || rank >= world)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_peer_create: bad argument (1 <= world <= 8)");
    static_assert(sizeof(cudaIpcMemHandle_t) == ACG_PEER_HANDLE_BYTES, "IPC handle size");
    *out = nullptr;
    acg_peer* p = new (std::nothrow) acg_peer();
    if (!p) return ACG_ERR_OOM;
    p->ctx = ctx;
    p->world = world;
    p->rank = rank;
    cudaError_t e = cudaMalloc(&p->local, kPeerBufferBytes);
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, kPeerBufferBytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) {
        cudaFree(p->local);
        delete p;
        return fail_cuda(ctx, e, "acg_peer_create");
    }
    std::memcpy(handle_out, &h, sizeof h);
    *out = p;
    return ACG_OK;
    ACG_CATCH(ctx)
}

int acg_peer_connect(acg_ctx* ctx, acg_peer* p, const uint8_t* handles) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!p || p->ctx != ctx || !handles || p->connected) return fail(ctx, ACG_ERR_BAD_ARG, "acg_peer_connect: bad argument");
    unsigned long long* bases[kMaxPeers] = {};
    for (uint32_t r = 0; r < p->world; ++r) {
        if (r == p->rank) {
            bases[r] = p->local;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * sizeof h, sizeof h);
        CU(ctx, cudaIpcOpenMemHandle(&p->mapped[r], h, cudaIpcMemLazyEnablePeerAccess));
        bases[r] = static_cast<unsigned long long*>(p->mapped[r]);
    }
    if (!p->d_table) CU(ctx, cudaMalloc(&p->d_table, sizeof bases));
    CU(ctx, cudaMemcpy(p->d_table, bases, sizeof bases, cudaMemcpyHostToDevice));
    p->slots.base = p->d_table;
    p->slots.own = p->local;
    p->slots.world = p->world;
    p->slots.rank = p->rank;
    p->connected = true;
    return ACG_OK;
    ACG_CATCH(ctx)
}

void acg_peer_free(acg_peer* p) {
    if (!p) return;
    if (p->ctx) cudaSetDevice(p->ctx->device);
    for (uint32_t r = 0; r < kMaxPeers; ++r)
        if (p->mapped[r]) cudaIpcCloseMemHandle(p->mapped[r]);
    cudaFree(p->local);
    cudaFree(p->d_table);
    delete p;
}

int acg_r1cs_check_async_allreduce(acg_ctx* ctx, const acg_r1cs* m, const acg_vec* w, acg_peer* peer,
                                   uint64_t* d_result, void* stream) {
    ACG_TRY
    int rc = activate(ctx);
    if (rc) return rc;
    if (!m || !w || !d_result || !peer || m->ctx != ctx || w->ctx != ctx || peer->ctx != ctx || w->n != m->n_cols ||
        !peer->connected)
        return fail(ctx, ACG_ERR_BAD_ARG, "acg_r1cs_check_async_allreduce: bad argument");
    uint32_t launches = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rc = enqueue_check(ctx, m, w, reinterpret_cast<unsigned long long*>(d_result), nullptr, nullptr, nullptr, s,
                       &launches, peer);
    if (rc) return rc;
    ctx->launches += launches;
    ctx->timing.kernel_launches = launches;
    return ACG_OK;
    ACG_CATCH(ctx)
}

}  // extern "C"
