// Synthetic circuit family S(n, seed, field) of SURVEY.md section 8(d): the workload bench.py and the
// parity tests run.  n Mul gates over 1024 input wires; each side of a gate is an optional constant
// (p = 1/4) plus two wire terms, a term being "near" (one of the previous 64 witness indices, p = 1/2)
// or "far" (any earlier index) with coefficient 1 (p = 1/2), -1 (p = 1/4) or a uniform field element
// (p = 1/4).  Expressible as a reference ArithCircuit (Add / ScalarMul / ConstGate / Var), which
// acg_synth_circuit_words emits; acg_synth_r1cs produces the lowered CSR system and the honest
// witness directly (same values, no per-gate allocation) so 2^24-gate instances build in seconds.
// The draw order is fixed and mirrored bit for bit by the test oracle's independent generator.
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../../include/acg.h"
#include "circuit.hpp"

// No C++ exception crosses the C ABI (include/acg.h): std::bad_alloc -> ACG_ERR_OOM, anything else -> ACG_ERR_INTERNAL.
#define ACG_TRY try {
#define ACG_CATCH()                                         \
    }                                                       \
    catch (const std::bad_alloc&) { return ACG_ERR_OOM; }   \
    catch (...) { return ACG_ERR_INTERNAL; }

using namespace acg;
using namespace acg::host;

namespace {

constexpr uint32_t kInputs = 1024;
constexpr uint32_t kNear = 64;

struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint64_t below(uint64_t n) { return next() % n; }
    template <class P>
    El field() {  // canonical, uniform on [0, r) by rejection
        const int top_bits = P::BITS - 192;
        for (;;) {
            El v;
            v.v[0] = next();
            v.v[1] = next();
            v.v[2] = next();
            v.v[3] = next() & ((1ull << top_bits) - 1ull);
            if (!Fr<P>::geq_mod(v)) return v;
        }
    }
};

struct Term {
    uint32_t idx;   // witness index, 0 = constant column
    int kind;       // 0/1: coefficient 1, 2: -1, 3: general
    El coef;        // canonical (kind 3, or the constant itself)
};

// one side of a gate, in draw order: [constant?] term term
template <class P>
int draw_side(SplitMix64& rng, uint32_t avail, bool dense, Term out[3]) {
    int n = 0;
    if ((rng.next() & 3ull) == 0ull) {
        out[n].idx = 0;
        out[n].kind = 3;
        out[n].coef = rng.field<P>();
        ++n;
    }
    for (int t = 0; t < 2; ++t) {
        const bool near = (rng.next() & 1ull) != 0;
        uint32_t idx;
        if (near) {
            const uint32_t lo = avail > kNear ? avail - kNear : 1u;
            idx = lo + (uint32_t)rng.below(avail - lo);
        } else {
            idx = 1u + (uint32_t)rng.below(avail - 1u);
        }
        const int kind = dense ? 3 : (int)(rng.next() & 3ull);
        out[n].idx = idx;
        out[n].kind = kind;
        if (kind == 3) out[n].coef = rng.field<P>();
        ++n;
    }
    return n;
}

template <class P>
El term_coef_mont(const Term& t) {
    if (t.kind <= 1) return Fr<P>::one();
    if (t.kind == 2) return Fr<P>::minus_one();
    return Fr<P>::to_mont(t.coef);
}

template <class P>
int synth_r1cs_impl(uint32_t n, uint64_t seed, bool dense, uint32_t rb, uint32_t re, acg_r1cs_host* m, uint64_t* w) {
    SplitMix64 rng{seed};
    const uint32_t n_cols = 1 + kInputs + n;
    std::vector<El> wm((size_t)n_cols);  // Montgomery witness
    wm[0] = Fr<P>::one();
    for (uint32_t i = 0; i < kInputs; ++i) wm[1 + i] = Fr<P>::to_mont(rng.field<P>());
    for (int k = 0; k < 3; ++k) {
        m->rowptr[k].reserve((size_t)(re - rb) + 1);
        m->rowptr[k].push_back(0);
    }
    const size_t nl = re - rb;
    m->col[0].reserve(nl * 9 / 4 + 16);
    m->col[1].reserve(nl * 9 / 4 + 16);
    m->col[2].reserve(nl);
    m->val[0].reserve(nl * 9 + 64);
    m->val[1].reserve(nl * 9 + 64);
    m->val[2].reserve(nl * 4);
    for (uint32_t g = 0; g < n; ++g) {
        const uint32_t avail = 1 + kInputs + g;
        const bool keep = g >= rb && g < re;
        El side_val[2];
        for (int side = 0; side < 2; ++side) {
            Term t[3];
            const int nt = draw_side<P>(rng, avail, dense, t);
            // affineCircuitToAffineMap: merge duplicate wires by addition (src/Circuit/Affine.hs:98)
            uint32_t cols[3];
            El coefs[3];
            int nu = 0;
            for (int i = 0; i < nt; ++i) {
                const El c = term_coef_mont<P>(t[i]);
                int j = 0;
                for (; j < nu; ++j)
                    if (cols[j] == t[i].idx) break;
                if (j < nu) {
                    coefs[j] = Fr<P>::add(coefs[j], c);
                } else {
                    cols[nu] = t[i].idx;
                    coefs[nu] = c;
                    ++nu;
                }
            }
            // sort by column (constant column 0 first), drop zero coefficients, evaluate
            for (int i = 1; i < nu; ++i)
                for (int j = i; j > 0 && cols[j] < cols[j - 1]; --j) {
                    std::swap(cols[j], cols[j - 1]);
                    std::swap(coefs[j], coefs[j - 1]);
                }
            El acc = Fr<P>::zero();
            for (int i = 0; i < nu; ++i) {
                if (coefs[i].is_zero()) continue;
                acc = Fr<P>::add(acc, Fr<P>::mul(coefs[i], wm[cols[i]]));
                if (!keep) continue;
                m->col[side].push_back(cols[i]);
                const El c = Fr<P>::from_mont(coefs[i]);
                m->val[side].insert(m->val[side].end(), c.v, c.v + 4);
            }
            if (keep) m->rowptr[side].push_back((uint32_t)m->col[side].size());
            side_val[side] = acc;
        }
        const uint32_t out_col = 1 + kInputs + g;
        wm[out_col] = Fr<P>::mul(side_val[0], side_val[1]);
        if (keep) {
            m->col[2].push_back(out_col);
            const uint64_t one[4] = {1, 0, 0, 0};
            m->val[2].insert(m->val[2].end(), one, one + 4);
            m->rowptr[2].push_back((uint32_t)m->col[2].size());
        }
    }
    for (uint32_t i = 0; i < n_cols; ++i) {
        const El c = Fr<P>::from_mont(wm[i]);
        std::memcpy(w + 4ull * i, c.v, 32);
    }
    m->n_rows = re - rb;
    m->n_cols = n_cols;
    m->n_in = kInputs;
    m->n_mid = n ? n - 1 : 0;
    m->n_out = n ? 1 : 0;
    m->roots.resize(4ull * (re - rb));
    for (uint32_t r = rb; r < re; ++r) m->roots[4ull * (r - rb)] = r;
    return ACG_OK;
}

template <class P>
void emit_side(SplitMix64& rng, uint32_t avail, bool dense, std::vector<uint64_t>& words) {
    Term t[3];
    const int nt = draw_side<P>(rng, avail, dense, t);
    const size_t len_pos = words.size();
    words.push_back(0);
    for (int i = 0; i < nt; ++i) {
        if (t[i].idx == 0) {  // ConstGate c
            words.push_back(1);
            words.insert(words.end(), t[i].coef.v, t[i].coef.v + 4);
        } else {
            const uint32_t idx = t[i].idx;
            const uint64_t wire = idx <= kInputs ? ACG_WIRE(ACG_WIRE_INPUT, idx - 1)
                                                 : ACG_WIRE(ACG_WIRE_INTERMEDIATE, idx - 1 - kInputs);
            words.push_back(0);
            words.push_back(wire);
            if (t[i].kind >= 2) {  // ScalarMul c (Var w)
                El c = t[i].coef;
                if (t[i].kind == 2) {
                    c = Fr<P>::modulus();
                    c.v[0] -= 1;  // r - 1 (r is odd)
                }
                words.push_back(3);
                words.insert(words.end(), c.v, c.v + 4);
            }
        }
        if (i > 0) words.push_back(2);  // Add (left fold)
    }
    words[len_pos] = words.size() - len_pos - 1;
}

template <class P>
int synth_words_impl(uint32_t n, uint64_t seed, bool dense, std::vector<uint64_t>& words, std::vector<uint64_t>& in_vals) {
    SplitMix64 rng{seed};
    for (uint32_t i = 0; i < kInputs; ++i) {
        const El v = rng.field<P>();
        in_vals.insert(in_vals.end(), v.v, v.v + 4);
    }
    for (uint32_t g = 0; g < n; ++g) {
        const uint32_t avail = 1 + kInputs + g;
        words.push_back(1);  // Mul
        words.push_back(g == n - 1 ? ACG_WIRE(ACG_WIRE_OUTPUT, 0) : ACG_WIRE(ACG_WIRE_INTERMEDIATE, g));
        emit_side<P>(rng, avail, dense, words);
        emit_side<P>(rng, avail, dense, words);
    }
    return ACG_OK;
}

}  // namespace

extern "C" {

int acg_synth_r1cs(int field_id, uint32_t n, uint64_t seed, int dense, acg_r1cs_host** out_m, uint64_t** out_w) {
    ACG_TRY
    return acg_synth_r1cs_rows(field_id, n, seed, dense, 0, n, out_m, out_w);
    ACG_CATCH()
}

int acg_synth_r1cs_rows(int field_id, uint32_t n, uint64_t seed, int dense, uint32_t row_begin, uint32_t row_end,
                        acg_r1cs_host** out_m, uint64_t** out_w) {
    ACG_TRY
    if (!out_m || !out_w || n == 0 || n > 0xF0000000u - kInputs || row_begin > row_end || row_end > n)
        return ACG_ERR_BAD_ARG;
    *out_m = nullptr;
    *out_w = nullptr;
    acg_r1cs_host* m = new (std::nothrow) acg_r1cs_host();
    uint64_t* w = static_cast<uint64_t*>(std::malloc((size_t)(1 + kInputs + (size_t)n) * 32));
    if (!m || !w) {
        delete m;
        std::free(w);
        return ACG_ERR_OOM;
    }
    m->field = field_id;
    int rc = ACG_ERR_BAD_ARG;
    try {
        if (field_id == ACG_FIELD_BN254_FR) rc = synth_r1cs_impl<Bn254Fr>(n, seed, dense != 0, row_begin, row_end, m, w);
        if (field_id == ACG_FIELD_BLS12_381_FR) rc = synth_r1cs_impl<Bls12381Fr>(n, seed, dense != 0, row_begin, row_end, m, w);
    } catch (const std::bad_alloc&) {
        rc = ACG_ERR_OOM;
    }
    if (rc != ACG_OK) {
        delete m;
        std::free(w);
        return rc;
    }
    *out_m = m;
    *out_w = w;
    return ACG_OK;
    ACG_CATCH()
}

int acg_synth_circuit_words(int field_id, uint32_t n, uint64_t seed, int dense, uint64_t** out_words,
                            uint64_t* out_n_words, uint32_t** out_input_ix, uint64_t** out_input_vals,
                            uint32_t* out_n_inputs) {
    ACG_TRY
    if (!out_words || !out_n_words || !out_input_ix || !out_input_vals || !out_n_inputs || n == 0)
        return ACG_ERR_BAD_ARG;
    std::vector<uint64_t> words, in_vals;
    int rc = ACG_ERR_BAD_ARG;
    try {
        if (field_id == ACG_FIELD_BN254_FR) rc = synth_words_impl<Bn254Fr>(n, seed, dense != 0, words, in_vals);
        if (field_id == ACG_FIELD_BLS12_381_FR) rc = synth_words_impl<Bls12381Fr>(n, seed, dense != 0, words, in_vals);
    } catch (const std::bad_alloc&) {
        rc = ACG_ERR_OOM;
    }
    if (rc != ACG_OK) return rc;
    uint64_t* ww = static_cast<uint64_t*>(std::malloc(words.size() * 8 + 8));
    uint64_t* iv = static_cast<uint64_t*>(std::malloc(in_vals.size() * 8 + 8));
    uint32_t* ix = static_cast<uint32_t*>(std::malloc(kInputs * 4));
    if (!ww || !iv || !ix) {
        std::free(ww);
        std::free(iv);
        std::free(ix);
        return ACG_ERR_OOM;
    }
    std::memcpy(ww, words.data(), words.size() * 8);
    std::memcpy(iv, in_vals.data(), in_vals.size() * 8);
    for (uint32_t i = 0; i < kInputs; ++i) ix[i] = i;
    *out_words = ww;
    *out_n_words = words.size();
    *out_input_ix = ix;
    *out_input_vals = iv;
    *out_n_inputs = kInputs;
    return ACG_OK;
    ACG_CATCH()
}

}  // extern "C"
