// Host side of the path, C++ mirror of the reference's Haskell host logic (no GPU involved):
//   ArithCircuit / Gate / AffineCircuit IR      src/Circuit/Arithmetic.hs:32-59,149-150, Affine.hs:26-31
//   validArithCircuit                            src/Circuit/Arithmetic.hs:158-185
//   generateAssignment = foldl' evalGate         src/QAP.hs:597-603, Arithmetic.hs:106-145,221-235
//   gateToGenQAP / arithCircuitToGenQAP as CSR   src/QAP.hs:366-474, 530-539
//   qapSetToMap as a dense vector                src/QAP.hs:605-620
// The circuit crosses the ABI as the flat word stream documented in include/acg.h.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <thread>
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

enum : uint64_t { TOK_VAR = 0, TOK_CONST = 1, TOK_ADD = 2, TOK_SCALAR = 3 };
enum : uint64_t { G_MUL = 1, G_EQUAL = 2, G_SPLIT = 3 };

inline bool wire_ok(uint64_t w) { return (w >> 62) <= ACG_WIRE_OUTPUT && (w & 0x3FFFFFFFFFFFFFFFull) < 0x7FFFFFFFull; }

template <class P>
bool parse_affine(const uint64_t* w, uint64_t n, std::vector<AffTok>& out) {
    int64_t depth = 0;
    uint64_t i = 0;
    while (i < n) {
        AffTok t{};
        t.op = (uint8_t)w[i];
        switch (w[i]) {
            case TOK_VAR:
                if (i + 1 >= n || !wire_ok(w[i + 1])) return false;
                t.wire = w[i + 1];
                i += 2;
                ++depth;
                break;
            case TOK_CONST:
            case TOK_SCALAR: {
                if (i + 4 >= n) return false;
                El c{{w[i + 1], w[i + 2], w[i + 3], w[i + 4]}};
                if (Fr<P>::geq_mod(c)) return false;
                t.val = Fr<P>::to_mont(c);
                if (w[i] == TOK_CONST) {
                    ++depth;
                } else if (depth < 1) {
                    return false;
                }
                i += 5;
                break;
            }
            case TOK_ADD:
                if (depth < 2) return false;
                --depth;
                i += 1;
                break;
            default:
                return false;
        }
        out.push_back(t);
    }
    return depth == 1;
}

template <class P>
int parse_impl(const uint64_t* w, uint64_t n, acg_circuit* c) {
    uint64_t i = 0;
    while (i < n) {
        GateH g{};
        g.kind = (uint8_t)w[i];
        switch (w[i]) {
            case G_MUL: {
                if (i + 2 >= n || !wire_ok(w[i + 1])) return ACG_ERR_BAD_ARG;
                g.w2 = w[i + 1];
                uint64_t nl = w[i + 2];
                i += 3;
                if (nl > n - i) return ACG_ERR_BAD_ARG;
                if (!parse_affine<P>(w + i, nl, g.l)) return ACG_ERR_BAD_ARG;
                i += nl;
                if (i >= n) return ACG_ERR_BAD_ARG;
                uint64_t nr = w[i];
                i += 1;
                if (nr > n - i) return ACG_ERR_BAD_ARG;
                if (!parse_affine<P>(w + i, nr, g.r)) return ACG_ERR_BAD_ARG;
                i += nr;
                c->n_roots += 1;
                break;
            }
            case G_EQUAL:
                if (i + 3 >= n || !wire_ok(w[i + 1]) || !wire_ok(w[i + 2]) || !wire_ok(w[i + 3])) return ACG_ERR_BAD_ARG;
                g.w0 = w[i + 1];
                g.w1 = w[i + 2];
                g.w2 = w[i + 3];
                i += 4;
                c->n_roots += 2;
                break;
            case G_SPLIT: {
                if (i + 2 >= n || !wire_ok(w[i + 1])) return ACG_ERR_BAD_ARG;
                g.w0 = w[i + 1];
                uint64_t no = w[i + 2];
                i += 3;
                if (no > n - i) return ACG_ERR_BAD_ARG;
                for (uint64_t k = 0; k < no; ++k) {
                    if (!wire_ok(w[i + k])) return ACG_ERR_BAD_ARG;
                    g.outs.push_back(w[i + k]);
                }
                i += no;
                c->n_roots += 1 + no;
                break;
            }
            default:
                return ACG_ERR_BAD_ARG;
        }
        c->gates.push_back(std::move(g));
    }
    return ACG_OK;
}

inline uint32_t wkind(uint64_t w) { return (uint32_t)(w >> 62); }
inline uint32_t wix(uint64_t w) { return (uint32_t)(w & 0xFFFFFFFFull); }

// evalAffineCircuit, src/Circuit/Affine.hs:73-86 (failed lookups are 0)
template <class P>
El eval_affine(const std::vector<AffTok>& toks, const acg_assignment* a, std::vector<El>& stack) {
    stack.clear();
    for (const AffTok& t : toks) {
        switch (t.op) {
            case TOK_VAR: {
                const WireMap& m = a->part[wkind(t.wire)];
                const uint32_t ix = wix(t.wire);
                stack.push_back(ix < m.val.size() && m.present[ix] ? m.val[ix] : Fr<P>::zero());
                break;
            }
            case TOK_CONST: stack.push_back(t.val); break;
            case TOK_ADD: {
                El r = stack.back();
                stack.pop_back();
                stack.back() = Fr<P>::add(stack.back(), r);
                break;
            }
            default: stack.back() = Fr<P>::mul(stack.back(), t.val); break;
        }
    }
    return stack.back();
}

void wm_set(WireMap& m, uint32_t ix, const El& v) {
    if (ix >= m.val.size()) {
        m.val.resize((size_t)ix + 1, El{{0, 0, 0, 0}});
        m.present.resize((size_t)ix + 1, 0);
    }
    m.val[ix] = v;
    m.present[ix] = 1;
}
bool wm_get(const WireMap& m, uint32_t ix, El& out) {
    if (ix >= m.val.size() || !m.present[ix]) return false;
    out = m.val[ix];
    return true;
}

// evalGate, src/Circuit/Arithmetic.hs:106-145
template <class P>
int eval_impl(const acg_circuit* c, acg_assignment* a) {
    std::vector<El> stack;
    for (const GateH& g : c->gates) {
        if (g.kind == G_MUL) {
            const El l = eval_affine<P>(g.l, a, stack);
            const El r = eval_affine<P>(g.r, a, stack);
            wm_set(a->part[wkind(g.w2)], wix(g.w2), Fr<P>::mul(l, r));
        } else if (g.kind == G_EQUAL) {
            El inp;
            if (!wm_get(a->part[wkind(g.w0)], wix(g.w0), inp)) return ACG_ERR_BAD_ARG;  // "the impossible happened"
            const bool z = inp.is_zero();
            wm_set(a->part[wkind(g.w1)], wix(g.w1), z ? Fr<P>::zero() : Fr<P>::inv(inp));
            wm_set(a->part[wkind(g.w2)], wix(g.w2), z ? Fr<P>::zero() : Fr<P>::one());
        } else {
            El inp;
            if (!wm_get(a->part[wkind(g.w0)], wix(g.w0), inp)) return ACG_ERR_BAD_ARG;
            const El canon = Fr<P>::from_mont(inp);  // testBit (fromP inp) ix
            for (size_t ix = 0; ix < g.outs.size(); ++ix) {
                const bool bit = ix < 256 && ((canon.v[ix >> 6] >> (ix & 63)) & 1);
                wm_set(a->part[wkind(g.outs[ix])], wix(g.outs[ix]), bit ? Fr<P>::one() : Fr<P>::zero());
            }
        }
    }
    return ACG_OK;
}

struct Layout {
    uint32_t n_in, n_mid, n_out;
    uint32_t col(uint64_t w) const {
        const uint32_t ix = wix(w);
        switch (wkind(w)) {
            case ACG_WIRE_INPUT: return 1 + ix;
            case ACG_WIRE_INTERMEDIATE: return 1 + n_in + ix;
            default: return 1 + n_in + n_mid + ix;
        }
    }
    bool covers(uint64_t w) const {
        const uint32_t ix = wix(w);
        switch (wkind(w)) {
            case ACG_WIRE_INPUT: return ix < n_in;
            case ACG_WIRE_INTERMEDIATE: return ix < n_mid;
            default: return ix < n_out;
        }
    }
};

void note_wire(uint32_t dims[3], uint64_t w) { dims[wkind(w)] = std::max(dims[wkind(w)], wix(w) + 1); }

void circuit_dims(const acg_circuit* c, uint32_t dims[3]) {
    dims[0] = dims[1] = dims[2] = 0;
    for (const GateH& g : c->gates) {
        if (g.kind == G_MUL) {
            for (const AffTok& t : g.l)
                if (t.op == TOK_VAR) note_wire(dims, t.wire);
            for (const AffTok& t : g.r)
                if (t.op == TOK_VAR) note_wire(dims, t.wire);
            note_wire(dims, g.w2);
        } else if (g.kind == G_EQUAL) {
            note_wire(dims, g.w0);
            note_wire(dims, g.w1);
            note_wire(dims, g.w2);
        } else {
            note_wire(dims, g.w0);
            for (uint64_t o : g.outs) note_wire(dims, o);
        }
    }
}

// A sparse row under construction: (column, coefficient) with "later update wins" (updateAtWire).
typedef std::vector<std::pair<uint32_t, El>> Row;
void row_set(Row& r, uint32_t col, const El& v) {
    for (auto& e : r)
        if (e.first == col) {
            e.second = v;
            return;
        }
    r.emplace_back(col, v);
}

// affineCircuitToAffineMap, src/Circuit/Affine.hs:90-105: duplicates merge with +, scalars distribute.
template <class P>
void affine_to_row(const std::vector<AffTok>& toks, const Layout& lay, Row& out) {
    struct Part {
        El c;
        Row v;
    };
    std::vector<Part> st;
    for (const AffTok& t : toks) {
        switch (t.op) {
            case TOK_VAR: {
                Part p{Fr<P>::zero(), {}};
                p.v.emplace_back(lay.col(t.wire), Fr<P>::one());
                st.push_back(std::move(p));
                break;
            }
            case TOK_CONST: st.push_back(Part{t.val, {}}); break;
            case TOK_ADD: {
                Part r = std::move(st.back());
                st.pop_back();
                Part& l = st.back();
                l.c = Fr<P>::add(l.c, r.c);
                for (auto& e : r.v) {
                    bool found = false;
                    for (auto& f : l.v)
                        if (f.first == e.first) {
                            f.second = Fr<P>::add(f.second, e.second);
                            found = true;
                            break;
                        }
                    if (!found) l.v.push_back(e);
                }
                break;
            }
            default: {
                Part& p = st.back();
                p.c = Fr<P>::mul(t.val, p.c);
                for (auto& e : p.v) e.second = Fr<P>::mul(t.val, e.second);
                break;
            }
        }
    }
    out = std::move(st.back().v);
    row_set(out, 0, st.back().c);  // the constant column (constantQapSet (root, const), src/QAP.hs:375-376)
}

struct RowSink {
    std::vector<uint32_t> rowptr[3], col[3];
    std::vector<uint64_t> val[3];
    RowSink() {
        for (auto& r : rowptr) r.push_back(0);
    }
    template <class P>
    void push(int which, Row& r) {
        std::sort(r.begin(), r.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (auto& e : r) {
            if (e.second.is_zero()) continue;  // explicit zeros carry no information in the R1CS form
            col[which].push_back(e.first);
            const El c = Fr<P>::from_mont(e.second);
            val[which].insert(val[which].end(), c.v, c.v + 4);
        }
        rowptr[which].push_back((uint32_t)col[which].size());
    }
};

// gateToGenQAP, src/QAP.hs:366-474, emitted row by row for the gates [g0, g1)
template <class P>
int lower_range(const acg_circuit* c, const Layout& lay, size_t g0, size_t g1, RowSink& sink) {
    const El one = Fr<P>::one(), m1 = Fr<P>::minus_one(), zero = Fr<P>::zero();
    for (size_t gi = g0; gi < g1; ++gi) {
        const GateH& g = c->gates[gi];
        if (g.kind == G_MUL) {  // :371-395
            Row A, B, C;
            affine_to_row<P>(g.l, lay, A);
            affine_to_row<P>(g.r, lay, B);
            row_set(C, 0, zero);
            row_set(C, lay.col(g.w2), one);
            sink.push<P>(0, A);
            sink.push<P>(1, B);
            sink.push<P>(2, C);
        } else if (g.kind == G_EQUAL) {  // :396-442
            const uint32_t i = lay.col(g.w0), m = lay.col(g.w1), o = lay.col(g.w2);
            Row A, B, C;
            row_set(A, 0, zero); row_set(A, i, one);  row_set(A, m, zero); row_set(A, o, zero);
            row_set(B, 0, zero); row_set(B, i, zero); row_set(B, m, one);  row_set(B, o, zero);
            row_set(C, 0, zero); row_set(C, i, zero); row_set(C, m, zero); row_set(C, o, one);
            sink.push<P>(0, A); sink.push<P>(1, B); sink.push<P>(2, C);
            Row A1, B1, C1;
            row_set(A1, 0, one);  row_set(A1, i, zero); row_set(A1, m, zero); row_set(A1, o, m1);
            row_set(B1, 0, zero); row_set(B1, i, one);  row_set(B1, m, zero); row_set(B1, o, zero);
            row_set(C1, 0, zero); row_set(C1, i, zero); row_set(C1, m, zero); row_set(C1, o, zero);
            sink.push<P>(0, A1); sink.push<P>(1, B1); sink.push<P>(2, C1);
        } else {  // Split, :443-473
            const uint32_t inp = lay.col(g.w0);
            Row A, B, C;
            row_set(A, 0, zero);
            row_set(A, inp, zero);
            El pw = one;  // 2^ix
            for (uint64_t o : g.outs) {
                row_set(A, lay.col(o), pw);
                pw = Fr<P>::add(pw, pw);
            }
            row_set(B, 0, one);
            row_set(B, inp, zero);
            row_set(C, 0, zero);
            row_set(C, inp, one);
            sink.push<P>(0, A); sink.push<P>(1, B); sink.push<P>(2, C);
            for (uint64_t o : g.outs) {
                const uint32_t oc = lay.col(o);
                Row A2, B2, C2;
                row_set(A2, 0, zero); row_set(A2, oc, one);
                row_set(B2, 0, one);  row_set(B2, oc, m1);
                row_set(C2, 0, zero); row_set(C2, oc, zero);
                sink.push<P>(0, A2); sink.push<P>(1, B2); sink.push<P>(2, C2);
            }
        }
    }
    return ACG_OK;
}

// Worker threads of the host-side lowering: ACG_HOST_THREADS, else the hardware concurrency (at most 64).
unsigned host_threads() {
    if (const char* e = std::getenv("ACG_HOST_THREADS")) {
        const long v = std::strtol(e, nullptr, 10);
        if (v >= 1) return (unsigned)std::min<long>(v, 256);
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hc ? hc : 1u, 64u));
}

// The gates lower independently (a gate's rows depend on the gate and the layout only), so contiguous chunks of the
// gate list are lowered by worker threads into private sinks and concatenated in order: the result is identical to
// the sequential pass (at 2^24 gates the lowering, not the check, is what a caller waits for).
template <class P>
int lower_impl(const acg_circuit* c, const Layout& lay, RowSink& sink) {
    const size_t n = c->gates.size();
    const size_t T = std::min<size_t>(host_threads(), n / 2048 + 1);
    if (T <= 1) return lower_range<P>(c, lay, 0, n, sink);
    std::vector<RowSink> parts(T);
    std::vector<int> rcs(T, ACG_OK);
    std::vector<std::thread> workers;
    workers.reserve(T);
    bool spawn_failed = false;
    for (size_t t = 0; t < T && !spawn_failed; ++t) {
        try {  // nothing may escape a worker (std::terminate) or leave joinable threads behind
            workers.emplace_back([&, t]() {
                try {
                    rcs[t] = lower_range<P>(c, lay, n * t / T, n * (t + 1) / T, parts[t]);
                } catch (const std::bad_alloc&) {
                    rcs[t] = ACG_ERR_OOM;
                } catch (...) {
                    rcs[t] = ACG_ERR_INTERNAL;
                }
            });
        } catch (...) {  // the system refused another thread
            spawn_failed = true;
        }
    }
    for (auto& w : workers) w.join();
    if (spawn_failed) {  // fall back to the sequential pass (identical output)
        parts.clear();
        return lower_range<P>(c, lay, 0, n, sink);
    }
    for (int rc : rcs)
        if (rc != ACG_OK) return rc;
    for (int k = 0; k < 3; ++k) {
        size_t rows = 0, nnz = 0;
        for (const RowSink& p : parts) {
            rows += p.rowptr[k].size() - 1;
            nnz += p.col[k].size();
        }
        if (nnz > 0xFFFFFFF0ull) return ACG_ERR_UNSUPPORTED;
        sink.rowptr[k].reserve(rows + 1);
        sink.col[k].reserve(nnz);
        sink.val[k].reserve(4 * nnz);
        for (RowSink& p : parts) {
            const uint32_t base = (uint32_t)sink.col[k].size();
            for (size_t r = 1; r < p.rowptr[k].size(); ++r) sink.rowptr[k].push_back(base + p.rowptr[k][r]);
            sink.col[k].insert(sink.col[k].end(), p.col[k].begin(), p.col[k].end());
            sink.val[k].insert(sink.val[k].end(), p.val[k].begin(), p.val[k].end());
            std::vector<uint32_t>().swap(p.col[k]);
            std::vector<uint64_t>().swap(p.val[k]);
        }
    }
    return ACG_OK;
}

inline bool lt256(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] < b[i]) return true;
        if (a[i] > b[i]) return false;
    }
    return false;
}

template <class F>
int dispatch(int field, F&& f) {
    if (field == ACG_FIELD_BN254_FR) return f(Bn254Fr{});
    if (field == ACG_FIELD_BLS12_381_FR) return f(Bls12381Fr{});
    return ACG_ERR_BAD_ARG;
}

}  // namespace

namespace acg {
namespace host {
int build_gate_plan(const acg_circuit* c, uint32_t n_in, uint32_t n_mid, uint32_t n_out, GatePlan& out) {
    if (!c) return ACG_ERR_BAD_ARG;
    uint32_t dims[3];
    circuit_dims(c, dims);
    const Layout lay{std::max(n_in, dims[0]), std::max(n_mid, dims[1]), std::max(n_out, dims[2])};
    out = GatePlan{};
    out.n_in = lay.n_in;
    out.n_mid = lay.n_mid;
    out.n_out = lay.n_out;
    const uint32_t n_cols = 1 + lay.n_in + lay.n_mid + lay.n_out;
    const size_t n = c->gates.size();
    std::vector<uint32_t> col_level(n_cols, 0), gate_level(n, 0);
    std::vector<uint8_t> written(n_cols, 0), read(n_cols, 0);
    std::vector<GateRec> recs(n);
    uint32_t n_levels = 0;
    int rc = dispatch(c->field, [&](auto p) {
        using P = decltype(p);
        Row row;
        auto produce = [&](uint32_t col, uint32_t lvl) {
            // a wire assigned twice, or read (as 0) before its gate: only the sequential fold is faithful
            if (written[col] || read[col] || col <= lay.n_in) return false;
            written[col] = 1;
            col_level[col] = lvl;
            return true;
        };
        for (size_t gi = 0; gi < n; ++gi) {
            const GateH& g = c->gates[gi];
            GateRec& r = recs[gi];
            r = GateRec{};
            r.kind = g.kind;
            uint32_t lvl = 0;
            if (g.kind == G_MUL) {
                for (int side = 0; side < 2; ++side) {
                    affine_to_row<P>(side ? g.r : g.l, lay, row);
                    (side ? r.r0 : r.l0) = (uint32_t)out.term_col.size();
                    for (auto& e : row) {
                        if (e.first != 0 && !written[e.first] && e.first > lay.n_in) read[e.first] = 1;
                        lvl = std::max(lvl, col_level[e.first]);
                        out.term_col.push_back(e.first);
                        out.term_coef.insert(out.term_coef.end(), e.second.v, e.second.v + 4);
                    }
                    (side ? r.r1 : r.l1) = (uint32_t)out.term_col.size();
                }
                r.out = lay.col(g.w2);
                if (!produce(r.out, lvl + 1)) return (int)ACG_ERR_UNSUPPORTED;
            } else if (g.kind == G_EQUAL) {
                r.in = lay.col(g.w0);
                r.magic = lay.col(g.w1);
                r.out = lay.col(g.w2);
                if (r.in > lay.n_in && !written[r.in]) return (int)ACG_ERR_BAD_ARG;  // lookup fails: reference panics
                if (r.in >= 1 && r.in <= lay.n_in) out.required_inputs.push_back(r.in);
                lvl = col_level[r.in];
                if (!produce(r.magic, lvl + 1) || !produce(r.out, lvl + 1)) return (int)ACG_ERR_UNSUPPORTED;
            } else {
                r.in = lay.col(g.w0);
                if (r.in > lay.n_in && !written[r.in]) return (int)ACG_ERR_BAD_ARG;
                if (r.in >= 1 && r.in <= lay.n_in) out.required_inputs.push_back(r.in);
                lvl = col_level[r.in];
                r.l0 = (uint32_t)out.split_outs.size();
                for (uint64_t o : g.outs) {
                    const uint32_t oc = lay.col(o);
                    if (!produce(oc, lvl + 1)) return (int)ACG_ERR_UNSUPPORTED;
                    out.split_outs.push_back(oc);
                }
                r.l1 = (uint32_t)out.split_outs.size();
            }
            gate_level[gi] = lvl;  // 0-based level of the gate
            n_levels = std::max(n_levels, lvl + 1);
        }
        return (int)ACG_OK;
    });
    if (rc != ACG_OK) return rc;
    // counting sort by level (stable: circuit order inside a level)
    out.level_ptr.assign((size_t)n_levels + 1, 0);
    for (size_t gi = 0; gi < n; ++gi) ++out.level_ptr[gate_level[gi] + 1];
    for (uint32_t l = 0; l < n_levels; ++l) {
        out.max_width = std::max(out.max_width, out.level_ptr[l + 1]);
        out.level_ptr[l + 1] += out.level_ptr[l];
    }
    out.gates.resize(n);
    std::vector<uint32_t> cursor(out.level_ptr.begin(), out.level_ptr.end() - (n_levels ? 1 : 0));
    for (size_t gi = 0; gi < n; ++gi) out.gates[cursor[gate_level[gi]]++] = recs[gi];
    return ACG_OK;
}
}  // namespace host
}  // namespace acg

extern "C" {

int acg_circuit_parse(int field_id, const uint64_t* words, uint64_t n_words, acg_circuit** out) {
    ACG_TRY
    if (!out || (!words && n_words)) return ACG_ERR_BAD_ARG;
    *out = nullptr;
    acg_circuit* c = new (std::nothrow) acg_circuit();
    if (!c) return ACG_ERR_OOM;
    c->field = field_id;
    int rc = dispatch(field_id, [&](auto p) { return parse_impl<decltype(p)>(words, n_words, c); });
    if (rc != ACG_OK) {
        delete c;
        return rc;
    }
    *out = c;
    return ACG_OK;
    ACG_CATCH()
}
void acg_circuit_free(acg_circuit* c) { delete c; }
uint64_t acg_circuit_num_gates(const acg_circuit* c) { return c ? c->gates.size() : 0; }
uint64_t acg_circuit_num_roots(const acg_circuit* c) { return c ? c->n_roots : 0; }

// validArithCircuit, src/Circuit/Arithmetic.hs:158-185
int acg_circuit_valid(const acg_circuit* c) {
    ACG_TRY
    if (!c) return 0;
    std::vector<uint8_t> defined_mid;
    auto is_defined = [&](uint64_t w) {
        if (wkind(w) == ACG_WIRE_INPUT) return true;
        if (wkind(w) == ACG_WIRE_OUTPUT) return false;
        return wix(w) < defined_mid.size() && defined_mid[wix(w)] != 0;
    };
    bool ok = true;
    for (const GateH& g : c->gates) {
        std::vector<uint64_t> outs, used;
        if (g.kind == G_MUL) {
            outs.push_back(g.w2);
            for (const AffTok& t : g.l)
                if (t.op == TOK_VAR) used.push_back(t.wire);
            for (const AffTok& t : g.r)
                if (t.op == TOK_VAR) used.push_back(t.wire);
        } else if (g.kind == G_EQUAL) {
            outs.push_back(g.w2);
            used.push_back(g.w0);  // the magic wire is filled in by evaluation (:178-180)
        } else {
            outs = g.outs;
            used.push_back(g.w0);
        }
        for (uint64_t o : outs) ok = ok && wkind(o) != ACG_WIRE_INPUT;
        for (uint64_t u : used) ok = ok && is_defined(u);
        for (uint64_t o : outs)
            if (wkind(o) == ACG_WIRE_INTERMEDIATE) {
                if (wix(o) >= defined_mid.size()) defined_mid.resize((size_t)wix(o) + 1, 0);
                defined_mid[wix(o)] = 1;
            }
    }
    return ok ? 1 : 0;
    ACG_CATCH()
}

int acg_generate_assignment(const acg_circuit* c, const uint32_t* input_ix, const uint64_t* input_vals,
                            uint32_t n_inputs, acg_assignment** out) {
    ACG_TRY
    if (!c || !out || (n_inputs && (!input_ix || !input_vals))) return ACG_ERR_BAD_ARG;
    *out = nullptr;
    acg_assignment* a = new (std::nothrow) acg_assignment();
    if (!a) return ACG_ERR_OOM;
    a->field = c->field;
    int rc = dispatch(c->field, [&](auto p) {
        using P = decltype(p);
        for (uint32_t i = 0; i < n_inputs; ++i) {  // initialQapSet, src/QAP.hs:591-595
            El v{{input_vals[4 * i], input_vals[4 * i + 1], input_vals[4 * i + 2], input_vals[4 * i + 3]}};
            if (Fr<P>::geq_mod(v)) return (int)ACG_ERR_NON_CANONICAL;
            if (input_ix[i] >= 0x7FFFFFFFu) return (int)ACG_ERR_BAD_ARG;
            wm_set(a->part[ACG_WIRE_INPUT], input_ix[i], Fr<P>::to_mont(v));
        }
        return eval_impl<P>(c, a);
    });
    if (rc != ACG_OK) {
        delete a;
        return rc;
    }
    *out = a;
    return ACG_OK;
    ACG_CATCH()
}
void acg_assignment_free(acg_assignment* a) { delete a; }

int acg_circuit_plan_stats(const acg_circuit* c, uint32_t* n_levels, uint32_t* max_width) {
    ACG_TRY
    if (!c) return ACG_ERR_BAD_ARG;
    acg::host::GatePlan plan;
    const int rc = acg::host::build_gate_plan(c, 0, 0, 0, plan);
    if (rc != ACG_OK) return rc;
    if (n_levels) *n_levels = plan.level_ptr.empty() ? 0u : (uint32_t)plan.level_ptr.size() - 1u;
    if (max_width) *max_width = plan.max_width;
    return ACG_OK;
    ACG_CATCH()
}

int acg_assignment_dims(const acg_assignment* a, uint32_t* n_in, uint32_t* n_mid, uint32_t* n_out) {
    ACG_TRY
    if (!a) return ACG_ERR_BAD_ARG;
    uint32_t* outs[3] = {n_in, n_mid, n_out};
    for (int k = 0; k < 3; ++k) {
        uint32_t mx = 0;  // maxKey + 1 over PRESENT keys (src/QAP.hs:612-620)
        for (size_t i = a->part[k].present.size(); i-- > 0;)
            if (a->part[k].present[i]) {
                mx = (uint32_t)i + 1;
                break;
            }
        if (outs[k]) *outs[k] = mx;
    }
    return ACG_OK;
    ACG_CATCH()
}

int acg_assignment_lookup(const acg_assignment* a, uint64_t wire, uint64_t out[4]) {
    ACG_TRY
    if (!a || !out || !wire_ok(wire)) return ACG_ERR_BAD_ARG;
    El v;
    if (!wm_get(a->part[wkind(wire)], wix(wire), v)) return 0;
    dispatch(a->field, [&](auto p) {
        const El c = Fr<decltype(p)>::from_mont(v);
        std::memcpy(out, c.v, 32);
        return 0;
    });
    return 1;
    ACG_CATCH()
}

int acg_assignment_update(acg_assignment* a, uint64_t wire, const uint64_t val[4]) {
    ACG_TRY
    if (!a || !val || !wire_ok(wire)) return ACG_ERR_BAD_ARG;
    return dispatch(a->field, [&](auto p) {
        using P = decltype(p);
        El v{{val[0], val[1], val[2], val[3]}};
        if (Fr<P>::geq_mod(v)) return (int)ACG_ERR_NON_CANONICAL;
        wm_set(a->part[wkind(wire)], wix(wire), Fr<P>::to_mont(v));
        return (int)ACG_OK;
    });
    ACG_CATCH()
}

int acg_assignment_to_vector(const acg_assignment* a, uint32_t n_in, uint32_t n_mid, uint32_t n_out, uint64_t* w) {
    ACG_TRY
    if (!a || !w) return ACG_ERR_BAD_ARG;
    const uint32_t dims[3] = {n_in, n_mid, n_out};
    for (int k = 0; k < 3; ++k)
        for (size_t i = dims[k]; i < a->part[k].present.size(); ++i)
            if (a->part[k].present[i]) return ACG_ERR_BAD_ARG;  // layout does not cover the assignment
    const uint64_t n_cols = 1ull + n_in + n_mid + n_out;
    std::memset(w, 0, n_cols * 32);
    w[0] = 1;  // qapSetConstant = 1 (initialQapSet)
    return dispatch(a->field, [&](auto p) {
        using P = decltype(p);
        uint64_t base = 1;
        for (int k = 0; k < 3; ++k) {
            const WireMap& m = a->part[k];
            for (size_t i = 0; i < m.val.size() && i < dims[k]; ++i)
                if (m.present[i]) {
                    const El c = Fr<P>::from_mont(m.val[i]);
                    std::memcpy(w + 4 * (base + i), c.v, 32);
                }
            base += dims[k];
        }
        return (int)ACG_OK;
    });
    ACG_CATCH()
}

int acg_circuit_to_r1cs(const acg_circuit* c, const uint64_t* roots, uint64_t root_start, uint32_t n_in,
                        uint32_t n_mid, uint32_t n_out, acg_r1cs_host** out) {
    ACG_TRY
    if (!c || !out) return ACG_ERR_BAD_ARG;
    *out = nullptr;
    uint32_t dims[3];
    circuit_dims(c, dims);
    Layout lay{n_in ? n_in : dims[0], n_mid ? n_mid : dims[1], n_out ? n_out : dims[2]};
    if (lay.n_in < dims[0] || lay.n_mid < dims[1] || lay.n_out < dims[2]) return ACG_ERR_BAD_ARG;
    if (c->n_roots > 0xFFFFFFF0ull) return ACG_ERR_UNSUPPORTED;
    RowSink sink;
    int rc = dispatch(c->field, [&](auto p) { return lower_impl<decltype(p)>(c, lay, sink); });
    if (rc != ACG_OK) return rc;
    const uint32_t n_rows = (uint32_t)c->n_roots;
    acg_r1cs_host* m = new (std::nothrow) acg_r1cs_host();
    if (!m) return ACG_ERR_OOM;
    m->field = c->field;
    m->n_rows = n_rows;
    m->n_in = lay.n_in;
    m->n_mid = lay.n_mid;
    m->n_out = lay.n_out;
    m->n_cols = 1 + lay.n_in + lay.n_mid + lay.n_out;
    m->roots.resize(4ull * n_rows);
    std::vector<uint32_t> order(n_rows);
    std::iota(order.begin(), order.end(), 0u);
    bool sorted = true;
    if (roots) {
        rc = dispatch(c->field, [&](auto p) {
            for (uint32_t r = 0; r < n_rows; ++r) {
                El v{{roots[4 * r], roots[4 * r + 1], roots[4 * r + 2], roots[4 * r + 3]}};
                if (Fr<decltype(p)>::geq_mod(v)) return (int)ACG_ERR_NON_CANONICAL;
            }
            return (int)ACG_OK;
        });
        if (rc != ACG_OK) {
            delete m;
            return rc;
        }
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t x, uint32_t y) { return lt256(roots + 4ull * x, roots + 4ull * y); });
        for (uint32_t r = 0; r < n_rows; ++r) {
            if (order[r] != r) sorted = false;
            std::memcpy(&m->roots[4ull * r], roots + 4ull * order[r], 32);
            // Map.fromList would silently merge rows that share a root (src/QAP.hs:233-239): reject
            if (r && !lt256(&m->roots[4ull * (r - 1)], &m->roots[4ull * r])) {
                delete m;
                return ACG_ERR_BAD_ARG;
            }
        }
    } else {
        for (uint32_t r = 0; r < n_rows; ++r) m->roots[4ull * r] = root_start + r;  // fromIntegral <$> fresh
    }
    for (int k = 0; k < 3; ++k) {
        if (sorted) {
            m->rowptr[k] = std::move(sink.rowptr[k]);
            m->col[k] = std::move(sink.col[k]);
            m->val[k] = std::move(sink.val[k]);
        } else {  // rows in ascending-root order (Map key order)
            m->rowptr[k].assign(1, 0);
            for (uint32_t r = 0; r < n_rows; ++r) {
                const uint32_t s = sink.rowptr[k][order[r]], e = sink.rowptr[k][order[r] + 1];
                m->col[k].insert(m->col[k].end(), sink.col[k].begin() + s, sink.col[k].begin() + e);
                m->val[k].insert(m->val[k].end(), sink.val[k].begin() + 4ull * s, sink.val[k].begin() + 4ull * e);
                m->rowptr[k].push_back((uint32_t)m->col[k].size());
            }
        }
    }
    *out = m;
    return ACG_OK;
    ACG_CATCH()
}

void acg_r1cs_host_free(acg_r1cs_host* m) { delete m; }

int acg_r1cs_host_dims(const acg_r1cs_host* m, uint32_t* n_rows, uint32_t* n_cols, uint32_t* n_in, uint32_t* n_mid,
                       uint32_t* n_out) {
    ACG_TRY
    if (!m) return ACG_ERR_BAD_ARG;
    if (n_rows) *n_rows = m->n_rows;
    if (n_cols) *n_cols = m->n_cols;
    if (n_in) *n_in = m->n_in;
    if (n_mid) *n_mid = m->n_mid;
    if (n_out) *n_out = m->n_out;
    return ACG_OK;
    ACG_CATCH()
}

int acg_r1cs_host_csr(const acg_r1cs_host* m, int which, acg_csr* out) {
    ACG_TRY
    if (!m || !out || which < 0 || which > 2) return ACG_ERR_BAD_ARG;
    out->rowptr = m->rowptr[which].data();
    out->col = m->col[which].data();
    out->val = m->val[which].data();
    out->nnz = m->col[which].size();
    return ACG_OK;
    ACG_CATCH()
}

const uint64_t* acg_r1cs_host_roots(const acg_r1cs_host* m) { return m ? m->roots.data() : nullptr; }

void acg_free(void* p) { std::free(p); }

}  // extern "C"
