// Shared host-side structures behind the opaque handles of include/acg.h (circuit IR mirror).
#pragma once
#include <cstdint>
#include <vector>

#include "fr_host.hpp"

struct acg_circuit;
namespace acg {
namespace host {

struct AffTok {   // one post-order token of an AffineCircuit (src/Circuit/Affine.hs:26-31)
    uint8_t op;   // 0 Var, 1 ConstGate, 2 Add, 3 ScalarMul
    uint64_t wire;
    El val;       // Montgomery form
};

struct GateH {    // Gate Wire f (src/Circuit/Arithmetic.hs:44-59)
    uint8_t kind; // 1 Mul, 2 Equal, 3 Split
    uint64_t w0;  // Equal/Split: input wire
    uint64_t w1;  // Equal: magic wire
    uint64_t w2;  // Mul/Equal: output wire
    std::vector<AffTok> l, r;
    std::vector<uint64_t> outs;
};

struct WireMap {  // Map Int f with dense storage
    std::vector<El> val;  // Montgomery form
    std::vector<uint8_t> present;
};

// Device-ready, level-ordered description of a circuit for the witness-generation kernel (K6).
// Level of a gate = 1 + the highest level among the gates that produce its input wires (inputs: level 0), so the
// gates of one level only read wires of earlier levels and can be evaluated in parallel.
struct GateRec {          // 32 bytes
    uint32_t kind;        // 1 Mul, 2 Equal, 3 Split
    uint32_t out;         // Mul / Equal: witness column of the output wire
    uint32_t l0, l1;      // Mul: left terms [l0, l1);  Split: outputs [l0, l1) in split_outs
    uint32_t r0, r1;      // Mul: right terms [r0, r1)
    uint32_t in;          // Equal / Split: witness column of the input wire
    uint32_t magic;       // Equal: witness column of the magic wire
};
struct GatePlan {
    uint32_t n_in = 0, n_mid = 0, n_out = 0;
    std::vector<uint32_t> level_ptr;    // gates of level l: [level_ptr[l], level_ptr[l + 1])
    std::vector<GateRec> gates;         // level order
    std::vector<uint32_t> term_col;     // affine-map terms (column 0 = the constant wire)
    std::vector<uint64_t> term_coef;    // 4 limbs each, Montgomery form
    std::vector<uint32_t> split_outs;   // witness columns of Split outputs
    uint32_t max_width = 0;             // widest level
    // input wires (columns 1 .. n_in) that an Equal or Split gate takes as its input: the reference looks those up
    // and panics when the assignment lacks them (src/QAP.hs:445,474) -- unlike the terms of a Mul gate, for which a
    // missing wire counts as 0 (src/Circuit/Affine.hs:121-125)
    std::vector<uint32_t> required_inputs;
};
// affineCircuitToAffineMap per Mul side + levelisation.  0 / ACG_ERR_*; ACG_ERR_UNSUPPORTED when the gate list is
// not in single-assignment, define-before-use form (then only the sequential host fold is faithful).
int build_gate_plan(const struct ::acg_circuit* c, uint32_t n_in, uint32_t n_mid, uint32_t n_out, GatePlan& out);

}  // namespace host
}  // namespace acg

struct acg_circuit {
    int field = 0;
    std::vector<acg::host::GateH> gates;
    uint64_t n_roots = 0;
};

struct acg_assignment {  // QapSet f; the constant is always 1 (initialQapSet)
    int field = 0;
    acg::host::WireMap part[3];  // indexed by ACG_WIRE_*
};

struct acg_r1cs_host {
    int field = 0;
    uint32_t n_rows = 0, n_cols = 0, n_in = 0, n_mid = 0, n_out = 0;
    std::vector<uint32_t> rowptr[3], col[3];
    std::vector<uint64_t> val[3];   // canonical limbs
    std::vector<uint64_t> roots;    // canonical limbs, ascending
};
