// Shared host-side structures behind the opaque handles of include/acg.h (circuit IR mirror).
#pragma once
#include <cstdint>
#include <vector>

#include "fr_host.hpp"

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
