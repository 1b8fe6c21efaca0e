// K6 -- witness generation on the device: the left fold of evalGate over the gate list
// (reference src/Circuit/Arithmetic.hs:106-145, 221-235; wrappers generateAssignment / generateAssignmentGate,
// src/QAP.hs:579-603) restated level by level.  The host (csrc/host/circuit.cpp: build_gate_plan) flattens both sides
// of every Mul gate to their affine maps (affineCircuitToAffineMap, src/Circuit/Affine.hs:90-105) and sorts the gates
// by dependency level; the gates of one level only read wires of earlier levels, so one thread evaluates one gate:
//     Mul    out   = (sum coef * w[col])_left * (sum coef * w[col])_right           (:120-124; missing wire = 0)
//     Equal  out   = (in != 0),  magic = in^-1 or 0                                   (:125-133)
//     Split  out_i = bit i of the canonical residue of in                             (:134-145)
// Levels are separated by a grid-wide barrier (cooperative launch) or, for circuits whose levels are narrower than
// one thread block (long dependency chains), by __syncthreads in a single CTA -- a level then costs ~1 us instead
// of a kernel launch.  The witness stays on the device in Montgomery form, in qapSetToMap order, ready for K2.
#include <cooperative_groups.h>

#include "dev.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace acg {

// L2-coherent load of a witness element another thread block may have written in the previous level
__device__ __forceinline__ fr_t ld_w_cg(const fr_t* p) {
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p)), b = __ldcg(reinterpret_cast<const uint4*>(p) + 1);
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

template <class P>
__device__ __forceinline__ fr_t affine_dot(const uint32_t* __restrict__ col, const fr_t* __restrict__ coef, uint32_t t0,
                                           uint32_t t1, const fr_t* w) {
    fr_t acc = fr_zero<P>();
    for (uint32_t t = t0; t < t1; ++t) acc = fr_add<P>(acc, fr_mul<P>(coef[t], ld_w_cg(w + col[t])));
    return acc;
}

template <class P>
__device__ __forceinline__ void eval_gate(const WitnessGate& g, const uint32_t* __restrict__ term_col,
                                          const fr_t* __restrict__ term_coef, const uint32_t* __restrict__ split_outs,
                                          fr_t* w) {
    if (g.kind == 1u) {
        const fr_t l = affine_dot<P>(term_col, term_coef, g.l0, g.l1, w);
        const fr_t r = affine_dot<P>(term_col, term_coef, g.r0, g.r1, w);
        w[g.out] = fr_mul<P>(l, r);
    } else if (g.kind == 2u) {
        const fr_t x = ld_w_cg(w + g.in);
        const bool z = fr_is_zero(x);
        w[g.magic] = z ? fr_zero<P>() : fr_inv<P>(x);
        w[g.out] = z ? fr_zero<P>() : fr_one<P>();
    } else {
        const fr_t x = fr_from_mont<P>(ld_w_cg(w + g.in));  // testBit (fromP in) i
        for (uint32_t i = g.l0; i < g.l1; ++i) {
            const uint32_t bit = i - g.l0;
            const bool on = bit < 256u && ((x.l[bit >> 5] >> (bit & 31u)) & 1u);
            w[split_outs[i]] = on ? fr_one<P>() : fr_zero<P>();
        }
    }
}

template <class P, bool GRID>
__global__ void __launch_bounds__(GRID ? 256 : 1024)
    k_witness_levels(const WitnessGate* __restrict__ gates, const uint32_t* __restrict__ level_ptr, uint32_t n_levels,
                     const uint32_t* __restrict__ term_col, const fr_t* __restrict__ term_coef,
                     const uint32_t* __restrict__ split_outs, fr_t* w) {
    const uint32_t tid = GRID ? blockIdx.x * blockDim.x + threadIdx.x : threadIdx.x;
    const uint32_t stride = GRID ? gridDim.x * blockDim.x : blockDim.x;
    for (uint32_t l = 0; l < n_levels; ++l) {
        const uint32_t g1 = level_ptr[l + 1];
        for (uint32_t g = level_ptr[l] + tid; g < g1; g += stride) eval_gate<P>(gates[g], term_col, term_coef, split_outs, w);
        if (GRID) {
            cg::this_grid().sync();
        } else {
            __threadfence_block();
            __syncthreads();
        }
    }
}

cudaError_t launch_witness_levels(int field, const WitnessGate* gates, const uint32_t* level_ptr, uint32_t n_levels,
                                  uint32_t max_width, const uint32_t* term_col, const fr_t* term_coef,
                                  const uint32_t* split_outs, fr_t* w, int sm_count, cudaStream_t s) {
    if (n_levels == 0) return cudaSuccess;
    void* args[] = {(void*)&gates, (void*)&level_ptr, (void*)&n_levels, (void*)&term_col,
                    (void*)&term_coef, (void*)&split_outs, (void*)&w};
    const bool grid_mode = max_width > 2048u;  // some level is wide enough to feed more than a couple of CTAs
    if (!grid_mode) {
        if (field == 0)
            k_witness_levels<Bn254Fr, false><<<1, 1024, 0, s>>>(gates, level_ptr, n_levels, term_col, term_coef, split_outs, w);
        else if (field == 1)
            k_witness_levels<Bls12381Fr, false><<<1, 1024, 0, s>>>(gates, level_ptr, n_levels, term_col, term_coef, split_outs, w);
        else
            return cudaErrorInvalidValue;
        return cudaGetLastError();
    }
    const void* fn = field == 0 ? (const void*)k_witness_levels<Bn254Fr, true>
                   : field == 1 ? (const void*)k_witness_levels<Bls12381Fr, true> : nullptr;
    if (!fn) return cudaErrorInvalidValue;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    unsigned grid = (unsigned)(sm_count * per_sm);
    const unsigned need = (max_width + 255u) / 256u;
    if (grid > need) grid = need;
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, s);
}

}  // namespace acg
