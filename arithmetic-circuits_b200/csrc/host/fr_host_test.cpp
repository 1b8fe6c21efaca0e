// CPU-only harness: runs the *device* field algorithms of fr.cuh against the emulated carry flag so
// tests/test_fr_host.py can compare them with Python big integers without a GPU.
#include "../fr.cuh"
#include <cstring>

using namespace acg;

template <class P>
static int binop(int op, const uint32_t* a, const uint32_t* b, uint32_t* o, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        fr_t x, y, z;
        std::memcpy(x.l, a + 8 * i, 32);
        std::memcpy(y.l, b + 8 * i, 32);
        if (op < 7 && (!fr_is_canonical<P>(x) || (op != 3 && op != 4 && op != 5 && !fr_is_canonical<P>(y)))) return -2;
        switch (op) {
            case 0: z = fr_add<P>(x, y); break;
            case 1: z = fr_sub<P>(x, y); break;
            case 2: z = fr_from_mont<P>(fr_mul<P>(fr_to_mont<P>(x), fr_to_mont<P>(y))); break;
            case 3: z = fr_from_mont<P>(fr_inv<P>(fr_to_mont<P>(x))); break;
            case 4: z = fr_to_mont<P>(x); break;
            case 5: z = fr_neg<P>(x); break;
            case 6: z = fr_mul<P>(x, y); break;  // raw Montgomery product x*y/R
            case 7: z = fr_mul<P>(x, y); break;  // x <= p, y any 256-bit value (unreduced row sum)
            case 8: z = fr_add<P>(x, y); break;  // x, y <= p: result in [0, p]
            default: return -1;
        }
        std::memcpy(o + 8 * i, z.l, 32);
    }
    return 0;
}

extern "C" int acg_hosttest_binop(int field_id, int op, const uint32_t* a, const uint32_t* b, uint32_t* o, uint64_t n) {
    if (field_id == 0) return binop<Bn254Fr>(op, a, b, o, n);
    if (field_id == 1) return binop<Bls12381Fr>(op, a, b, o, n);
    return -1;
}
