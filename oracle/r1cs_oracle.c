/* CPU ORACLE (test infrastructure, NOT product code) -- plain-C restatement of the reference's
 * R1CS witness check and QAP polynomial path at O(nnz) / O(n log n), for sizes the Python big-int
 * oracle (oracle/qap_oracle.py) cannot reach, and as the CPU baseline bench.py reports.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product never links or dlopens it.
 *
 * What it restates (reference = sdiehl/arithmetic-circuits @ 18e15de):
 *   orc_r1cs_eval / orc_r1cs_check : per row, dotProduct (src/Circuit/Affine.hs:121-125) of the sparse
 *       A/B/C row with the witness vector (missing wire = 0), then the evaluation-domain form of
 *       verificationWitnessZk's predicate (src/QAP.hs:309-327): (A.w)_g * (B.w)_g - (C.w)_g == 0 at
 *       every root g  <=>  remainder == 0.
 *   orc_ntt : the DFT behind FFT.interpolate (src/QAP.hs:521-523; galois-fft-0.1.0, un-vendored --
 *       published radix-2 algorithm restated; omega = getRootOfUnity k = g^((r-1)/2^k), pairing-1.0.0).
 *   orc_qap_witness : a = iNTT(A.w) + d1*T, b, c likewise, h = (a*b - c) / T with T = X^N - 1
 *       (src/QAP.hs:314-327 with the linearity collapse of SURVEY R9), via a coset evaluation.
 *   orc_poly_mul_divmod_check : the reference-SHAPED check -- schoolbook product and long division
 *       (src/QAP.hs:325-327) -- for small n, used to validate the coset path and to show the O(n^2) wall.
 *   orc_fr_* : Prime r arithmetic (galois-field-1.0.2) as 4x64-bit Montgomery with unsigned __int128.
 *
 * PARITY STATUS: Boolean results pinned by the reference's own unit tests (through
 * tests/test_oracle.py); field-element / coefficient values "parity unpinned" (no reference test or
 * runnable reference pins them) but cross-checked against the independent Python big-int oracle.
 *
 * ABI: field elements are 4 little-endian uint64 limbs, canonical (< r) on both sides.
 * Build: oracle/Makefile -> oracle/_build/liboracle.so  (gcc -O3 -fopenmp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;

typedef struct {
    u64 p[4];      /* modulus r */
    u64 one[4];    /* R mod r  (Montgomery 1) */
    u64 r2[4];     /* R^2 mod r */
    u64 ninv;      /* -r^-1 mod 2^64 */
    u64 gen;       /* multiplicative generator defining the roots of unity */
    int two_adicity;
    int ready;
} orc_field;

static orc_field g_fields[2] = {
    /* BN254 Fr, README.tex.md:59-61 */
    {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
     {0}, {0}, 0, 5, 28, 0},
    /* BLS12-381 Fr */
    {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
     {0}, {0}, 0, 7, 32, 0},
};

/* ---- 256-bit helpers ------------------------------------------------------------------- */
static inline int ge4(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline u64 add4(u64 o[4], const u64 a[4], const u64 b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a[i] + b[i]; o[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static inline u64 sub4(u64 o[4], const u64 a[4], const u64 b[4]) {
    u64 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - br;
        o[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
    return br;
}
static inline int is_zero4(const u64 a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }

static inline void fr_add(const orc_field* F, u64 o[4], const u64 a[4], const u64 b[4]) {
    u64 t[4], s[4];
    u64 carry = add4(t, a, b);
    u64 borrow = sub4(s, t, F->p);
    if (carry || !borrow) memcpy(o, s, 32); else memcpy(o, t, 32);
}
static inline void fr_sub(const orc_field* F, u64 o[4], const u64 a[4], const u64 b[4]) {
    u64 t[4];
    if (sub4(t, a, b)) add4(t, t, F->p);
    memcpy(o, t, 32);
}
/* Montgomery product a*b*R^-1 mod r, CIOS over 64-bit limbs */
static inline void fr_mul(const orc_field* F, u64 o[4], const u64 a[4], const u64 b[4]) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (u64)c; c >>= 64;
        }
        c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
        u64 m = t[0] * F->ninv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (u64)c; c >>= 64;
        }
        c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
    }
    u64 s[4];
    u64 borrow = sub4(s, t, F->p);
    if (t[4] || !borrow) memcpy(o, s, 32); else memcpy(o, t, 32);
}
static void fr_pow_limbs(const orc_field* F, u64 o[4], const u64 base[4], const u64 e[4]) {
    u64 acc[4], b[4];
    memcpy(acc, F->one, 32); memcpy(b, base, 32);
    for (int i = 0; i < 256; ++i) {
        if ((e[i >> 6] >> (i & 63)) & 1) fr_mul(F, acc, acc, b);
        fr_mul(F, b, b, b);
    }
    memcpy(o, acc, 32);
}
static void fr_inv(const orc_field* F, u64 o[4], const u64 a[4]) { /* a^(r-2) */
    u64 e[4], two[4] = {2, 0, 0, 0};
    sub4(e, F->p, two);
    fr_pow_limbs(F, o, a, e);
}
static inline void fr_to_mont(const orc_field* F, u64 o[4], const u64 a[4]) { fr_mul(F, o, a, F->r2); }
static inline void fr_from_mont(const orc_field* F, u64 o[4], const u64 a[4]) {
    static const u64 lit1[4] = {1, 0, 0, 0};
    fr_mul(F, o, a, lit1);
}

static const orc_field* field_get(int id) {
    if (id < 0 || id > 1) return NULL;
    orc_field* F = &g_fields[id];
    if (!F->ready) {
#pragma omp critical(orc_field_init)
        if (!F->ready) {
            u64 inv = 1; /* Newton: inv = p^-1 mod 2^64 */
            for (int i = 0; i < 6; ++i) inv *= 2 - F->p[0] * inv;
            F->ninv = (u64)0 - inv;
            u64 x[4] = {1, 0, 0, 0};
            for (int i = 0; i < 512; ++i) {        /* x = 2^i mod r by modular doubling */
                if (i == 256) memcpy(F->one, x, 32);
                u64 t[4], s[4];
                u64 carry = add4(t, x, x);
                u64 borrow = sub4(s, t, F->p);
                if (carry || !borrow) memcpy(x, s, 32); else memcpy(x, t, 32);
            }
            memcpy(F->r2, x, 32);
            __sync_synchronize();
            F->ready = 1;
        }
    }
    return F;
}

/* omega_k in Montgomery form: gen^((r-1)/2^k) */
static int root_of_unity_mont(const orc_field* F, int k, u64 o[4]) {
    if (k < 0 || k > F->two_adicity) return -1;
    u64 e[4], one1[4] = {1, 0, 0, 0}, g[4] = {F->gen, 0, 0, 0}, gm[4];
    sub4(e, F->p, one1);
    for (int s = 0; s < k; ++s) { /* e >>= 1 */
        for (int i = 0; i < 4; ++i) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);
    }
    fr_to_mont(F, gm, g);
    fr_pow_limbs(F, o, gm, e);
    return 0;
}

/* ---- exported field helpers (tests pin these against the Python big-int oracle) ---------- */
int orc_field_constants(int field_id, u64 p[4], u64 one[4], u64 r2[4], u64* ninv) {
    const orc_field* F = field_get(field_id);
    if (!F) return -1;
    memcpy(p, F->p, 32); memcpy(one, F->one, 32); memcpy(r2, F->r2, 32); *ninv = F->ninv;
    return 0;
}
/* op: 0 add, 1 sub, 2 mul, 3 inv(a) ; canonical in/out, elementwise over n */
int orc_fr_binop(int field_id, int op, const u64* a, const u64* b, u64* o, u64 n) {
    const orc_field* F = field_get(field_id);
    if (!F) return -1;
    for (u64 i = 0; i < n; ++i) {
        u64 x[4], y[4], z[4];
        if (ge4(a + 4 * i, F->p)) return -2;
        fr_to_mont(F, x, a + 4 * i);
        if (op != 3) { if (ge4(b + 4 * i, F->p)) return -2; fr_to_mont(F, y, b + 4 * i); }
        switch (op) {
            case 0: fr_add(F, z, x, y); break;
            case 1: fr_sub(F, z, x, y); break;
            case 2: fr_mul(F, z, x, y); break;
            case 3: if (is_zero4(x)) memset(z, 0, 32); else fr_inv(F, z, x); break;
            default: return -1;
        }
        fr_from_mont(F, o + 4 * i, z);
    }
    return 0;
}
int orc_root_of_unity(int field_id, int k, u64 o[4]) {
    const orc_field* F = field_get(field_id);
    if (!F) return -1;
    u64 w[4];
    if (root_of_unity_mont(F, k, w)) return -1;
    fr_from_mont(F, o, w);
    return 0;
}

/* ---- R1CS: CSR x witness ---------------------------------------------------------------- */
typedef struct {
    const u32* rowptr;  /* n_rows + 1 */
    const u32* col;     /* nnz */
    const u64* val;     /* nnz * 4 limbs, canonical */
    u64 nnz;
} orc_csr;

static inline void row_dot(const orc_field* F, const orc_csr* M, u32 row, const u64* w_mont, u64 acc[4]) {
    memset(acc, 0, 32);
    for (u32 k = M->rowptr[row]; k < M->rowptr[row + 1]; ++k) {
        u64 t[4];
        /* canonical coeff * Montgomery witness * R^-1 = canonical product */
        fr_mul(F, t, M->val + 4 * (u64)k, w_mont + 4 * (u64)M->col[k]);
        fr_add(F, acc, acc, t);
    }
}

/* Aw/Bw/Cw (may be NULL) receive canonical values.  Returns 0, or <0 on bad argument. */
int orc_r1cs_eval_check(int field_id, u32 n_rows, u32 n_cols, const orc_csr* A, const orc_csr* B,
                        const orc_csr* C, const u64* w, u64* Aw, u64* Bw, u64* Cw,
                        u64* n_violations, u64* first_bad_row, int n_threads) {
    const orc_field* F = field_get(field_id);
    if (!F) return -1;
    u64* wm = (u64*)malloc((size_t)n_cols * 32 + 32);
    if (!wm) return -3;
    int bad_arg = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static) reduction(| : bad_arg)
    for (long i = 0; i < (long)n_cols; ++i) {
        if (ge4(w + 4 * i, F->p)) bad_arg |= 1;
        fr_to_mont(F, wm + 4 * i, w + 4 * i);
    }
    if (bad_arg) { free(wm); return -2; }
    u64 viol = 0, first = UINT64_MAX;
#pragma omp parallel for schedule(static) reduction(+ : viol) reduction(min : first)
    for (long r = 0; r < (long)n_rows; ++r) {
        u64 a[4], b[4], c[4], ab[4], am[4];
        row_dot(F, A, (u32)r, wm, a);
        row_dot(F, B, (u32)r, wm, b);
        row_dot(F, C, (u32)r, wm, c);
        if (Aw) memcpy(Aw + 4 * r, a, 32);
        if (Bw) memcpy(Bw + 4 * r, b, 32);
        if (Cw) memcpy(Cw + 4 * r, c, 32);
        fr_to_mont(F, am, a);
        fr_mul(F, ab, am, b); /* canonical a*b */
        if (memcmp(ab, c, 32) != 0) { viol += 1; if ((u64)r < first) first = (u64)r; }
    }
    free(wm);
    if (n_violations) *n_violations = viol;
    if (first_bad_row) *first_bad_row = first;
    return 0;
}

/* ---- NTT ---------------------------------------------------------------------------------- */
static void bit_reverse_permute(u64* a, int log_n) {
    u64 n = 1ULL << log_n;
    for (u64 i = 0; i < n; ++i) {
        u64 j = 0;
        for (int b = 0; b < log_n; ++b) j |= ((i >> b) & 1ULL) << (log_n - 1 - b);
        if (i < j) {
            u64 t[4];
            memcpy(t, a + 4 * i, 32); memcpy(a + 4 * i, a + 4 * j, 32); memcpy(a + 4 * j, t, 32);
        }
    }
}

/* in-place on Montgomery-form data, natural order in and out; omega_m = primitive 2^log_n root */
static int ntt_mont(const orc_field* F, u64* a, int log_n, const u64 omega_m[4]) {
    u64 n = 1ULL << log_n;
    if (log_n == 0) return 0;
    u64* tw = (u64*)malloc((size_t)(n / 2) * 32);
    if (!tw) return -3;
    memcpy(tw, F->one, 32);
    for (u64 i = 1; i < n / 2; ++i) fr_mul(F, tw + 4 * i, tw + 4 * (i - 1), omega_m);
    bit_reverse_permute(a, log_n);
    for (int s = 1; s <= log_n; ++s) {
        u64 len = 1ULL << s, half = len >> 1, stride = n >> s;
#pragma omp parallel for schedule(static)
        for (long blk = 0; blk < (long)(n / len); ++blk) {
            u64* base = a + 4 * (u64)blk * len;
            for (u64 j = 0; j < half; ++j) {
                u64 v[4], u[4];
                fr_mul(F, v, base + 4 * (j + half), tw + 4 * (j * stride));
                memcpy(u, base + 4 * j, 32);
                fr_add(F, base + 4 * j, u, v);
                fr_sub(F, base + 4 * (j + half), u, v);
            }
        }
    }
    free(tw);
    return 0;
}

/* canonical in/out.  inverse != 0: inverse DFT (omega^-1, scaled by 1/n). */
int orc_ntt(int field_id, u64* data, int log_n, int inverse, int n_threads) {
    const orc_field* F = field_get(field_id);
    if (!F) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    u64 n = 1ULL << log_n, om[4];
    if (root_of_unity_mont(F, log_n, om)) return -1;
    if (inverse) fr_inv(F, om, om);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) fr_to_mont(F, data + 4 * i, data + 4 * i);
    int rc = ntt_mont(F, data, log_n, om);
    if (rc) return rc;
    u64 ninv[4] = {n, 0, 0, 0};
    if (inverse) { fr_to_mont(F, ninv, ninv); fr_inv(F, ninv, ninv); }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) {
        if (inverse) fr_mul(F, data + 4 * i, data + 4 * i, ninv);
        fr_from_mont(F, data + 4 * i, data + 4 * i);
    }
    return 0;
}

/* ---- QAP witness polynomials via coset (T = X^N - 1) ------------------------------------------
 * in : aw,bw,cw = A.w, B.w, C.w zero-padded to N = 2^log_n (canonical), delta[3][4]
 * out: a,b,c (N+1 coefficients each, includes the delta*T terms), h (N+1 coefficients; deg h <= N
 *      only when delta1*delta2 != 0), *divisible = 1 iff every residual aw*bw-cw is 0.
 * When not divisible h is the polynomial part computed anyway (caller ignores it, like Nothing). */
int orc_qap_witness(int field_id, int log_n, const u64* aw, const u64* bw, const u64* cw,
                    const u64* delta, u64* a_out, u64* b_out, u64* c_out, u64* h_out,
                    int* divisible, int n_threads) {
    const orc_field* F = field_get(field_id);
    if (!F) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    u64 N = 1ULL << log_n;
    u64 om[4], omi[4], g[4], gi[4], gN[4], zinv[4], ninv[4] = {N, 0, 0, 0};
    if (root_of_unity_mont(F, log_n, om)) return -1;
    fr_inv(F, omi, om);
    fr_to_mont(F, ninv, ninv); fr_inv(F, ninv, ninv);
    { u64 gg[4] = {F->gen, 0, 0, 0}; fr_to_mont(F, g, gg); }   /* coset shift = generator */
    fr_inv(F, gi, g);
    { u64 e[4] = {N, 0, 0, 0}; fr_pow_limbs(F, gN, g, e); }
    { u64 t[4]; fr_sub(F, t, gN, F->one); fr_inv(F, zinv, t); }  /* 1/(g^N - 1) */

    u64* buf[3];
    int all_ok = 1;
    for (int k = 0; k < 3; ++k) {
        buf[k] = (u64*)malloc((size_t)N * 32);
        if (!buf[k]) return -3;
    }
    /* divisibility == all residuals zero */
#pragma omp parallel for schedule(static) reduction(& : all_ok)
    for (long i = 0; i < (long)N; ++i) {
        u64 x[4], y[4], z[4], xy[4];
        fr_to_mont(F, x, aw + 4 * i); fr_to_mont(F, y, bw + 4 * i); fr_to_mont(F, z, cw + 4 * i);
        fr_mul(F, xy, x, y);
        if (memcmp(xy, z, 32) != 0) all_ok = 0;
        memcpy(buf[0] + 4 * i, x, 32); memcpy(buf[1] + 4 * i, y, 32); memcpy(buf[2] + 4 * i, z, 32);
    }
    u64* outs[3] = {a_out, b_out, c_out};
    for (int k = 0; k < 3; ++k) {
        int rc = ntt_mont(F, buf[k], log_n, omi);
        if (rc) return rc;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i) fr_mul(F, buf[k] + 4 * i, buf[k] + 4 * i, ninv);
        /* coefficients (without delta terms) -> outs, canonical */
        for (u64 i = 0; i < N; ++i) fr_from_mont(F, outs[k] + 4 * i, buf[k] + 4 * i);
        memset(outs[k] + 4 * N, 0, 32);
    }
    /* coset evaluation: coeff_j *= g^j, forward NTT */
    u64* hb = (u64*)malloc((size_t)N * 32);
    if (!hb) return -3;
    for (int k = 0; k < 3; ++k) {
        u64 p[4]; memcpy(p, F->one, 32);
        for (u64 i = 0; i < N; ++i) { fr_mul(F, buf[k] + 4 * i, buf[k] + 4 * i, p); fr_mul(F, p, p, g); }
        int rc = ntt_mont(F, buf[k], log_n, om);
        if (rc) return rc;
    }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; ++i) {
        u64 t[4];
        fr_mul(F, t, buf[0] + 4 * i, buf[1] + 4 * i);
        fr_sub(F, t, t, buf[2] + 4 * i);
        fr_mul(F, hb + 4 * i, t, zinv);
    }
    { int rc = ntt_mont(F, hb, log_n, omi); if (rc) return rc; }
    {
        u64 p[4]; memcpy(p, ninv, 32);
        for (u64 i = 0; i < N; ++i) { fr_mul(F, hb + 4 * i, hb + 4 * i, p); fr_mul(F, p, p, gi); }
    }
    /* delta terms: a' = a + d1*T, b' = b + d2*T, c' = c + d3*T,
     * h' = h + d1*b + d2*a + d1*d2*T - d3  (T = X^N - 1) */
    u64 d[3][4], hN[4] = {0, 0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        if (ge4(delta + 4 * k, F->p)) return -2;
        fr_to_mont(F, d[k], delta + 4 * k);
    }
    u64 d12[4]; fr_mul(F, d12, d[0], d[1]);
    for (u64 i = 0; i < N; ++i) {
        u64 am[4], bm[4], t[4];
        fr_to_mont(F, am, a_out + 4 * i); fr_to_mont(F, bm, b_out + 4 * i);
        fr_mul(F, t, d[0], bm); fr_add(F, hb + 4 * i, hb + 4 * i, t);
        fr_mul(F, t, d[1], am); fr_add(F, hb + 4 * i, hb + 4 * i, t);
    }
    fr_sub(F, hb, hb, d12);   /* -d1*d2 at X^0 */
    fr_sub(F, hb, hb, d[2]);  /* -d3 */
    memcpy(hN, d12, 32);      /* +d1*d2 at X^N */
    for (u64 i = 0; i < N; ++i) fr_from_mont(F, h_out + 4 * i, hb + 4 * i);
    fr_from_mont(F, h_out + 4 * N, hN);
    for (int k = 0; k < 3; ++k) {
        u64 x0[4], xN[4], t0[4];
        fr_to_mont(F, t0, outs[k]);
        fr_sub(F, x0, t0, d[k]);           /* X^0: -delta */
        fr_from_mont(F, outs[k], x0);
        memcpy(xN, d[k], 32);              /* X^N: +delta */
        fr_from_mont(F, outs[k] + 4 * N, xN);
    }
    free(hb);
    for (int k = 0; k < 3; ++k) free(buf[k]);
    if (divisible) *divisible = all_ok;
    return 0;
}

/* ---- reference-SHAPED divisibility check for small n (src/QAP.hs:325-327) ----------------------
 * p = a*b - c by schoolbook product, then long division by the monic target t (degree nt).
 * a,b,c: na,nb,nc canonical coefficients; q_out: max(0, na+nb-1-nt) coeffs... caller sizes
 * q_out >= na+nb, rem_out >= nt.  Returns *rem_is_zero. */
int orc_poly_mul_divmod_check(int field_id, const u64* a, u64 na, const u64* b, u64 nb,
                              const u64* c, u64 nc, const u64* t, u64 nt_plus1,
                              u64* q_out, u64* n_q, u64* rem_out, int* rem_is_zero) {
    const orc_field* F = field_get(field_id);
    if (!F || nt_plus1 < 1) return -1;
    u64 np = (na && nb) ? na + nb - 1 : 0;
    if (nc > np) np = nc;
    u64* p = (u64*)calloc((size_t)(np ? np : 1) * 4, 8);
    u64* am = (u64*)malloc((size_t)(na ? na : 1) * 32);
    u64* bm = (u64*)malloc((size_t)(nb ? nb : 1) * 32);
    u64* tm = (u64*)malloc((size_t)nt_plus1 * 32);
    if (!p || !am || !bm || !tm) return -3;
    for (u64 i = 0; i < na; ++i) fr_to_mont(F, am + 4 * i, a + 4 * i);
    for (u64 i = 0; i < nb; ++i) fr_to_mont(F, bm + 4 * i, b + 4 * i);
    for (u64 i = 0; i < nt_plus1; ++i) fr_to_mont(F, tm + 4 * i, t + 4 * i);
    for (u64 i = 0; i < na; ++i)
        for (u64 j = 0; j < nb; ++j) {
            u64 x[4];
            fr_mul(F, x, am + 4 * i, bm + 4 * j);
            fr_add(F, p + 4 * (i + j), p + 4 * (i + j), x);
        }
    for (u64 i = 0; i < nc; ++i) { u64 x[4]; fr_to_mont(F, x, c + 4 * i); fr_sub(F, p + 4 * i, p + 4 * i, x); }
    u64 nt = nt_plus1 - 1;   /* degree of t */
    u64 lead_inv[4];
    fr_inv(F, lead_inv, tm + 4 * nt);
    u64 nq = np > nt ? np - nt : 0;
    for (u64 kk = nq; kk-- > 0;) {
        u64 cq[4];
        fr_mul(F, cq, p + 4 * (kk + nt), lead_inv);
        if (q_out) fr_from_mont(F, q_out + 4 * kk, cq);
        if (!is_zero4(cq))
            for (u64 j = 0; j <= nt; ++j) {
                u64 x[4];
                fr_mul(F, x, cq, tm + 4 * j);
                fr_sub(F, p + 4 * (kk + j), p + 4 * (kk + j), x);
            }
    }
    int zero = 1;
    for (u64 i = 0; i < nt && i < np; ++i) {
        if (!is_zero4(p + 4 * i)) zero = 0;
        if (rem_out) fr_from_mont(F, rem_out + 4 * i, p + 4 * i);
    }
    if (n_q) *n_q = nq;
    if (rem_is_zero) *rem_is_zero = zero;
    free(p); free(am); free(bm); free(tm);
    return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
