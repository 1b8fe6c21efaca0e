"""ctypes loader for the C oracle (oracle/r1cs_oracle.c).  TEST INFRASTRUCTURE ONLY -- see the header
of r1cs_oracle.c for who may import this.  Field elements cross as numpy uint64 arrays of shape
(n, 4): little-endian limbs, canonical."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


class OrcCsr(C.Structure):
    _fields_ = [("rowptr", u32p), ("col", u32p), ("val", u64p), ("nnz", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "r1cs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ})
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p64(a):
    return a.ctypes.data_as(u64p) if a is not None else None


def _p32(a):
    return a.ctypes.data_as(u32p)


def ints_to_limbs(vals) -> np.ndarray:
    """list of Python ints -> (n, 4) uint64"""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    m = (1 << 64) - 1
    for i, v in enumerate(vals):
        out[i, 0] = v & m
        out[i, 1] = (v >> 64) & m
        out[i, 2] = (v >> 128) & m
        out[i, 3] = (v >> 192) & m
    return out


def limbs_to_ints(a: np.ndarray):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def _csr(rowptr, col, val):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.uint32)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    val = np.ascontiguousarray(val, dtype=np.uint64).reshape(-1, 4)
    s = OrcCsr(_p32(rowptr), _p32(col), _p64(val), len(col))
    s._keep = (rowptr, col, val)
    return s


def field_constants(field_id: int):
    p = np.zeros(4, np.uint64); one = np.zeros(4, np.uint64); r2 = np.zeros(4, np.uint64)
    ninv = C.c_uint64()
    rc = lib().orc_field_constants(field_id, _p64(p), _p64(one), _p64(r2), C.byref(ninv))
    assert rc == 0
    return limbs_to_ints(p)[0], limbs_to_ints(one)[0], limbs_to_ints(r2)[0], ninv.value


def fr_binop(field_id: int, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, np.uint64).reshape(-1, 4)
    o = np.empty_like(a)
    rc = lib().orc_fr_binop(field_id, op, _p64(a), _p64(b), _p64(o), C.c_uint64(a.shape[0]))
    if rc:
        raise ValueError("orc_fr_binop rc=%d" % rc)
    return o


def root_of_unity(field_id: int, k: int) -> int:
    o = np.zeros(4, np.uint64)
    rc = lib().orc_root_of_unity(field_id, k, _p64(o))
    if rc:
        raise ValueError("orc_root_of_unity rc=%d" % rc)
    return limbs_to_ints(o)[0]


def r1cs_eval_check(field_id, n_rows, n_cols, A, B, Cm, w, want_vectors=False, n_threads=1):
    """A, B, Cm: (rowptr, col, val) numpy triples.  Returns dict(n_violations, first_bad_row[, Aw, Bw, Cw])."""
    a, b, c = _csr(*A), _csr(*B), _csr(*Cm)
    w = np.ascontiguousarray(w, np.uint64).reshape(-1, 4)
    assert w.shape[0] >= n_cols
    Aw = Bw = Cw = None
    if want_vectors:
        Aw = np.empty((n_rows, 4), np.uint64); Bw = np.empty_like(Aw); Cw = np.empty_like(Aw)
    nv, fb = C.c_uint64(), C.c_uint64()
    rc = lib().orc_r1cs_eval_check(field_id, C.c_uint32(n_rows), C.c_uint32(n_cols), C.byref(a), C.byref(b),
                                   C.byref(c), _p64(w), _p64(Aw), _p64(Bw), _p64(Cw), C.byref(nv), C.byref(fb),
                                   n_threads)
    if rc:
        raise ValueError("orc_r1cs_eval_check rc=%d" % rc)
    out = {"n_violations": nv.value, "first_bad_row": (-1 if fb.value == 2**64 - 1 else fb.value)}
    if want_vectors:
        out.update(Aw=Aw, Bw=Bw, Cw=Cw)
    return out


def ntt(field_id: int, data: np.ndarray, inverse: bool, n_threads: int = 1) -> np.ndarray:
    d = np.array(data, dtype=np.uint64, copy=True).reshape(-1, 4)
    n = d.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().orc_ntt(field_id, _p64(d), log_n, int(bool(inverse)), n_threads)
    if rc:
        raise ValueError("orc_ntt rc=%d" % rc)
    return d


def qap_witness(field_id, aw, bw, cw, delta=(0, 0, 0), n_threads=1):
    aw = np.ascontiguousarray(aw, np.uint64).reshape(-1, 4)
    bw = np.ascontiguousarray(bw, np.uint64).reshape(-1, 4)
    cw = np.ascontiguousarray(cw, np.uint64).reshape(-1, 4)
    N = aw.shape[0]
    log_n = N.bit_length() - 1
    assert 1 << log_n == N
    d = ints_to_limbs(list(delta))
    a = np.empty((N + 1, 4), np.uint64); b = np.empty_like(a); c = np.empty_like(a); h = np.empty_like(a)
    div = C.c_int()
    rc = lib().orc_qap_witness(field_id, log_n, _p64(aw), _p64(bw), _p64(cw), _p64(d), _p64(a), _p64(b), _p64(c),
                               _p64(h), C.byref(div), n_threads)
    if rc:
        raise ValueError("orc_qap_witness rc=%d" % rc)
    return a, b, c, h, bool(div.value)


def poly_mul_divmod_check(field_id, a, b, c, t):
    """reference-shaped: (q, rem, rem_is_zero) of (a*b - c) divmod t; polys as lists of ints."""
    al, bl, cl, tl = (ints_to_limbs(x) if len(x) else np.zeros((0, 4), np.uint64) for x in (a, b, c, t))
    q = np.zeros((len(a) + len(b) + len(c) + 1, 4), np.uint64)
    rem = np.zeros((max(1, len(t)), 4), np.uint64)
    nq, z = C.c_uint64(), C.c_int()
    rc = lib().orc_poly_mul_divmod_check(field_id, _p64(al), C.c_uint64(len(a)), _p64(bl), C.c_uint64(len(b)),
                                         _p64(cl), C.c_uint64(len(c)), _p64(tl), C.c_uint64(len(t)), _p64(q),
                                         C.byref(nq), _p64(rem), C.byref(z))
    if rc:
        raise ValueError("orc_poly_mul_divmod_check rc=%d" % rc)
    return limbs_to_ints(q[:nq.value]), limbs_to_ints(rem[:len(t) - 1]), bool(z.value)


def max_threads() -> int:
    return lib().orc_max_threads()
