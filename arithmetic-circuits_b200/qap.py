"""Host-side mirror of the reference's QAP / Circuit API for the accelerated path (Python binding of the
C ABI; the Haskell binding with the same shape is hs/QAP/GPU.hs).  Names follow the reference
(src/QAP.hs:11-39, src/Circuit/Arithmetic.hs:7-20) in snake_case:

    reference                      here
    ---------------------------    ---------------------------------------------
    ArithCircuit [Gate Wire f]     ArithCircuit([Mul(..), Equal(..), Split(..)])
    generateAssignment c inputs    generate_assignment(c, inputs) -> QapSet
    arithCircuitToGenQAP roots c   arith_circuit_to_gen_qap(c, roots) -> GenQAP (CSR rows per root)
    verifyAssignment qap asg       verify_assignment(ctx, gen_qap, asg) -> bool          [GPU]
    verificationWitness[Zk]        verification_witness_zk(ctx, d1,d2,d3, gen_qap, asg)  [GPU]
    createPolynomialsFFT           create_polynomials_fft(ctx, columns)                   [GPU]
    createPolynomials (Lagrange)   create_polynomials(ctx, xs, ys)                        [GPU]
    qapSetToMap                    QapSet.to_vector(layout)

Every numeric step runs in libacg.so (CUDA for bulk arithmetic, C++ for per-gate host logic).  Field
elements are Python ints at this level and (n, 4) uint64 little-endian limb arrays underneath."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import AcgCsr, AcgTiming

BN254_FR = 0
BLS12_381_FR = 1
CHECK_AUTO, CHECK_ROWWISE, CHECK_TILED = 0, 1, 2
UINT64_MAX = (1 << 64) - 1
_M64 = UINT64_MAX


class AcgError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = _lib.lib().acg_strerror(code).decode()
        super().__init__("%s (%d)%s" % (msg, code, (": " + detail) if detail else ""))


def _check(rc: int, ctx: Optional["Context"] = None):
    if rc != 0:
        detail = ""
        if ctx is not None and ctx._h:
            detail = _lib.lib().acg_last_error(ctx._h).decode()
        raise AcgError(rc, detail)


# ---------------------------------------------------------------------------------------------------
# limb helpers
# ---------------------------------------------------------------------------------------------------
def to_limbs(vals: Iterable[int]) -> np.ndarray:
    vals = list(vals)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i, 0] = v & _M64
        out[i, 1] = (v >> 64) & _M64
        out[i, 2] = (v >> 128) & _M64
        out[i, 3] = (v >> 192) & _M64
    return out


def from_limbs(a: np.ndarray) -> List[int]:
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def field_constants(field: int) -> Dict[str, int]:
    p = np.zeros(4, np.uint64); r = np.zeros(4, np.uint64); r2 = np.zeros(4, np.uint64)
    ninv = C.c_uint64(); ta = C.c_uint32()
    _check(_lib.lib().acg_field_constants(field, p.ctypes.data_as(_lib.u64p), r.ctypes.data_as(_lib.u64p),
                                          r2.ctypes.data_as(_lib.u64p), C.byref(ninv), C.byref(ta)))
    return {"modulus": from_limbs(p)[0], "mont_r": from_limbs(r)[0], "mont_r2": from_limbs(r2)[0],
            "ninv64": ninv.value, "two_adicity": ta.value}


def get_root_of_unity(field: int, k: int) -> int:
    """getRootOfUnity k (pairing-1.0.0; Example.hs:26)."""
    o = np.zeros(4, np.uint64)
    _check(_lib.lib().acg_root_of_unity(field, k, o.ctypes.data_as(_lib.u64p)))
    return from_limbs(o)[0]


# ---------------------------------------------------------------------------------------------------
# circuit IR (word-stream marshalling, include/acg.h)
# ---------------------------------------------------------------------------------------------------
def InputWire(i: int) -> int:
    return (0 << 62) | i


def IntermediateWire(i: int) -> int:
    return (1 << 62) | i


def OutputWire(i: int) -> int:
    return (2 << 62) | i


def Var(w):
    return ("var", w)


def ConstGate(f):
    return ("const", f)


def Add(l, r):
    return ("add", l, r)


def ScalarMul(s, c):
    return ("scalar", s, c)


def Mul(l, r, out):
    return ("mul", l, r, out)


def Equal(i, m, out):
    return ("equal", i, m, out)


def Split(i, outs):
    return ("split", i, list(outs))


def _limbs4(v: int) -> List[int]:
    return [v & _M64, (v >> 64) & _M64, (v >> 128) & _M64, (v >> 192) & _M64]


def _affine_words(c, modulus: int) -> List[int]:
    out: List[int] = []
    stack = [(c, False)]
    while stack:  # iterative post-order (unsplit chains are 256 deep)
        node, done = stack.pop()
        tag = node[0]
        if tag == "var":
            out += [0, node[1]]
        elif tag == "const":
            out += [1] + _limbs4(node[1] % modulus)
        elif tag == "add":
            if done:
                out.append(2)
            else:
                stack += [(node, True), (node[2], False), (node[1], False)]
        elif tag == "scalar":
            if done:
                out += [3] + _limbs4(node[1] % modulus)
            else:
                stack += [(node, True), (node[2], False)]
        else:
            raise ValueError("bad affine node %r" % (tag,))
    return out


class ArithCircuit:
    """ArithCircuit f = [Gate Wire f] (src/Circuit/Arithmetic.hs:149-150), parsed by the C++ host side."""

    def __init__(self, field: int, gates: Sequence):
        self.field = field
        self.gates = list(gates)
        modulus = field_constants(field)["modulus"]
        words: List[int] = []
        for g in self.gates:
            if g[0] == "mul":
                l, r = _affine_words(g[1], modulus), _affine_words(g[2], modulus)
                words += [1, g[3], len(l)] + l + [len(r)] + r
            elif g[0] == "equal":
                words += [2, g[1], g[2], g[3]]
            elif g[0] == "split":
                words += [3, g[1], len(g[2])] + list(g[2])
            else:
                raise ValueError("bad gate %r" % (g[0],))
        self.words = np.array(words, dtype=np.uint64)
        h = C.c_void_p()
        _check(_lib.lib().acg_circuit_parse(field, _ptr(self.words), len(words), C.byref(h)))
        self._h = h

    @classmethod
    def from_words(cls, field: int, words: np.ndarray) -> "ArithCircuit":
        self = cls.__new__(cls)
        self.field = field
        self.gates = None
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        h = C.c_void_p()
        _check(_lib.lib().acg_circuit_parse(field, _ptr(self.words), len(self.words), C.byref(h)))
        self._h = h
        return self

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:  # (_lib is None while the interpreter shuts down)
            _lib.lib().acg_circuit_free(self._h)
            self._h = None

    @property
    def num_gates(self) -> int:
        return _lib.lib().acg_circuit_num_gates(self._h)

    @property
    def num_roots(self) -> int:
        return _lib.lib().acg_circuit_num_roots(self._h)

    def plan_stats(self) -> Tuple[int, int]:
        """(dependency levels, widest level) of the device witness generation; raises AcgError(-6) for a gate list that
        is not in single-assignment, define-before-use form."""
        a, b = C.c_uint32(), C.c_uint32()
        _check(_lib.lib().acg_circuit_plan_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def valid(self) -> bool:
        """validArithCircuit (src/Circuit/Arithmetic.hs:158-185)."""
        return bool(_lib.lib().acg_circuit_valid(self._h))


def unsplit(wires: Sequence[int]):
    """unsplit (src/Circuit/Arithmetic.hs:238-244)."""
    acc = ConstGate(0)
    for ix, w in enumerate(wires):
        acc = Add(acc, ScalarMul(1 << ix, Var(w)))
    return acc


class QapSet:
    """QapSet f of witness values (src/QAP.hs:66-71); constant = 1."""

    def __init__(self, field: int, handle):
        self.field = field
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.lib().acg_assignment_free(self._h)
            self._h = None

    def dims(self) -> Tuple[int, int, int]:
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(_lib.lib().acg_assignment_dims(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def lookup(self, wire: int) -> Optional[int]:
        """lookupAtWire (src/QAP.hs:331-337)."""
        o = np.zeros(4, np.uint64)
        rc = _lib.lib().acg_assignment_lookup(self._h, wire, o.ctypes.data_as(_lib.u64p))
        if rc < 0:
            _check(rc)
        return from_limbs(o)[0] if rc == 1 else None

    def update(self, wire: int, value: int) -> "QapSet":
        """updateAtWire (src/QAP.hs:341-347), in place."""
        v = to_limbs([value])
        _check(_lib.lib().acg_assignment_update(self._h, wire, v.ctypes.data_as(_lib.u64p)))
        return self

    def to_vector(self, layout: Optional[Tuple[int, int, int]] = None) -> np.ndarray:
        """qapSetToMap (src/QAP.hs:605-620) as a dense (n_cols, 4) limb array."""
        n_in, n_mid, n_out = layout if layout is not None else self.dims()
        w = np.zeros((1 + n_in + n_mid + n_out, 4), np.uint64)
        _check(_lib.lib().acg_assignment_to_vector(self._h, n_in, n_mid, n_out, _ptr(w)))
        return w


def generate_assignment(circuit: ArithCircuit, inputs: Dict[int, int]) -> QapSet:
    """generateAssignment (src/QAP.hs:597-603)."""
    ix = np.array(sorted(inputs), dtype=np.uint32)
    vals = to_limbs([inputs[int(i)] for i in ix])
    h = C.c_void_p()
    _check(_lib.lib().acg_generate_assignment(circuit._h, _ptr(ix), _ptr(vals), len(ix), C.byref(h)))
    return QapSet(circuit.field, h)


class GenQAP:
    """The R1CS rows of GenQAP (Map k) k (src/QAP.hs:94-99): three host CSR matrices, one row per root in
    ascending-root order.  Owns (or borrows) the arrays."""

    def __init__(self, field: int, n_rows: int, n_cols: int, layout: Tuple[int, int, int], mats, roots=None,
                 owner=None):
        self.field = field
        self.n_rows, self.n_cols = n_rows, n_cols
        self.layout = layout
        self.mats = mats          # [(rowptr u32, col u32, val (nnz,4) u64)] * 3
        self.roots = roots
        self._owner = owner       # keeps the C++ object (and so the borrowed arrays) alive

    def csr_structs(self):
        out = []
        for rowptr, col, val in self.mats:
            out.append(AcgCsr(rowptr.ctypes.data_as(_lib.u32p), col.ctypes.data_as(_lib.u32p),
                              val.ctypes.data_as(_lib.u64p), len(col)))
        return out

    @property
    def nnz(self) -> Tuple[int, int, int]:
        return tuple(len(m[1]) for m in self.mats)

    def tile_stream_digest(self, variant: int = 0, n_threads: int = 0, rows: Optional[Tuple[int, int]] = None):
        """Host-only: digests of the tile stream acg_r1cs_upload would build for this system (acg_tile_stream_digest):
        (hash of the blobs, hash of the records, blob bytes, tiles).  Independent of n_threads by construction."""
        a, b, c = self.csr_structs()
        out = (C.c_uint64 * 4)()
        r0, r1 = rows if rows is not None else (0, self.n_rows)
        _check(_lib.lib().acg_tile_stream_digest(self.field, variant, self.n_rows, self.n_cols, C.byref(a), C.byref(b),
                                                 C.byref(c), r0, r1, n_threads, out), None)
        return tuple(int(x) for x in out)


class _HostR1cs:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.lib().acg_r1cs_host_free(self._h)
            self._h = None


def _gen_qap_from_handle(field: int, h) -> GenQAP:
    L = _lib.lib()
    owner = _HostR1cs(h)
    n_rows, n_cols, a, b, c = (C.c_uint32() for _ in range(5))
    _check(L.acg_r1cs_host_dims(h, C.byref(n_rows), C.byref(n_cols), C.byref(a), C.byref(b), C.byref(c)))
    mats = []
    for k in range(3):
        s = AcgCsr()
        _check(L.acg_r1cs_host_csr(h, k, C.byref(s)))
        nnz = int(s.nnz)
        rowptr = np.ctypeslib.as_array(s.rowptr, shape=(n_rows.value + 1,))
        col = np.ctypeslib.as_array(s.col, shape=(nnz,)) if nnz else np.zeros(0, np.uint32)
        val = np.ctypeslib.as_array(s.val, shape=(nnz, 4)) if nnz else np.zeros((0, 4), np.uint64)
        mats.append((rowptr, col, val))
    rp = L.acg_r1cs_host_roots(h)
    roots = np.ctypeslib.as_array(rp, shape=(n_rows.value, 4)) if n_rows.value else np.zeros((0, 4), np.uint64)
    return GenQAP(field, n_rows.value, n_cols.value, (a.value, b.value, c.value), mats, roots, owner)


def arith_circuit_to_gen_qap(circuit: ArithCircuit, roots: Optional[Sequence[Sequence[int]]] = None,
                             root_start: int = 0, layout: Tuple[int, int, int] = (0, 0, 0)) -> GenQAP:
    """arithCircuitToGenQAP (src/QAP.hs:530-539) as sparse rows.  roots: [[k]] per gate as in the
    reference, or None for `fromIntegral <$> fresh` counting from root_start."""
    flat = None
    if roots is not None:
        flat_list = [r for per_gate in roots for r in per_gate]
        if len(flat_list) != circuit.num_roots:
            raise AcgError(-1, "gateToGenQAP: wrong number of roots supplied")
        modulus = field_constants(circuit.field)["modulus"]
        flat = to_limbs([r % modulus for r in flat_list])
    h = C.c_void_p()
    _check(_lib.lib().acg_circuit_to_r1cs(circuit._h, _ptr(flat), root_start, layout[0], layout[1], layout[2],
                                          C.byref(h)))
    return _gen_qap_from_handle(circuit.field, h)


def synth_r1cs(field: int, n: int, seed: int, dense: bool = False,
               rows: Optional[Tuple[int, int]] = None) -> Tuple[GenQAP, np.ndarray]:
    """S(n, seed, field) of SURVEY.md 8(d): (rows, honest witness vector).  rows=(begin, end) materialises
    only that row shard (GenQAP.n_rows = end - begin); the witness is always complete."""
    h = C.c_void_p()
    wp = _lib.u64p()
    rb, re = rows if rows is not None else (0, n)
    _check(_lib.lib().acg_synth_r1cs_rows(field, n, seed, int(dense), rb, re, C.byref(h), C.byref(wp)))
    g = _gen_qap_from_handle(field, h)
    w = np.ctypeslib.as_array(wp, shape=(g.n_cols, 4)).copy()
    _lib.lib().acg_free(wp)
    return g, w


def synth_circuit(field: int, n: int, seed: int, dense: bool = False) -> Tuple[ArithCircuit, Dict[int, int]]:
    """The same family as a reference-style ArithCircuit plus its inputs (small n; cross-checks)."""
    L = _lib.lib()
    wp, ivp, ixp = _lib.u64p(), _lib.u64p(), _lib.u32p()
    nw, ni = C.c_uint64(), C.c_uint32()
    _check(L.acg_synth_circuit_words(field, n, seed, int(dense), C.byref(wp), C.byref(nw), C.byref(ixp),
                                     C.byref(ivp), C.byref(ni)))
    words = np.ctypeslib.as_array(wp, shape=(nw.value,)).copy()
    ix = np.ctypeslib.as_array(ixp, shape=(ni.value,)).copy()
    iv = np.ctypeslib.as_array(ivp, shape=(ni.value, 4)).copy()
    for p in (wp, ivp, ixp):
        L.acg_free(p)
    return ArithCircuit.from_words(field, words), dict(zip((int(i) for i in ix), from_limbs(iv)))


def synth_mixed_circuit(field: int, n_rows: int, seed: int, dist: Tuple[int, int, int] = (50, 10, 1),
                        num_inputs: int = 1024, split_bits: int = 256) -> Tuple[ArithCircuit, Dict[int, int]]:
    """A random circuit with the gate mix of the reference's own generator, arbArithCircuit
    (test/Test/Circuit/Arithmetic.hs:77-126): Mul : Equal : Split drawn with frequencies `dist` (the reference's
    property tests use 50 : 10 : 1), Split into 256 bits, gate inputs affine circuits of size 1 over the inputs and ALL
    earlier intermediate wires (`elements mids`: no locality), every gate output a fresh intermediate wire.  Gates are
    added until the circuit lowers to at least n_rows R1CS rows (Mul 1 row, Equal 2, Split 1 + 256).  Returns the
    circuit and an input assignment."""
    import random
    rnd = random.Random(seed)
    r = field_constants(field)["modulus"]
    words: List[int] = []
    n_mid, rows = 0, 0

    def leaf() -> List[int]:
        k = rnd.randrange(3 if n_mid else 2)
        if k == 0:
            return [1] + _limbs4(rnd.randrange(r))
        if k == 1:
            return [0, InputWire(rnd.randrange(num_inputs))]
        return [0, IntermediateWire(rnd.randrange(n_mid))]

    def affine() -> List[int]:   # arbAffineCircuitWithMids numInps mids 1
        if rnd.randrange(2) == 0:
            return leaf() + [3] + _limbs4(rnd.randrange(r))   # ScalarMul s leaf
        return leaf() + leaf() + [2]                              # Add leaf leaf

    total = sum(dist)
    while rows < n_rows:
        x = rnd.randrange(total) if n_mid else 0
        if x < dist[0]:
            l, rr = affine(), affine()
            words += [1, IntermediateWire(n_mid), len(l)] + l + [len(rr)] + rr
            n_mid += 1
            rows += 1
        elif x < dist[0] + dist[1]:
            words += [2, IntermediateWire(rnd.randrange(n_mid)), IntermediateWire(n_mid), IntermediateWire(n_mid + 1)]
            n_mid += 2
            rows += 2
        else:
            words += [3, IntermediateWire(rnd.randrange(n_mid)), split_bits] + \
                     [IntermediateWire(n_mid + i) for i in range(split_bits)]
            n_mid += split_bits
            rows += 1 + split_bits
    inputs = {i: rnd.randrange(r) for i in range(num_inputs)}
    return ArithCircuit.from_words(field, np.array(words, dtype=np.uint64)), inputs


def synth_mixed_r1cs(field: int, n_rows: int, seed: int, dist: Tuple[int, int, int] = (50, 10, 1)):
    """synth_mixed_circuit lowered through the host mirror (arithCircuitToGenQAP with roots 1, 2, 3, ... and
    generateAssignment): (rows, honest witness vector)."""
    circuit, inputs = synth_mixed_circuit(field, n_rows, seed, dist)
    g = arith_circuit_to_gen_qap(circuit, None, 1)
    w = witness_vector(generate_assignment(circuit, inputs), g.layout)
    g._circuit = circuit
    return g, w


# ---------------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------------
class Context:
    """One CUDA device + one field (acg_ctx).  Raises AcgError(ACG_ERR_NO_DEVICE) without a GPU."""

    def __init__(self, field: int = BN254_FR, device: int = 0):
        self.field = field
        self.device = device
        self._h = C.c_void_p()
        _check(_lib.lib().acg_ctx_create(field, device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().acg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        if _lib is not None:
            self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_check_kernel(self, which: int):
        _check(_lib.lib().acg_ctx_set_check_kernel(self._h, which), self)

    def set_overlap_checks(self, on: bool):
        """Overlap of back-to-back checks of the same system (include/acg.h); off by default."""
        _check(_lib.lib().acg_ctx_set_overlap_checks(self._h, int(on)), self)

    def set_tiled_variant(self, variant: int):
        """Tile geometry for systems uploaded AFTER this call (0: 128-row tiles, 1: 256-row tiles)."""
        _check(_lib.lib().acg_ctx_set_tiled_variant(self._h, variant), self)


    def last_timing(self) -> Dict[str, float]:
        t = AcgTiming()
        _check(_lib.lib().acg_last_timing(self._h, C.byref(t)), self)
        return {"h2d_ms": t.h2d_ms, "kernel_ms": t.kernel_ms, "d2h_ms": t.d2h_ms, "kernel_launches": t.kernel_launches}

    def kernel_launch_count(self) -> int:
        return _lib.lib().acg_kernel_launch_count(self._h)

    def profile_begin(self, max_launches: int):
        _check(_lib.lib().acg_profile_begin(self._h, max_launches), self)

    def profile_end(self, capacity: int) -> List[float]:
        ms = (C.c_float * max(1, capacity))()
        n = C.c_uint32()
        _check(_lib.lib().acg_profile_end(self._h, ms, capacity, C.byref(n)), self)
        return [ms[i] for i in range(n.value)]

    # ---- uploads
    def upload_r1cs(self, g: GenQAP, row_begin: int = 0, row_end: Optional[int] = None) -> "DeviceR1cs":
        row_end = g.n_rows if row_end is None else row_end
        a, b, c = g.csr_structs()
        h = C.c_void_p()
        _check(_lib.lib().acg_r1cs_upload(self._h, g.n_rows, g.n_cols, C.byref(a), C.byref(b), C.byref(c), row_begin,
                                          row_end, C.byref(h)), self)
        return DeviceR1cs(self, h, g, row_begin, row_end)

    def upload_witness(self, w: np.ndarray) -> "DeviceVec":
        w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
        h = C.c_void_p()
        _check(_lib.lib().acg_witness_upload(self._h, _ptr(w), w.shape[0], C.byref(h)), self)
        return DeviceVec(self, h)

    # ---- R1CS check
    def generate_assignment(self, circuit: ArithCircuit, inputs: Dict[int, int],
                            layout: Tuple[int, int, int] = (0, 0, 0)) -> Tuple["DeviceVec", int]:
        """generateAssignment (src/QAP.hs:597-603) on the device, level by level (K6).  Returns the device witness
        (qapSetToMap order) and the number of dependency levels."""
        ix = np.array(sorted(inputs), dtype=np.uint32)
        vals = to_limbs([inputs[int(i)] for i in ix])
        h, lv = C.c_void_p(), C.c_uint32()
        _check(_lib.lib().acg_generate_assignment_device(self._h, circuit._h, _ptr(ix), _ptr(vals), len(ix), layout[0],
                                                         layout[1], layout[2], C.byref(h), C.byref(lv)), self)
        return DeviceVec(self, h), lv.value

    def r1cs_check(self, m: "DeviceR1cs", w: "DeviceVec") -> Tuple[int, int]:
        """(number of violated rows, first violated global row or -1)."""
        nv, fb = C.c_uint64(), C.c_uint64()
        _check(_lib.lib().acg_r1cs_check(self._h, m._h, w._h, C.byref(nv), C.byref(fb)), self)
        return nv.value, (-1 if fb.value == UINT64_MAX else fb.value)

    def r1cs_check_async(self, m: "DeviceR1cs", w: "DeviceVec", d_result_ptr: int, stream: int):
        _check(_lib.lib().acg_r1cs_check_async(self._h, m._h, w._h, C.c_void_p(d_result_ptr), C.c_void_p(stream)), self)

    def r1cs_check_async_allreduce(self, m: "DeviceR1cs", w: "DeviceVec", peer: "PeerExchange", d_result_ptr: int,
                                   stream: int):
        """Shard check + all-reduce of the result pair over peer memory (acg_r1cs_check_async_allreduce)."""
        _check(_lib.lib().acg_r1cs_check_async_allreduce(self._h, m._h, w._h, peer._h, C.c_void_p(d_result_ptr),
                                                         C.c_void_p(stream)), self)

    def r1cs_check_host(self, g: GenQAP, w: np.ndarray) -> Tuple[int, int]:
        w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
        a, b, c = g.csr_structs()
        nv, fb = C.c_uint64(), C.c_uint64()
        _check(_lib.lib().acg_r1cs_check_host(self._h, g.n_rows, g.n_cols, C.byref(a), C.byref(b), C.byref(c),
                                              _ptr(w), C.byref(nv), C.byref(fb)), self)
        return nv.value, (-1 if fb.value == UINT64_MAX else fb.value)

    def r1cs_eval(self, m: "DeviceR1cs", w: "DeviceVec"):
        n = m.row_end - m.row_begin
        outs = [np.zeros((n, 4), np.uint64) for _ in range(3)]
        _check(_lib.lib().acg_r1cs_eval(self._h, m._h, w._h, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])), self)
        return outs

    # ---- NTT / interpolation
    def ntt(self, data: np.ndarray, inverse: bool = False) -> np.ndarray:
        d = np.array(data, dtype=np.uint64, copy=True).reshape(-1, 4)
        n = d.shape[0]
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise AcgError(-1, "NTT length must be a power of two")
        _check(_lib.lib().acg_ntt(self._h, _ptr(d), log_n, int(inverse)), self)
        return d

    def interpolate_columns(self, cols: np.ndarray) -> np.ndarray:
        """cols: (n_cols_batch, N, 4) values in ascending-root order, N a power of two."""
        d = np.array(cols, dtype=np.uint64, copy=True)
        nb, n = d.shape[0], d.shape[1]
        log_n = n.bit_length() - 1
        if (1 << log_n) != n:
            raise AcgError(-1, "column length must be a power of two")
        _check(_lib.lib().acg_interpolate_columns(self._h, _ptr(d), log_n, nb), self)
        return d

    def qap_witness(self, m: "DeviceR1cs", w: "DeviceVec", delta=(0, 0, 0), want=("a", "b", "c", "h")):
        N = 1
        while N < m.gen_qap.n_rows:
            N <<= 1
        bufs = {k: (np.zeros((N + 1, 4), np.uint64) if k in want else None) for k in ("a", "b", "c", "h")}
        d = to_limbs(list(delta))
        div = C.c_int()
        _check(_lib.lib().acg_qap_witness(self._h, m._h, w._h, _ptr(d), _ptr(bufs["a"]), _ptr(bufs["b"]),
                                          _ptr(bufs["c"]), _ptr(bufs["h"]), C.byref(div)), self)
        return bufs, bool(div.value)

    def lagrange(self, xs: Sequence[int], ys: Sequence[Sequence[int]], want_target: bool = True):
        n = len(xs)
        x = to_limbs(xs)
        y = to_limbs([v for row in ys for v in row]) if ys else np.zeros((0, 4), np.uint64)
        co = np.zeros((max(1, len(ys) * n), 4), np.uint64)
        tg = np.zeros((n + 1, 4), np.uint64) if want_target else None
        _check(_lib.lib().acg_lagrange(self._h, _ptr(x), _ptr(y), n, len(ys), _ptr(co), _ptr(tg)), self)
        polys = [from_limbs(co[i * n:(i + 1) * n]) for i in range(len(ys))]
        return polys, (from_limbs(tg) if want_target else None)

    def fft_target(self, n_roots: int) -> List[int]:
        """FFT.fftTargetPoly primRoots n (src/QAP.hs:524): prod_{i < n} (X - omega^i), stripped."""
        out = np.zeros((n_roots + 1, 4), np.uint64)
        _check(_lib.lib().acg_fft_target(self._h, n_roots, _ptr(out)), self)
        return strip(from_limbs(out))

    def fr_binop(self, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, np.uint64).reshape(-1, 4)
        o = np.empty_like(a)
        _check(_lib.lib().acg_fr_binop(self._h, op, _ptr(a), _ptr(b), _ptr(o), a.shape[0]), self)
        return o


class PeerExchange:
    """Exchange buffers of a group of row-shard ranks (include/acg.h "multi-GPU row shards").  `all_gather_bytes` is a
    callable taking this rank's 64-byte handle (bytes) and returning the list of all ranks' handles in rank order --
    sharding.connect_peers() supplies one over torch.distributed."""

    def __init__(self, ctx: Context, world: int, rank: int, all_gather_bytes):
        self.ctx = ctx
        h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _check(_lib.lib().acg_peer_create(ctx._h, world, rank, C.byref(h), handle), ctx)
        self._h = h
        handles = all_gather_bytes(bytes(handle))
        if len(handles) != world or any(len(x) != 64 for x in handles):
            raise AcgError(-1, "PeerExchange: need one 64-byte handle per rank")
        blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
        _check(_lib.lib().acg_peer_connect(ctx._h, self._h, blob), ctx)

    def free(self):
        if getattr(self, "_h", None) and self.ctx._h:
            _lib.lib().acg_peer_free(self._h)
        self._h = None

    def __del__(self):
        if _lib is not None:
            self.free()


class DeviceR1cs:
    def __init__(self, ctx: Context, h, g: GenQAP, row_begin: int, row_end: int):
        self.ctx, self._h, self.gen_qap = ctx, h, g
        self.row_begin, self.row_end = row_begin, row_end

    def free(self):
        if getattr(self, "_h", None) and self.ctx._h:
            _lib.lib().acg_r1cs_free(self._h)
        self._h = None

    def __del__(self):
        if _lib is not None:
            self.free()

    def set_row_offset(self, offset: int):
        """Global index of this row block's first row (acg_r1cs_set_row_offset): added to reported first bad rows."""
        _check(_lib.lib().acg_r1cs_set_row_offset(self._h, offset), self.ctx)

    @property
    def algorithmic_bytes(self) -> int:
        return _lib.lib().acg_r1cs_algorithmic_bytes(self._h)

    @property
    def stream_bytes(self) -> int:
        return _lib.lib().acg_r1cs_stream_bytes(self._h)


class DeviceVec:
    def __init__(self, ctx: Context, h):
        self.ctx, self._h = ctx, h

    def free(self):
        if getattr(self, "_h", None) and self.ctx._h:
            _lib.lib().acg_vec_free(self._h)
        self._h = None

    def __del__(self):
        if _lib is not None:
            self.free()

    def __len__(self):
        return _lib.lib().acg_vec_len(self._h)

    @property
    def device_ptr(self) -> int:
        return _lib.lib().acg_vec_device_ptr(self._h)

    def update_range(self, w_slice: np.ndarray, first: int):
        """Overwrite elements [first, first + len(w_slice)) from host memory (canonical limbs)."""
        w_slice = np.ascontiguousarray(w_slice, np.uint64).reshape(-1, 4)
        _check(_lib.lib().acg_witness_update_range(self.ctx._h, self._h, _ptr(w_slice), first, len(w_slice)), self.ctx)

    def update_async(self, w: np.ndarray, first: int = 0):
        """Enqueue an update of elements [first, first + len(w)) on the context's copy stream (acg_witness_update_async):
        no host synchronisation; later checks of this vector wait for it.  `w` must stay alive (ideally pinned) until
        the next blocking call on this vector."""
        w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
        keep = getattr(self, "_async_src", [])   # keep the host buffers alive while the copies are in flight
        self._async_src = (keep + [w])[-8:]
        _check(_lib.lib().acg_witness_update_async(self.ctx._h, self._h, _ptr(w), first, w.shape[0]), self.ctx)

    def stream_wait(self, stream: int):
        """Make the CUDA stream `stream` wait for the asynchronous updates enqueued so far (acg_vec_stream_wait)."""
        _check(_lib.lib().acg_vec_stream_wait(self.ctx._h, self._h, C.c_void_p(stream)), self.ctx)

    def status(self):
        """Wait for the asynchronous updates; raises AcgError(-2) if one of them met an element >= r."""
        _check(_lib.lib().acg_vec_status(self.ctx._h, self._h), self.ctx)

    def as_torch_bytes(self):
        """A torch uint8 tensor aliasing the device storage (Montgomery form, 32 bytes per element) -- for
        device-to-device exchange with torch.distributed.  The DeviceVec must outlive the tensor."""
        import torch

        class _Alias:
            pass
        a = _Alias()
        a.__cuda_array_interface__ = {"shape": (len(self) * 32,), "typestr": "|u1", "data": (int(self.device_ptr), False),
                                      "version": 2}
        return torch.as_tensor(a, device="cuda")

    def download(self) -> np.ndarray:
        """The vector as canonical limbs, shape (n, 4)."""
        out = np.zeros((len(self), 4), np.uint64)
        _check(_lib.lib().acg_vec_download(self.ctx._h, self._h, _ptr(out), len(self)), self.ctx)
        return out

    def update(self, w: np.ndarray):
        w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
        _check(_lib.lib().acg_witness_update(self.ctx._h, self._h, _ptr(w), w.shape[0]), self.ctx)


# ---------------------------------------------------------------------------------------------------
# reference-named entry points (GPU)
# ---------------------------------------------------------------------------------------------------
def strip(poly: Sequence[int]) -> List[int]:
    """VPoly normal form: no trailing zero coefficients (poly-0.4; matters for `== 0`, src/QAP.hs:310)."""
    p = list(poly)
    while p and p[-1] == 0:
        p.pop()
    return p


def witness_vector(assignment: QapSet, layout: Tuple[int, int, int]) -> np.ndarray:
    """The assignment as the dense witness of a system with the given layout.  The reference pairs QAP and assignment
    wire by wire with defaults on BOTH sides (combineWithDefaults, src/QAP.hs:163-181, 314): a wire the assignment
    lacks counts as 0, and a wire the system does not know meets the zero polynomial -- so assignment keys beyond the
    layout are dropped instead of being an error."""
    dims = assignment.dims()
    lay = tuple(max(a, b) for a, b in zip(dims, layout))
    w = assignment.to_vector(lay)
    if lay == tuple(layout):
        return w
    n_in, n_mid, n_out = layout
    keep = np.concatenate([np.arange(0, 1 + n_in), 1 + lay[0] + np.arange(n_mid),
                           1 + lay[0] + lay[1] + np.arange(n_out)]).astype(np.int64)
    return np.ascontiguousarray(w[keep])


def verify_assignment(ctx: Context, g: GenQAP, assignment: QapSet) -> bool:
    """verifyAssignment (src/QAP.hs:276-282) in R1CS form: one upload + one check."""
    w = witness_vector(assignment, g.layout)
    nv, _ = ctx.r1cs_check_host(g, w)
    return nv == 0


def verification_witness_zk(ctx: Context, d1: int, d2: int, d3: int, g: GenQAP, assignment: QapSet):
    """verificationWitnessZk (src/QAP.hs:300-327) on the FFT-built QAP: `Just h` (stripped coefficient
    list) or None."""
    m = ctx.upload_r1cs(g)
    w = ctx.upload_witness(witness_vector(assignment, g.layout))
    try:
        bufs, ok = ctx.qap_witness(m, w, (d1, d2, d3), want=("h",))
    finally:
        w.free()
        m.free()
    return strip(from_limbs(bufs["h"])) if ok else None


def verification_witness(ctx: Context, g: GenQAP, assignment: QapSet):
    """verificationWitness (src/QAP.hs:292-298)."""
    return verification_witness_zk(ctx, 0, 0, 0, g, assignment)


def create_polynomials_fft(ctx: Context, columns: Sequence[Sequence[int]]) -> List[List[int]]:
    """The per-wire work of createPolynomialsFFT (src/QAP.hs:512-525): each column is a wire's values in
    ascending-root order; returns the stripped interpolants."""
    n = max((len(c) for c in columns), default=1)
    N = 1
    while N < n:
        N <<= 1
    arr = np.zeros((len(columns), N, 4), np.uint64)
    for i, col in enumerate(columns):
        if len(col):
            arr[i, :len(col)] = to_limbs(col)
    out = ctx.interpolate_columns(arr)
    return [strip(from_limbs(out[i])) for i in range(len(columns))]


def create_polynomials(ctx: Context, xs: Sequence[int], ys: Sequence[Sequence[int]]):
    """createPolynomials' Lagrange build (src/QAP.hs:486-508): (stripped interpolants, target)."""
    polys, target = ctx.lagrange(xs, ys, True)
    return [strip(p) for p in polys], strip(target)


# ---------------------------------------------------------------------------------------------------
# Bulletproofs backend: linear constraints (src/Circuit/Bulletproofs.hs:116-129, 329-338)
# ---------------------------------------------------------------------------------------------------
SECP256K1_FN = 2   # modulus id of the secp256k1 scalar field (acg_linear_constraints_check only)
_MODULI = {SECP256K1_FN: 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141}


def modulus_of(modulus_id: int) -> int:
    return _MODULI[modulus_id] if modulus_id in _MODULI else field_constants(modulus_id)["modulus"]


def check_linear_constraints(ctx: "Context", modulus_id: int, constraints: Sequence[Dict], assignment: Dict) -> Tuple[int, int]:
    """checkLinearConstraint (src/Circuit/Bulletproofs.hs:329-338) for a batch of LinearConstraint values against one
    Assignment, on the device: constraint i holds <=> wL_i.aL + wR_i.aR + wO_i.aO == wV_i.v + c_i over the field
    `modulus_id` (SECP256K1_FN for the reference's Bulletproofs circuits; BN254_FR / BLS12_381_FR also work).
    constraints: dicts {"wL", "wR", "wO", "wV": {index: weight}, "c": constant} (LinearConstraint, :116-129);
    assignment: {"aL", "aR", "aO", "v": {index: value}} -- a wire it lacks counts as 0 (dotProduct,
    src/Circuit/Affine.hs:121-125).  Returns (number of violated constraints, first violated index or -1)."""
    r = modulus_of(modulus_id)
    sides = ("L", "R", "O")
    dims = {k: 1 + max([-1] + [ix for c in constraints for ix in c["w" + k]]) for k in sides}   # columns the weights use
    off = {"L": 0, "R": dims["L"], "O": dims["L"] + dims["R"]}
    n_lhs = dims["L"] + dims["R"] + dims["O"]
    n_rhs = 1 + max([-1] + [ix for c in constraints for ix in c["wV"]])
    x = [0] * n_lhs
    for k in sides:
        for ix, val in assignment.get("a" + k, {}).items():
            if ix < dims[k]:
                x[off[k] + ix] = val % r
    v = [0] * n_rhs
    for ix, val in assignment.get("v", {}).items():
        if ix < n_rhs:
            v[ix] = val % r

    def csr(rows):
        rowptr, cols, vals = [0], [], []
        for row in rows:
            for col, wgt in sorted(row):
                if wgt % r:
                    cols.append(col)
                    vals.append(wgt % r)
            rowptr.append(len(cols))
        return (np.array(rowptr, np.uint32), np.array(cols, np.uint32),
                to_limbs(vals) if vals else np.zeros((0, 4), np.uint64))
    lhs = csr([[(off[k] + ix, wgt) for k in sides for ix, wgt in c["w" + k].items()] for c in constraints])
    rhs = csr([list(c["wV"].items()) for c in constraints])
    cst = to_limbs([c["c"] % r for c in constraints]) if constraints else np.zeros((0, 4), np.uint64)
    xs = to_limbs(x) if x else np.zeros((0, 4), np.uint64)
    vs = to_limbs(v) if v else np.zeros((0, 4), np.uint64)
    structs = [AcgCsr(m[0].ctypes.data_as(_lib.u32p), m[1].ctypes.data_as(_lib.u32p), m[2].ctypes.data_as(_lib.u64p),
                      len(m[1])) for m in (lhs, rhs)]
    nv, fb = C.c_uint64(), C.c_uint64()
    _check(_lib.lib().acg_linear_constraints_check(ctx._h, modulus_id, len(constraints), n_lhs, n_rhs, C.byref(structs[0]),
                                                   C.byref(structs[1]), _ptr(cst), _ptr(xs), _ptr(vs), C.byref(nv),
                                                   C.byref(fb)), ctx)
    return nv.value, (-1 if fb.value == UINT64_MAX else fb.value)


def check_linear_constraint(ctx: "Context", modulus_id: int, constraint: Dict, assignment: Dict) -> bool:
    """checkLinearConstraint (src/Circuit/Bulletproofs.hs:329-338) for one constraint."""
    return check_linear_constraints(ctx, modulus_id, [constraint], assignment)[0] == 0


# ---------------------------------------------------------------------------------------------------
# per-wire QAP value: QAP f (src/QAP.hs:74-79) as createPolynomials[FFT] returns it
# ---------------------------------------------------------------------------------------------------
class QAP:
    """QAP f: one polynomial per wire for the left / right / output sets plus the target.  Storage is dense:
    `left`, `right`, `out` are (n_cols, N, 4) limb arrays in qapSetToMap wire order (index 0 = the constant), zero-padded
    little-endian coefficients; `target` is a stripped integer coefficient list.  kind: "fft" (roots = powers of the
    2^k-th root of unity, src/QAP.hs:512-525) or "lagrange" (arbitrary roots, :486-508)."""

    def __init__(self, field: int, layout, n_rows: int, left, right, out, target: List[int], kind: str, roots=None,
                 present=None):
        self.field, self.layout, self.n_rows = field, tuple(layout), n_rows
        self.left, self.right, self.out = left, right, out
        self.target, self.kind, self.roots = target, kind, roots
        # per set, which wires the reference's Map has a key for: the wires that occur in some gate's row of that set
        # (createMapGenQap, src/QAP.hs:233-239; addMissingZeroes fills in ROOTS of those wires, not wires).  None: all.
        self.present = present
        self._dev = None   # (Context, acg_qap handle): the value resident on the device, uploaded on first use

    def device_handle(self, ctx: "Context"):
        """The acg_qap handle of this value on `ctx` (acg_qap_upload on first use)."""
        if self._dev is not None and self._dev[0] is ctx and ctx._h:
            return self._dev[1]
        self.free_device()
        tgt = to_limbs(self.target)
        h = C.c_void_p()
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (self.left, self.right, self.out)]
        _check(_lib.lib().acg_qap_upload(ctx._h, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), arrs[0].shape[0],
                                         arrs[0].shape[1], _ptr(tgt), len(self.target), C.byref(h)), ctx)
        self._dev = (ctx, h)
        return h

    def free_device(self):
        if getattr(self, "_dev", None) is not None:
            ctx, h = self._dev
            if ctx._h and _lib is not None:
                _lib.lib().acg_qap_free(h)
            self._dev = None

    def __del__(self):
        if _lib is not None:
            self.free_device()

    @property
    def n_cols(self) -> int:
        return self.left.shape[0]

    def wire_poly(self, which: str, col: int) -> List[int]:
        """The VPoly of one wire (stripped), which in {"left", "right", "out"}; col in qapSetToMap order."""
        return strip(from_limbs(getattr(self, which)[col]))

    def sets(self):
        """-> three (constant, {i: poly}, {i: poly}, {i: poly}) tuples in the reference's QapSet shape -- what
        json_io.qap_to_json takes.  A set holds the wires that occur in it (`present`): the reference's Maps are sparse
        in wires (src/QAP.hs:563-566), a wire without a key counts as the zero polynomial."""
        n_in, n_mid, n_out = self.layout
        res = []
        for k, which in enumerate(("left", "right", "out")):
            arr = getattr(self, which)
            polys = [strip(from_limbs(arr[c])) for c in range(self.n_cols)]
            has = (lambda c: True) if self.present is None else (lambda c, k=k: bool(self.present[k][c]))
            res.append((polys[0], {i: polys[1 + i] for i in range(n_in) if has(1 + i)},
                        {i: polys[1 + n_in + i] for i in range(n_mid) if has(1 + n_in + i)},
                        {i: polys[1 + n_in + n_mid + i] for i in range(n_out) if has(1 + n_in + n_mid + i)}))
        return res


def _present_wires(g: GenQAP):
    """Per set, the columns with an entry in some row (QAP.present)."""
    res = []
    for _, col, _ in g.mats:
        mask = np.zeros(g.n_cols, bool)
        mask[np.asarray(col, dtype=np.int64)] = True
        res.append(mask)
    return tuple(res)


def _dense_columns(g: GenQAP, N: int) -> List[np.ndarray]:
    """The three GenQAP column sets as dense (n_cols, N, 4) arrays: entry [col, row] = coefficient (row = position of
    the root in ascending order, `Map.elems`), zero where the wire does not occur (addMissingZeroes) and in the padding."""
    res = []
    for rowptr, col, val in g.mats:
        arr = np.zeros((g.n_cols, N, 4), np.uint64)
        rows = np.repeat(np.arange(g.n_rows, dtype=np.int64), np.diff(rowptr.astype(np.int64)))
        arr[col.astype(np.int64), rows] = val
        res.append(arr)
    return res


def create_polynomials_fft_qap(ctx: Context, g: GenQAP, target_full_domain: bool = False) -> QAP:
    """createPolynomialsFFT (src/QAP.hs:512-525): every wire's column interpolated on the 2^k-th roots of unity by one
    batched inverse NTT on the device (3 * n_cols transforms of size N).  Memory is 3 * n_cols * N * 32 bytes: this is
    the API-completeness path for moderate n_cols * N; checks at scale use the GenQAP form (linearity collapse)."""
    N = 1
    while N < max(1, g.n_rows):
        N <<= 1
    cols = _dense_columns(g, N)
    polys = [ctx.interpolate_columns(c) if N > 1 else c for c in cols]
    # FFT.fftTargetPoly (src/QAP.hs:524): prod_{i < n} (X - omega^i), or X^N - 1 with the convention switch of
    # DESIGN.md section 3 (identical when n is a power of two)
    target = ctx.fft_target(N if target_full_domain else g.n_rows)
    return QAP(g.field, g.layout, g.n_rows, polys[0], polys[1], polys[2], target, "fft", present=_present_wires(g))


def arith_circuit_to_qap_fft(ctx: Context, circuit: ArithCircuit, roots: Optional[Sequence[Sequence[int]]] = None,
                             root_start: int = 1) -> QAP:
    """arithCircuitToQAPFFT (src/QAP.hs:552-561); the primitive-root function is getRootOfUnity of the field."""
    return create_polynomials_fft_qap(ctx, arith_circuit_to_gen_qap(circuit, roots, root_start))


def create_polynomials_qap(ctx: Context, g: GenQAP) -> QAP:
    """createPolynomials (src/QAP.hs:486-508): Lagrange interpolation through the GenQAP's own roots on the device (K5),
    target = prod (X - root).  n_rows <= 4096."""
    xs = from_limbs(g.roots)
    n = g.n_rows
    cols = _dense_columns(g, n)
    outs, target = [], None
    for c in cols:
        ys = [from_limbs(c[k]) for k in range(g.n_cols)]
        polys, target = ctx.lagrange(xs, ys, True)
        arr = np.zeros((g.n_cols, n, 4), np.uint64)
        for k, p in enumerate(polys):
            arr[k, :len(p)] = to_limbs(p) if len(p) else np.zeros((0, 4), np.uint64)
        outs.append(arr)
    return QAP(g.field, g.layout, n, outs[0], outs[1], outs[2], strip(target), "lagrange", xs, present=_present_wires(g))


def arith_circuit_to_qap(ctx: Context, circuit: ArithCircuit, roots: Optional[Sequence[Sequence[int]]] = None,
                         root_start: int = 1) -> QAP:
    """arithCircuitToQAP (src/QAP.hs:542-549)."""
    return create_polynomials_qap(ctx, arith_circuit_to_gen_qap(circuit, roots, root_start))


def verification_witness_zk_qap(ctx: Context, d1: int, d2: int, d3: int, qap: QAP, assignment: QapSet):
    """verificationWitnessZk (src/QAP.hs:300-327) on a per-wire QAP value, one C call (acg_qap_verify), everything on
    the device: a = d1*T + sum_k w_k L_k (scale-and-sum), likewise b, c; p = a*b - c by an NTT product; (h, rem) =
    p divMod T by long division for a target of any shape (prod (X - root) of the Lagrange build, fftTargetPoly,
    X^N - 1).  Returns `h` (stripped coefficient list, the reference's `Just quotient`) or None when rem != 0.
    A wire the assignment lacks counts as 0 and a wire the QAP lacks as the zero polynomial (combineWithDefaults,
    src/QAP.hs:163-181): assignment keys beyond the QAP's layout are dropped."""
    w = witness_vector(assignment, qap.layout)
    h_dev = qap.device_handle(ctx)
    cap = _lib.lib().acg_qap_quotient_len(h_dev)
    h = np.zeros((max(cap, 1), 4), np.uint64)
    delta = to_limbs([d1, d2, d3])
    n, div = C.c_uint32(), C.c_int()
    _check(_lib.lib().acg_qap_verify(ctx._h, h_dev, _ptr(w), _ptr(delta), _ptr(h), cap, C.byref(n), C.byref(div)), ctx)
    return strip(from_limbs(h[:n.value])) if div.value else None


def verify_assignment_qap(ctx: Context, qap: QAP, assignment: QapSet) -> bool:
    """verifyAssignment (src/QAP.hs:276-282) on a per-wire QAP value."""
    return verification_witness_zk_qap(ctx, 0, 0, 0, qap, assignment) is not None
