"""CPU ORACLE (test infrastructure, NOT product code) -- Python big-int restatement of the
reference's R1CS / QAP hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product path (arithmetic-circuits_b200/) never does; it fails loudly
when its CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to the reference
checkout, sdiehl/arithmetic-circuits @ 18e15de).  Field/polynomial/FFT arithmetic lives in
un-vendored Hackage packages (galois-field-1.0.2, poly-0.4.0.0, semirings-0.5.3,
galois-fft-0.1.0, pairing-1.0.0; stack.yaml:3-14); their published algorithms are restated
here with Python integers, which makes this file the independent arbiter of "bit-exact".

PARITY STATUS.  Boolean-level parity is PINNED: this oracle reproduces the outcome of every
QAP-path HUnit test the reference holds (test/Test/QAP.hs:68-90, unit_eqGate, unit_splitUnsplit)
and the README/bench example.  Value-level parity (per-wire QAP coefficients, qapTarget, h) is
"parity unpinned": the reference's tests assert no field element, no polynomial coefficient and no
NTT output, and no Haskell toolchain exists in this image to run the reference.  Values are
nevertheless mathematically unique wherever this repo uses them (exact field arithmetic, unique
interpolants); the one free convention (galois-fft orientation / target for n not a power of two)
is behind `FFT_TARGET_FULL_DOMAIN` below.

Representation.  Field elements are Python ints in [0, r).  Wires are tuples ("in"|"mid"|"out", ix)
(src/Circuit/Arithmetic.hs:32-36).  Affine circuits are tuples ("add", l, r) | ("scalar", s, c) |
("const", f) | ("var", wire) (src/Circuit/Affine.hs:26-31).  Gates are ("mul", l, r, out) |
("equal", i, m, out) | ("split", i, [outs]) (src/Circuit/Arithmetic.hs:44-59).  Polynomials are
little-endian coefficient lists with no trailing zeros (poly's VPoly normal form; zero = []).
"""
from __future__ import annotations

from dataclasses import dataclass, field as _dc_field
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------------------------
# Fields (R1 / R12 in SURVEY.md section 8a).  Constants cross-checked in tests/test_oracle.py (KAT-6).
# ----------------------------------------------------------------------------------------------

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
BLS12_381_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


@dataclass(frozen=True)
class Field:
    """A prime field Fr.  `generator` is the multiplicative generator whose powers define the
    2-adic roots of unity (pairing-1.0.0 Data.Pairing.BN254.getRootOfUnity; recalled: 5 for BN254,
    7 for BLS12-381), `two_adicity` the largest s with 2^s | r-1."""

    name: str
    r: int
    generator: int
    two_adicity: int
    field_id: int

    def root_of_unity(self, k: int) -> int:
        """getRootOfUnity k: primitive 2^k-th root of unity (call sites Example.hs:26,
        bench/Circuit.hs:33, test/Test/QAP.hs:101)."""
        if not 0 <= k <= self.two_adicity:
            raise ValueError("no 2^%d-th root of unity in %s" % (k, self.name))
        return pow(self.generator, (self.r - 1) >> k, self.r)

    # Montgomery constants for 4x64-bit limbs (KAT-6)
    @property
    def mont_R(self) -> int:
        return (1 << 256) % self.r

    @property
    def mont_R2(self) -> int:
        return pow(1 << 256, 2, self.r)

    @property
    def mont_ninv64(self) -> int:
        return (-pow(self.r, -1, 1 << 64)) % (1 << 64)

    @property
    def mont_ninv32(self) -> int:
        return (-pow(self.r, -1, 1 << 32)) % (1 << 32)


BN254 = Field("bn254_fr", BN254_R, 5, 28, 0)
BLS12_381 = Field("bls12_381_fr", BLS12_381_R, 7, 32, 1)
FIELDS = {0: BN254, 1: BLS12_381, "bn254": BN254, "bls12_381": BLS12_381}

# galois-fft's fftTargetPoly for n not a power of two: False = prod_{i<n}(X - w^i) (default, see
# SURVEY 8c), True = X^N - 1.  Identical when n is a power of two; the validity Bool is identical
# either way.
FFT_TARGET_FULL_DOMAIN = False

Wire = Tuple[str, int]


def inw(i: int) -> Wire:
    return ("in", i)


def midw(i: int) -> Wire:
    return ("mid", i)


def outw(i: int) -> Wire:
    return ("out", i)


# Haskell's derived Ord on Wire: InputWire < IntermediateWire < OutputWire, then by index
_WIRE_ORD = {"in": 0, "mid": 1, "out": 2}


def wire_key(w: Wire):
    return (_WIRE_ORD[w[0]], w[1])


# ----------------------------------------------------------------------------------------------
# Affine circuits (src/Circuit/Affine.hs)
# ----------------------------------------------------------------------------------------------

def Add(l, r):
    return ("add", l, r)


def ScalarMul(s, c):
    return ("scalar", s, c)


def ConstGate(f):
    return ("const", f)


def Var(w):
    return ("var", w)


def eval_affine_circuit(F: Field, lookup: Callable[[Wire], Optional[int]], c) -> int:
    """evalAffineCircuit, src/Circuit/Affine.hs:73-86.  Failed lookups are 0 (:84)."""
    r = F.r
    # iterative post-order to survive deep `unsplit` chains (256 nested Adds)
    stack = [(c, False)]
    vals: List[int] = []
    while stack:
        node, done = stack.pop()
        tag = node[0]
        if tag == "const":
            vals.append(node[1] % r)
        elif tag == "var":
            v = lookup(node[1])
            vals.append(0 if v is None else v % r)
        elif tag == "add":
            if done:
                b = vals.pop()
                a = vals.pop()
                vals.append((a + b) % r)
            else:
                stack.append((node, True))
                stack.append((node[2], False))
                stack.append((node[1], False))
        elif tag == "scalar":
            if done:
                a = vals.pop()
                vals.append((a * node[1]) % r)
            else:
                stack.append((node, True))
                stack.append((node[2], False))
        else:
            raise ValueError("bad affine node %r" % (tag,))
    assert len(vals) == 1
    return vals[0]


def affine_circuit_to_affine_map(F: Field, c) -> Tuple[int, Dict[Wire, int]]:
    """affineCircuitToAffineMap, src/Circuit/Affine.hs:90-105: (constant, sparse vector).
    Duplicate wires merge with + (:98), scalars distribute (:102).  Coefficients that cancel to 0
    stay in the map, as in the reference (Map.unionWith never deletes)."""
    r = F.r
    stack = [(c, False)]
    vals: List[Tuple[int, Dict[Wire, int]]] = []
    while stack:
        node, done = stack.pop()
        tag = node[0]
        if tag == "var":
            vals.append((0, {node[1]: 1}))
        elif tag == "const":
            vals.append((node[1] % r, {}))
        elif tag == "add":
            if done:
                cr, vr = vals.pop()
                cl, vl = vals.pop()
                merged = dict(vl)
                for k, v in vr.items():
                    merged[k] = (merged[k] + v) % r if k in merged else v
                vals.append(((cl + cr) % r, merged))
            else:
                stack.append((node, True))
                stack.append((node[2], False))
                stack.append((node[1], False))
        elif tag == "scalar":
            if done:
                ce, ve = vals.pop()
                s = node[1] % r
                vals.append(((s * ce) % r, {k: (s * v) % r for k, v in ve.items()}))
            else:
                stack.append((node, True))
                stack.append((node[2], False))
        else:
            raise ValueError("bad affine node %r" % (tag,))
    assert len(vals) == 1
    return vals[0]


def dot_product(F: Field, inp: Dict, comp: Dict) -> int:
    """dotProduct, src/Circuit/Affine.hs:121-125: sum_ix comp[ix] * inp.get(ix, 0)."""
    return sum(c * inp.get(ix, 0) for ix, c in comp.items()) % F.r


def eval_affine_map(F: Field, amap: Tuple[int, Dict], inp: Dict) -> int:
    """evalAffineMap, src/Circuit/Affine.hs:111-119."""
    return (amap[0] + dot_product(F, inp, amap[1])) % F.r


# ----------------------------------------------------------------------------------------------
# QapSet (src/QAP.hs:66-71) and wire access (:331-347)
# ----------------------------------------------------------------------------------------------

@dataclass
class QapSet:
    constant: object
    inputs: Dict[int, object] = _dc_field(default_factory=dict)
    mids: Dict[int, object] = _dc_field(default_factory=dict)
    outputs: Dict[int, object] = _dc_field(default_factory=dict)

    def copy(self) -> "QapSet":
        return QapSet(self.constant, dict(self.inputs), dict(self.mids), dict(self.outputs))

    def fmap(self, f) -> "QapSet":
        return QapSet(f(self.constant), {k: f(v) for k, v in self.inputs.items()},
                      {k: f(v) for k, v in self.mids.items()},
                      {k: f(v) for k, v in self.outputs.items()})

    def _part(self, kind: str) -> Dict[int, object]:
        return {"in": self.inputs, "mid": self.mids, "out": self.outputs}[kind]


def lookup_at_wire(w: Wire, qs: QapSet):
    """lookupAtWire, src/QAP.hs:331-337."""
    return qs._part(w[0]).get(w[1])


def update_at_wire(w: Wire, a, qs: QapSet) -> QapSet:
    """updateAtWire, src/QAP.hs:341-347 (in place here; callers own their QapSet)."""
    qs._part(w[0])[w[1]] = a
    return qs


def initial_qap_set(inputs: Dict[int, int]) -> QapSet:
    """initialQapSet, src/QAP.hs:591-595."""
    return QapSet(1, dict(inputs), {}, {})


def qap_set_to_map(qs: QapSet) -> Dict[int, object]:
    """qapSetToMap, src/QAP.hs:605-620: index 0 constant, then inputs, intermediates, outputs,
    each block sized maxKey+1 (0 when empty)."""
    def max_key(m):
        return max(m) + 1 if m else 0
    n_in = max_key(qs.inputs)
    n_mid = max_key(qs.mids)
    out = {0: qs.constant}
    for k, v in qs.inputs.items():
        out[1 + k] = v
    for k, v in qs.mids.items():
        out[1 + n_in + k] = v
    for k, v in qs.outputs.items():
        out[1 + n_in + n_mid + k] = v
    return out


# ----------------------------------------------------------------------------------------------
# Gates / circuits (src/Circuit/Arithmetic.hs)
# ----------------------------------------------------------------------------------------------

def Mul(l, r, out: Wire):
    return ("mul", l, r, out)


def Equal(i: Wire, m: Wire, out: Wire):
    return ("equal", i, m, out)


def Split(i: Wire, outs: Sequence[Wire]):
    return ("split", i, list(outs))


def output_wires(g) -> List[Wire]:
    """outputWires, src/Circuit/Arithmetic.hs:67-71."""
    if g[0] == "mul":
        return [g[3]]
    if g[0] == "equal":
        return [g[3]]
    return list(g[2])


def eval_gate(F: Field, vars_: QapSet, gate) -> QapSet:
    """evalGate, src/Circuit/Arithmetic.hs:106-145, specialised to lookupAtWire/updateAtWire."""
    r = F.r
    look = lambda w: lookup_at_wire(w, vars_)
    if gate[0] == "mul":  # :120-124
        lval = eval_affine_circuit(F, look, gate[1])
        rval = eval_affine_circuit(F, look, gate[2])
        return update_at_wire(gate[3], (lval * rval) % r, vars_)
    if gate[0] == "equal":  # :125-133
        inp = look(gate[1])
        if inp is None:
            raise RuntimeError("evalGate: the impossible happened")
        res = 0 if inp == 0 else 1
        mid = 0 if inp == 0 else pow(inp, -1, r)
        update_at_wire(gate[2], mid, vars_)
        return update_at_wire(gate[3], res, vars_)
    if gate[0] == "split":  # :134-145, bit ix of the canonical residue (fromP)
        inp = look(gate[1])
        if inp is None:
            raise RuntimeError("evalGate: the impossible happened")
        for ix, o in enumerate(gate[2]):
            update_at_wire(o, (inp >> ix) & 1, vars_)
        return vars_
    raise ValueError("bad gate %r" % (gate[0],))


def eval_arith_circuit(F: Field, gates: Sequence, vars_: QapSet) -> QapSet:
    """evalArithCircuit, src/Circuit/Arithmetic.hs:221-235: left fold of evalGate."""
    for g in gates:
        vars_ = eval_gate(F, vars_, g)
    return vars_


def generate_assignment(F: Field, gates: Sequence, inputs: Dict[int, int]) -> QapSet:
    """generateAssignment, src/QAP.hs:597-603."""
    return eval_arith_circuit(F, gates, initial_qap_set({k: v % F.r for k, v in inputs.items()}))


def generate_roots(take_root: Callable[[], int], gates: Sequence) -> List[List[int]]:
    """generateRoots, src/Circuit/Arithmetic.hs:194-216: 1 root per Mul, 2 per Equal,
    1 + #outputs per Split, drawn in order."""
    out = []
    for g in gates:
        if g[0] == "mul":
            out.append([take_root()])
        elif g[0] == "equal":
            out.append([take_root(), take_root()])
        else:
            out.append([take_root() for _ in range(1 + len(g[2]))])
    return out


def fresh_roots(gates: Sequence, start: int = 0) -> List[List[int]]:
    """evalFresh (generateRoots (fromIntegral <$> fresh)) -- src/Fresh.hs:10-20 counts from 0
    (bench/Circuit.hs:31); Example.hs:24 uses (+1)."""
    ctr = [start]

    def take():
        v = ctr[0]
        ctr[0] += 1
        return v
    return generate_roots(take, gates)


def unsplit(wires: Sequence[Wire]):
    """unsplit, src/Circuit/Arithmetic.hs:238-244."""
    acc = ConstGate(0)
    for ix, w in enumerate(wires):
        acc = Add(acc, ScalarMul(1 << ix, Var(w)))
    return acc


def fetch_vars(c) -> List[Wire]:
    """fetchVars, src/Circuit/Arithmetic.hs:187-191."""
    out, stack = [], [c]
    while stack:
        n = stack.pop()
        if n[0] == "var":
            out.append(n[1])
        elif n[0] == "scalar":
            stack.append(n[2])
        elif n[0] == "add":
            stack.append(n[2])
            stack.append(n[1])
    return out


def valid_arith_circuit(gates: Sequence) -> bool:
    """validArithCircuit, src/Circuit/Arithmetic.hs:158-185."""
    defined = set()
    ok = True
    for g in gates:
        outs = output_wires(g)
        if g[0] == "mul":
            used = fetch_vars(g[1]) + fetch_vars(g[2])
        else:
            used = [g[1]]
        ok = ok and all(o[0] != "in" for o in outs)
        for w in used:
            if w[0] == "out":
                ok = False
            elif w[0] == "mid" and w not in defined:
                ok = False
        defined.update(outs)
    return ok


# ----------------------------------------------------------------------------------------------
# Gate -> R1CS rows (src/QAP.hs:366-474).  One GenQAP ((,) k) k per root: three QapSets of
# (root, coeff) pairs.  Rows are kept exactly as the reference builds them, explicit zeros included.
# ----------------------------------------------------------------------------------------------

@dataclass
class GenQapRow:
    root: int
    left: QapSet     # values: coefficient at this root
    right: QapSet
    out: QapSet


def _const_qs(v) -> QapSet:
    """constantQapSet, src/QAP.hs:114-120."""
    return QapSet(v, {}, {}, {})


def gate_to_gen_qap(F: Field, roots: Sequence[int], gate) -> List[GenQapRow]:
    """gateToGenQAP, src/QAP.hs:366-474."""
    r = F.r
    if gate[0] == "mul":  # :371-395
        if len(roots) != 1:
            raise ValueError("gateToGenQAP: wrong number of roots supplied")
        root = roots[0] % r
        lc, lv = affine_circuit_to_affine_map(F, gate[1])
        rc, rv = affine_circuit_to_affine_map(F, gate[2])
        left, right, out = _const_qs(lc), _const_qs(rc), _const_qs(0)
        for w, c in lv.items():
            update_at_wire(w, c, left)
        for w, c in rv.items():
            update_at_wire(w, c, right)
        update_at_wire(gate[3], 1, out)
        return [GenQapRow(root, left, right, out)]
    if gate[0] == "equal":  # :396-442
        if len(roots) != 2:
            raise ValueError("gateToGenQAP: wrong number of roots supplied")
        i, m, o = gate[1], gate[2], gate[3]
        r0, r1 = roots[0] % r, roots[1] % r

        def mk(cst, triples):
            qs = _const_qs(cst)
            for w, c in triples:       # updateAtWires folds left: later entries win (:350-352)
                update_at_wire(w, c % r, qs)
            return qs
        row0 = GenQapRow(r0, mk(0, [(i, 1), (m, 0), (o, 0)]), mk(0, [(i, 0), (m, 1), (o, 0)]),
                         mk(0, [(i, 0), (m, 0), (o, 1)]))
        row1 = GenQapRow(r1, mk(1, [(i, 0), (m, 0), (o, -1)]), mk(0, [(i, 1), (m, 0), (o, 0)]),
                         mk(0, [(i, 0), (m, 0), (o, 0)]))
        return [row0, row1]
    if gate[0] == "split":  # :443-473
        inp, outs = gate[1], gate[2]
        if len(roots) < 1 or len(roots) - 1 != len(outs):
            raise ValueError("gateToGenQAP: wrong number of roots supplied")
        root = roots[0] % r
        left = _const_qs(0)
        update_at_wire(inp, 0, left)
        for ix, o in enumerate(outs):
            update_at_wire(o, pow(2, ix, r), left)
        right = _const_qs(1)
        update_at_wire(inp, 0, right)
        out = _const_qs(0)
        update_at_wire(inp, 1, out)
        rows = [GenQapRow(root, left, right, out)]
        for rr, o in zip(roots[1:], outs):
            l2 = _const_qs(0)
            update_at_wire(o, 1, l2)
            r2 = _const_qs(1)
            update_at_wire(o, (-1) % r, r2)
            o2 = _const_qs(0)
            update_at_wire(o, 0, o2)
            rows.append(GenQapRow(rr % r, l2, r2, o2))
        return rows
    raise ValueError("gateToGenQAP: wrong number of roots supplied")


@dataclass
class GenQAP:
    """GenQAP (Map k) k, src/QAP.hs:94-99: per wire, a Map root -> coeff."""
    left: QapSet      # values: Dict[root, coeff]
    right: QapSet
    out: QapSet
    target: Dict[int, int]


def _sequence_to_maps(rows: Sequence[GenQapRow], pick) -> QapSet:
    """sequenceQapSet + fmap Map.fromList, src/QAP.hs:104-110, 233-239.  Map.fromList keeps the
    LAST value of a duplicated root."""
    cst: Dict[int, int] = {}
    parts = {"in": {}, "mid": {}, "out": {}}
    for row in rows:
        qs: QapSet = pick(row)
        cst[row.root] = qs.constant
        for kind in ("in", "mid", "out"):
            for ix, c in qs._part(kind).items():
                parts[kind].setdefault(ix, {})[row.root] = c
    return QapSet(cst, parts["in"], parts["mid"], parts["out"])


def create_map_gen_qap(rows: Sequence[GenQapRow]) -> GenQAP:
    """createMapGenQap, src/QAP.hs:233-239."""
    return GenQAP(_sequence_to_maps(rows, lambda x: x.left), _sequence_to_maps(rows, lambda x: x.right),
                  _sequence_to_maps(rows, lambda x: x.out), {row.root: 0 for row in rows})


def add_missing_zeroes(all_roots: Iterable[int], g: GenQAP) -> GenQAP:
    """addMissingZeroes, src/QAP.hs:566-576: left-biased union with {root: 0}."""
    roots = list(all_roots)

    def fill(m: Dict[int, int]) -> Dict[int, int]:
        out = dict(m)
        for rt in roots:
            out.setdefault(rt, 0)
        return out
    return GenQAP(g.left.fmap(fill), g.right.fmap(fill), g.out.fmap(fill), fill(g.target))


def arith_circuit_to_gen_qap(F: Field, roots_per_gate: Sequence[Sequence[int]], gates: Sequence,
                             densify: bool = True) -> GenQAP:
    """arithCircuitToGenQAP, src/QAP.hs:530-539.  `densify=False` skips addMissingZeroes (the
    O(m*n) step) for callers that only want the sparse rows."""
    rows: List[GenQapRow] = []
    for rts, g in zip(roots_per_gate, gates):  # zipWith truncates to the shorter list
        rows.extend(gate_to_gen_qap(F, rts, g))
    gq = create_map_gen_qap(rows)
    if densify:
        gq = add_missing_zeroes([x % F.r for rts in roots_per_gate for x in rts], gq)
    return gq


# ----------------------------------------------------------------------------------------------
# Polynomials: poly-0.4 VPoly semantics (dense, little-endian, normalised) -- R11
# ----------------------------------------------------------------------------------------------

def p_norm(F: Field, a: Sequence[int]) -> List[int]:
    a = [x % F.r for x in a]
    while a and a[-1] == 0:
        a.pop()
    return a


def p_add(F, a, b):
    n = max(len(a), len(b))
    return p_norm(F, [(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)])


def p_sub(F, a, b):
    n = max(len(a), len(b))
    return p_norm(F, [(a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0) for i in range(n)])


def p_scale(F, c, a):
    """scale 0 c p / monomial 0 c * p."""
    return p_norm(F, [c * x for x in a])


def p_mul(F, a, b):
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] += x * y
    return p_norm(F, out)


def p_eval(F, a, x):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % F.r
    return acc


def p_deriv(F, a):
    return p_norm(F, [i * a[i] for i in range(1, len(a))])


def p_quot_rem(F, a, b):
    """Euclidean quotRem on VPoly over a field (semirings Data.Euclidean; schoolbook long division)."""
    r = F.r
    b = p_norm(F, b)
    if not b:
        raise ZeroDivisionError("polynomial division by zero")
    rem = [x % r for x in a]
    if len(rem) < len(b):
        return [], p_norm(F, rem)
    q = [0] * (len(rem) - len(b) + 1)
    inv_lead = pow(b[-1], -1, r)
    for k in range(len(rem) - len(b), -1, -1):
        c = (rem[k + len(b) - 1] * inv_lead) % r
        q[k] = c
        if c:
            for j, y in enumerate(b):
                rem[k + j] = (rem[k + j] - c * y) % r
    return p_norm(F, q), p_norm(F, rem[:len(b) - 1])


# ----------------------------------------------------------------------------------------------
# NTT (galois-fft FFT module, recalled: inverse DFT on the zero-padded value list) -- R9
# ----------------------------------------------------------------------------------------------

def next_pow2(n: int) -> int:
    k = 1
    while k < n:
        k <<= 1
    return k


def log2_exact(n: int) -> int:
    assert n > 0 and n & (n - 1) == 0
    return n.bit_length() - 1


def ntt(F: Field, a: Sequence[int], omega: int) -> List[int]:
    """Forward DFT: out[i] = sum_j a[j] * omega^(i*j).  Iterative radix-2, len(a) = 2^k."""
    r = F.r
    n = len(a)
    if n == 1:
        return [a[0] % r]
    k = log2_exact(n)
    out = [0] * n
    for i in range(n):   # bit reversal
        out[int(format(i, "0%db" % k)[::-1], 2)] = a[i] % r
    length = 2
    while length <= n:
        w_len = pow(omega, n // length, r)
        half = length // 2
        for start in range(0, n, length):
            w = 1
            for j in range(start, start + half):
                u, v = out[j], out[j + half] * w % r
                out[j] = (u + v) % r
                out[j + half] = (u - v) % r
                w = w * w_len % r
        length <<= 1
    return out


def intt(F: Field, a: Sequence[int], omega: int) -> List[int]:
    """Inverse DFT: the unique coefficient vector c with sum_j c[j] omega^(ij) = a[i]."""
    r = F.r
    n = len(a)
    ninv = pow(n, -1, r)
    return [x * ninv % r for x in ntt(F, a, pow(omega, -1, r))]


def fft_interpolate(F: Field, values: Sequence[int]) -> List[int]:
    """FFT.interpolate primRoots (Map.elems col), src/QAP.hs:521-523: zero-pad to N = 2^ceil(log2 n),
    P(omega^i) = v_i with omega = getRootOfUnity(log2 N)."""
    n = next_pow2(max(1, len(values)))
    vals = [v % F.r for v in values] + [0] * (n - len(values))
    return p_norm(F, intt(F, vals, F.root_of_unity(log2_exact(n))))


def fft_target_poly(F: Field, n_roots: int) -> List[int]:
    """FFT.fftTargetPoly primRoots n, src/QAP.hs:524 (convention: see FFT_TARGET_FULL_DOMAIN)."""
    N = next_pow2(max(1, n_roots))
    if FFT_TARGET_FULL_DOMAIN or n_roots == N:
        return p_norm(F, [-1] + [0] * (N - 1) + [1])
    omega = F.root_of_unity(log2_exact(N))
    t, x = [1], 1
    for _ in range(n_roots):
        t = p_mul(F, t, [(-x) % F.r, 1])
        x = x * omega % F.r
    return t


# ----------------------------------------------------------------------------------------------
# QAP polynomials (src/QAP.hs:74-79, 486-525)
# ----------------------------------------------------------------------------------------------

@dataclass
class QAP:
    left: QapSet      # values: polynomials
    right: QapSet
    out: QapSet
    target: List[int]


def lagrange_interpolate(F: Field, xys: Sequence[Tuple[int, int]]) -> List[int]:
    """lagrangeInterpolate, src/QAP.hs:495-508: sum_i (y_i / phi_i) * (roots quot (X - x_i)),
    roots = prod (X - x_i), phi_i = roots'(x_i)."""
    r = F.r
    xs = [x % r for x, _ in xys]
    ys = [y % r for _, y in xys]
    roots = [1]
    for xi in xs:
        roots = p_mul(F, roots, [(-xi) % r, 1])
    d = p_deriv(F, roots)
    acc: List[int] = []
    for xi, yi in zip(xs, ys):
        phi = p_eval(F, d, xi)
        f = yi * pow(phi, -1, r) % r
        q, _ = p_quot_rem(F, roots, [(-xi) % r, 1])
        acc = p_add(F, acc, p_scale(F, f, q))
    return acc


def create_polynomials(F: Field, g: GenQAP) -> QAP:
    """createPolynomials, src/QAP.hs:486-508 (Map.toList => ascending root order)."""
    interp = lambda m: lagrange_interpolate(F, sorted(m.items()))
    target = [1]
    for root in sorted(g.target):
        target = p_mul(F, target, [(-root) % F.r, 1])
    return QAP(g.left.fmap(interp), g.right.fmap(interp), g.out.fmap(interp), target)


def create_polynomials_fft(F: Field, g: GenQAP) -> QAP:
    """createPolynomialsFFT, src/QAP.hs:512-525 (Map.elems => values in ascending root order; the
    root values themselves are ignored)."""
    interp = lambda m: fft_interpolate(F, [v for _, v in sorted(m.items())])
    return QAP(g.left.fmap(interp), g.right.fmap(interp), g.out.fmap(interp),
               fft_target_poly(F, len(g.target)))


def arith_circuit_to_qap(F, roots, gates) -> QAP:
    """arithCircuitToQAP, src/QAP.hs:542-549."""
    return create_polynomials(F, arith_circuit_to_gen_qap(F, roots, gates))


def arith_circuit_to_qap_fft(F, roots, gates) -> QAP:
    """arithCircuitToQAPFFT, src/QAP.hs:552-561."""
    return create_polynomials_fft(F, arith_circuit_to_gen_qap(F, roots, gates))


def gate_to_qap(F, roots, gate) -> QAP:
    """gateToQAP, src/QAP.hs:355-362."""
    rows = gate_to_gen_qap(F, roots, gate)
    return create_polynomials_fft(F, add_missing_zeroes([x % F.r for x in roots], create_map_gen_qap(rows)))


def generate_assignment_gate(F, gate, inputs) -> QapSet:
    """generateAssignmentGate, src/QAP.hs:579-589."""
    return eval_gate(F, initial_qap_set({k: v % F.r for k, v in inputs.items()}), gate)


# ----------------------------------------------------------------------------------------------
# Verification (src/QAP.hs:163-181, 226-230, 276-327)
# ----------------------------------------------------------------------------------------------

def combine_with_defaults(f, default_a, default_b, qa: QapSet, qb: QapSet) -> QapSet:
    """combineWithDefaults, src/QAP.hs:163-181."""
    def comb(ma, mb):
        out = {}
        for k in set(ma) | set(mb):
            out[k] = f(ma.get(k, default_a), mb.get(k, default_b))
        return out
    return QapSet(f(qa.constant, qb.constant), comb(qa.inputs, qb.inputs), comb(qa.mids, qb.mids),
                  comb(qa.outputs, qb.outputs))


def fold_qap_set(f, qs: QapSet):
    """foldQapSet = foldr1 over the derived Foldable (constant, inputs, intermediates, outputs)."""
    items = [qs.constant] + [qs.inputs[k] for k in sorted(qs.inputs)] + \
            [qs.mids[k] for k in sorted(qs.mids)] + [qs.outputs[k] for k in sorted(qs.outputs)]
    acc = items[-1]
    for x in reversed(items[:-1]):
        acc = f(x, acc)
    return acc


def verification_witness_zk(F: Field, d1: int, d2: int, d3: int, qap: QAP, assignment: QapSet):
    """verificationWitnessZk, src/QAP.hs:300-327.  Returns (h or None, a, b, c, remainder)."""
    def scaled(x: QapSet) -> QapSet:
        return combine_with_defaults(lambda a, b: p_scale(F, b, a), [], 0, x, assignment)
    summ = lambda qs: fold_qap_set(lambda x, y: p_add(F, x, y), qs)
    left = p_add(F, p_scale(F, d1, qap.target), summ(scaled(qap.left)))
    right = p_add(F, p_scale(F, d2, qap.target), summ(scaled(qap.right)))
    outp = p_add(F, p_scale(F, d3, qap.target), summ(scaled(qap.out)))
    io = p_sub(F, p_mul(F, left, right), outp)
    quotient, remainder = p_quot_rem(F, io, qap.target)
    return (quotient if not remainder else None), left, right, outp, remainder


def verification_witness(F, qap, assignment):
    """verificationWitness, src/QAP.hs:292-298."""
    return verification_witness_zk(F, 0, 0, 0, qap, assignment)[0]


def verify_assignment(F, qap, assignment) -> bool:
    """verifyAssignment, src/QAP.hs:276-282."""
    return verification_witness(F, qap, assignment) is not None


# ----------------------------------------------------------------------------------------------
# R1CS form: GenQAP columns -> CSR rows in ascending-root order, witness vector
# in qapSetToMap order, residuals (A.w o B.w - C.w).
# ----------------------------------------------------------------------------------------------

@dataclass
class Layout:
    """Witness index layout of qapSetToMap (src/QAP.hs:605-620) for fixed block sizes."""
    n_in: int
    n_mid: int
    n_out: int

    @property
    def n_cols(self) -> int:
        return 1 + self.n_in + self.n_mid + self.n_out

    def col(self, w: Wire) -> int:
        if w[0] == "in":
            return 1 + w[1]
        if w[0] == "mid":
            return 1 + self.n_in + w[1]
        return 1 + self.n_in + self.n_mid + w[1]


def layout_of(*qsets: QapSet) -> Layout:
    mk = lambda ms: max((max(m) + 1 for m in ms if m), default=0)
    return Layout(mk([q.inputs for q in qsets]), mk([q.mids for q in qsets]), mk([q.outputs for q in qsets]))


@dataclass
class CSR:
    rowptr: List[int]
    col: List[int]
    val: List[int]

    @property
    def nnz(self) -> int:
        return len(self.col)


def gen_qap_to_csr(F: Field, g: GenQAP, lay: Layout, keep_zeros: bool = False):
    """Transpose the column-major GenQAP into three CSR matrices, one row per root in ascending
    canonical-residue order (Map key order).  Explicit zeros are dropped unless keep_zeros."""
    roots = sorted(g.target)
    ridx = {rt: i for i, rt in enumerate(roots)}

    def one(qs: QapSet) -> CSR:
        rows: List[List[Tuple[int, int]]] = [[] for _ in roots]
        def put(colidx, m):
            for rt, c in m.items():
                c %= F.r
                if c or keep_zeros:
                    rows[ridx[rt]].append((colidx, c))
        put(0, qs.constant)
        for ix, m in qs.inputs.items():
            put(lay.col(("in", ix)), m)
        for ix, m in qs.mids.items():
            put(lay.col(("mid", ix)), m)
        for ix, m in qs.outputs.items():
            put(lay.col(("out", ix)), m)
        rowptr, col, val = [0], [], []
        for rw in rows:
            rw.sort()
            for c_, v_ in rw:
                col.append(c_)
                val.append(v_)
            rowptr.append(len(col))
        return CSR(rowptr, col, val)
    return one(g.left), one(g.right), one(g.out), roots


def witness_vector(F: Field, assignment: QapSet, lay: Layout) -> List[int]:
    """Dense w in qapSetToMap order; wires the assignment lacks are 0 (src/QAP.hs:314 default)."""
    w = [0] * lay.n_cols
    w[0] = assignment.constant % F.r
    for kind in ("in", "mid", "out"):
        for ix, v in assignment._part(kind).items():
            w[lay.col((kind, ix))] = v % F.r
    return w


def csr_matvec(F: Field, m: CSR, w: Sequence[int]) -> List[int]:
    r = F.r
    return [sum(m.val[k] * w[m.col[k]] for k in range(m.rowptr[i], m.rowptr[i + 1])) % r
            for i in range(len(m.rowptr) - 1)]


def r1cs_residuals(F: Field, A: CSR, B: CSR, C: CSR, w: Sequence[int]):
    """(A.w) o (B.w) - C.w per row; valid <=> all zero (equivalent to verifyAssignment)."""
    aw, bw, cw = csr_matvec(F, A, w), csr_matvec(F, B, w), csr_matvec(F, C, w)
    res = [(x * y - z) % F.r for x, y, z in zip(aw, bw, cw)]
    return res, aw, bw, cw


def r1cs_check(F: Field, A: CSR, B: CSR, C: CSR, w: Sequence[int]) -> Tuple[int, int]:
    """(number of violated rows, first violated row or -1)."""
    res = r1cs_residuals(F, A, B, C, w)[0]
    bad = [i for i, x in enumerate(res) if x]
    return len(bad), (bad[0] if bad else -1)


def qap_witness_ntt(F: Field, aw, bw, cw, d1=0, d2=0, d3=0):
    """The linearity-collapsed FFT path: a = iNTT(A.w) + d1*T etc. on the N = 2^k domain
    with T = X^N - 1, h = (a*b - c) / T.  Returns (a, b, c, h, divisible) as length-N / N+1 lists,
    un-normalised (fixed length), matching the device layout."""
    r = F.r
    n = len(aw)
    N = next_pow2(max(1, n))
    om = F.root_of_unity(log2_exact(N))
    pad = lambda v: [x % r for x in v] + [0] * (N - n)
    a, b, c = intt(F, pad(aw), om), intt(F, pad(bw), om), intt(F, pad(cw), om)
    T = [(-1) % r] + [0] * (N - 1) + [1]
    af = p_add(F, a, p_scale(F, d1, T))
    bf = p_add(F, b, p_scale(F, d2, T))
    cf = p_add(F, c, p_scale(F, d3, T))
    q, rem = p_quot_rem(F, p_sub(F, p_mul(F, af, bf), cf), T)
    return af, bf, cf, q, (not rem)


# ----------------------------------------------------------------------------------------------
# Synthetic family S(n, seed, field) -- SURVEY 8d.  Mirrored bit-for-bit by the product's C++
# generator (arithmetic-circuits_b200/csrc/host/synth.cpp) so tests can cross-check both.
# ----------------------------------------------------------------------------------------------

_M64 = (1 << 64) - 1
SYNTH_N_INPUTS = 1024
SYNTH_NEAR_WINDOW = 64


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _M64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        """Uniform-enough index: 64-bit draw mod n (n << 2^64)."""
        return self.next() % n

    def field(self, F: Field) -> int:
        """4 outputs as LE limbs, top limb masked to the field's bit length, rejection-sampled < r."""
        top_bits = F.r.bit_length() - 192
        while True:
            l0, l1, l2, l3 = self.next(), self.next(), self.next(), self.next()
            l3 &= (1 << top_bits) - 1
            v = l0 | (l1 << 64) | (l2 << 128) | (l3 << 192)
            if v < F.r:
                return v


def synth_circuit(F: Field, n: int, seed: int, dense: bool = False):
    """S(n, seed, field): returns (gates, inputs) expressed as a reference ArithCircuit of n Mul
    gates over SYNTH_N_INPUTS input wires.  Witness index of InputWire i is 1+i, of gate g's output
    1+1024+g (the last gate is OutputWire 0, which lands on the same index because
    n_mid = n-1).  Draw order per gate, per side (left then right): const? (1 draw, p=1/4:
    next()&3==0 -> field draw); then 2 wire terms, each: near? (next()&1), index draw, coefficient
    kind (next()&3: 0,1 -> 1; 2 -> r-1; 3 -> uniform field draw; `dense` forces uniform)."""
    rng = SplitMix64(seed)
    inputs = {i: rng.field(F) for i in range(SYNTH_N_INPUTS)}
    gates = []

    def wire_of(idx: int) -> Wire:   # witness index (>=1) -> wire
        return inw(idx - 1) if idx <= SYNTH_N_INPUTS else midw(idx - 1 - SYNTH_N_INPUTS)

    for g in range(n):
        avail = 1 + SYNTH_N_INPUTS + g       # witness indices [1, avail) are defined
        sides = []
        for _side in range(2):
            terms = []
            if rng.next() & 3 == 0:
                terms.append(ConstGate(rng.field(F)))
            for _t in range(2):
                near = rng.next() & 1
                if near:
                    lo = max(1, avail - SYNTH_NEAR_WINDOW)
                    idx = lo + rng.below(avail - lo)
                else:
                    idx = 1 + rng.below(avail - 1)
                kind = 3 if dense else (rng.next() & 3)
                if kind <= 1:
                    terms.append(Var(wire_of(idx)))
                elif kind == 2:
                    terms.append(ScalarMul(F.r - 1, Var(wire_of(idx))))
                else:
                    terms.append(ScalarMul(rng.field(F), Var(wire_of(idx))))
            acc = terms[0]
            for t in terms[1:]:
                acc = Add(acc, t)
            sides.append(acc)
        out = outw(0) if g == n - 1 else midw(g)
        gates.append(Mul(sides[0], sides[1], out))
    return gates, inputs


def synth_r1cs(F: Field, n: int, seed: int, dense: bool = False):
    """S(n, seed, field) lowered through the reference-shaped path of this oracle:
    (A, B, C, w, layout)."""
    gates, inputs = synth_circuit(F, n, seed, dense)
    roots = fresh_roots(gates, 0)
    gq = arith_circuit_to_gen_qap(F, roots, gates, densify=False)
    assignment = generate_assignment(F, gates, inputs)
    lay = Layout(SYNTH_N_INPUTS, max(n - 1, 0), 1)
    A, B, C, _ = gen_qap_to_csr(F, gq, lay)
    return A, B, C, witness_vector(F, assignment, lay), lay


# ----------------------------------------------------------------------------------------------
# Bulletproofs backend: linear constraints (src/Circuit/Bulletproofs.hs:116-129, 329-338)
# ----------------------------------------------------------------------------------------------
SECP256K1_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141   # order of the secp256k1 group


def check_linear_constraint(modulus: int, lc: Dict[str, Dict[int, int]], asg: Dict[str, Dict[int, int]]) -> bool:
    """checkLinearConstraint, src/Circuit/Bulletproofs.hs:329-338:
        wL `dotProduct` aL + wR `dotProduct` aR + wO `dotProduct` aO == wV `dotProduct` v + c
    with dotProduct as src/Circuit/Affine.hs:121-125 (a wire the assignment lacks counts as 0).
    lc: {"wL", "wR", "wO", "wV": {index: weight}, "c": constant};  asg: {"aL", "aR", "aO", "v": {index: value}}."""
    dot = lambda wgt, val: sum(c * val.get(ix, 0) for ix, c in wgt.items())
    lhs = dot(lc["wL"], asg["aL"]) + dot(lc["wR"], asg["aR"]) + dot(lc["wO"], asg["aO"])
    rhs = dot(lc["wV"], asg["v"]) + lc["c"]
    return (lhs - rhs) % modulus == 0


# ----------------------------------------------------------------------------------------------
# limb helpers shared by tests
# ----------------------------------------------------------------------------------------------

def to_limbs(v: int) -> List[int]:
    return [(v >> (64 * i)) & _M64 for i in range(4)]


def from_limbs(l: Sequence[int]) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)
