"""CPU tests of the product's host side: the C ABI library loads and exports every declared symbol, the
C++ mirror of the reference's host logic (circuit IR, generateAssignment, gateToGenQAP rows, qapSetToMap)
matches the oracle on the golden circuits, and the device entry points fail loudly without a GPU."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from oracle import qap_oracle as O
from helpers import FIELDS, csr_from_json, gates_acg, gates_oracle, genqap_equals_csr, golden, unhex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(acg):
    hdr = open(os.path.join(ROOT, "include", "acg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(acg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = C.CDLL(acg._lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(acg._lib.PROTOTYPES), declared ^ set(acg._lib.PROTOTYPES)
    assert acg._lib.lib().acg_abi_version() == 1


def test_no_device_fails_loudly(acg):
    """No CPU fallback: creating a context without a GPU is an error, not a silent slow path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(acg.AcgError) as e:
        acg.Context(acg.BN254_FR, 0)
    assert e.value.code == -5


def test_field_constants_and_roots(acg):
    for fid, F in FIELDS.items():
        k = acg.field_constants(fid)
        assert (k["modulus"], k["mont_r"], k["mont_r2"], k["ninv64"], k["two_adicity"]) == \
               (F.r, F.mont_R, F.mont_R2, F.mont_ninv64, F.two_adicity)
        for kk in (0, 1, 2, 3, 10, F.two_adicity):
            assert acg.get_root_of_unity(fid, kk) == F.root_of_unity(kk)
        with pytest.raises(acg.AcgError):
            acg.get_root_of_unity(fid, F.two_adicity + 1)


def _kat(acg):
    gates = [acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(1)), acg.IntermediateWire(0)),
             acg.Mul(acg.Var(acg.InputWire(2)), acg.Var(acg.InputWire(3)), acg.IntermediateWire(1)),
             acg.Mul(acg.Add(acg.ConstGate(10), acg.Var(acg.IntermediateWire(0))), acg.Var(acg.IntermediateWire(1)),
                     acg.OutputWire(0))]
    return acg.ArithCircuit(acg.BN254_FR, gates)


def test_kat1_lowering_and_assignment(acg):
    """test/Test/QAP.hs:48-62 through the C++ mirror == golden (oracle) CSR and witness."""
    k = golden("kats.json")["kat1"]
    c = _kat(acg)
    assert (c.num_gates, c.num_roots, c.valid()) == (3, 3, True)
    a = acg.generate_assignment(c, {0: 2, 1: 3, 2: 4, 3: 5})
    assert a.dims() == (4, 2, 1)
    assert [a.lookup(acg.IntermediateWire(0)), a.lookup(acg.IntermediateWire(1)), a.lookup(acg.OutputWire(0))] == [6, 20, 320]
    assert a.lookup(acg.OutputWire(1)) is None
    g = acg.arith_circuit_to_gen_qap(c, [[7], [8], [9]])
    genqap_equals_csr(acg, g, csr_from_json(k["A"]), csr_from_json(k["B"]), csr_from_json(k["C"]))
    assert acg.from_limbs(g.roots) == [7, 8, 9]
    assert acg.from_limbs(a.to_vector(g.layout)) == unhex(k["w"])
    # roots given out of order: rows come back in ascending-root order (Map key order)
    g2 = acg.arith_circuit_to_gen_qap(c, [[9], [7], [8]])
    assert acg.from_limbs(g2.roots) == [7, 8, 9]
    assert g2.mats[2][1].tolist() == [6, 7, 5]  # C columns: gate 1 (root 7), gate 2 (root 8), gate 0 (root 9)
    # wrong number of roots / duplicate roots: the reference panics or silently merges; here an error
    with pytest.raises(acg.AcgError):
        acg.arith_circuit_to_gen_qap(c, [[7], [8]])
    with pytest.raises(acg.AcgError):
        acg.arith_circuit_to_gen_qap(c, [[7], [7], [9]])
    # faulty assignment of unit_arithCircuitToQapNoFalsePositive
    a.update(acg.IntermediateWire(0), 7)
    assert acg.from_limbs(a.to_vector(g.layout)) == unhex(golden("kats.json")["kat2"]["w"])


def test_eq_gate_and_split_unsplit(acg):
    """unit_eqGate, unit_splitUnsplit (test/Test/Circuit/Arithmetic.hs:154-182) on the C++ evaluator."""
    eq = acg.ArithCircuit(0, [acg.Equal(acg.InputWire(0), acg.IntermediateWire(0), acg.OutputWire(0))])
    F = O.BN254
    for v, want in ((0, 0), (1, 1), (2, 1), (3, 1)):
        a = acg.generate_assignment(eq, {0: v})
        assert a.lookup(acg.OutputWire(0)) == want
        assert a.lookup(acg.IntermediateWire(0)) == (pow(v, -1, F.r) if v else 0)
    nbits = 16
    mids = [acg.IntermediateWire(i) for i in range(nbits)]
    c = acg.ArithCircuit(0, [acg.Split(acg.InputWire(0), mids), acg.Mul(acg.ConstGate(1), acg.unsplit(mids), acg.OutputWire(0))])
    assert c.valid() and c.num_roots == 1 + nbits + 1
    rnd = random.Random(2)
    for v in [0, 1, 65535, 0x8000] + [rnd.randrange(1 << 16) for _ in range(60)]:
        assert acg.generate_assignment(c, {0: v}).lookup(acg.OutputWire(0)) == v


def test_invalid_circuits(acg):
    bad1 = acg.ArithCircuit(0, [acg.Mul(acg.Var(acg.IntermediateWire(3)), acg.ConstGate(1), acg.OutputWire(0))])
    assert not bad1.valid()           # reference to an undefined intermediate wire
    bad2 = acg.ArithCircuit(0, [acg.Mul(acg.ConstGate(1), acg.ConstGate(1), acg.InputWire(0))])
    assert not bad2.valid()           # an input wire used as output
    with pytest.raises(acg.AcgError):
        acg.ArithCircuit.from_words(0, np.array([9, 0, 0], np.uint64))       # unknown gate tag
    with pytest.raises(acg.AcgError):
        acg.ArithCircuit.from_words(0, np.array([1, acg.OutputWire(0), 1, 2, 1, 2], np.uint64))  # Add on empty stack
    with pytest.raises(acg.AcgError):   # Equal on a missing wire: "the impossible happened"
        acg.generate_assignment(acg.ArithCircuit(0, [acg.Equal(acg.InputWire(5), acg.IntermediateWire(0), acg.OutputWire(0))]), {0: 1})
    F = O.BN254
    c = _kat(acg)
    with pytest.raises(acg.AcgError) as e:   # non-canonical input is rejected, not reduced
        import ctypes
        ix = np.array([0], np.uint32); vals = acg.to_limbs([F.r])
        h = ctypes.c_void_p()
        rc = acg._lib.lib().acg_generate_assignment(c._h, ix.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p), 1, ctypes.byref(h))
        acg.qap._check(rc)
    assert e.value.code == -2


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_mixed_circuits_match_oracle(acg, idx):
    case = golden("mixed_circuits.json")[idx]
    fid = case["field"]
    og = gates_oracle(case["gates"])
    c = acg.ArithCircuit(fid, gates_acg(acg, og))
    assert c.valid()
    inputs = {int(k): int(v, 16) for k, v in case["inputs"].items()}
    a = acg.generate_assignment(c, inputs)
    lay = tuple(case["layout"])
    assert a.dims() == lay
    g = acg.arith_circuit_to_gen_qap(c, None, 1)
    assert g.layout == lay
    genqap_equals_csr(acg, g, csr_from_json(case["A"]), csr_from_json(case["B"]), csr_from_json(case["C"]))
    assert acg.from_limbs(a.to_vector(lay)) == unhex(case["w"])


def test_synth_matches_oracle_and_golden(acg):
    for case in golden("synth.json"):
        fid, n = case["field"], case["n"]
        g, w = acg.synth_r1cs(fid, n, case["seed"], case["dense"])
        genqap_equals_csr(acg, g, csr_from_json(case["A"]), csr_from_json(case["B"]), csr_from_json(case["C"]))
        wi = acg.from_limbs(w)
        assert wi[1025:] == unhex(case["w_tail"]) and wi[:8] == unhex(case["w_head"])
        # the equivalent ArithCircuit through parse -> generateAssignment -> lowering gives the same system
        circ, inp = acg.synth_circuit(fid, n, case["seed"], case["dense"])
        assert circ.valid() and circ.num_gates == n
        g2 = acg.arith_circuit_to_gen_qap(circ, layout=g.layout)   # not every input wire is referenced at small n
        a2 = acg.generate_assignment(circ, inp)
        assert g2.layout == g.layout == (1024, n - 1, 1) and a2.dims() == g.layout
        for m1, m2 in zip(g.mats, g2.mats):
            assert m1[0].tolist() == m2[0].tolist() and m1[1].tolist() == m2[1].tolist() and (m1[2] == m2[2]).all()
        assert (a2.to_vector(g2.layout) == w).all()


def test_synth_statistics(acg):
    """S(n): ~5.5 nnz per row, coefficient mix 1/2 : 1/4 : 1/4."""
    n = 1 << 12
    g, w = acg.synth_r1cs(0, n, 20260002, False)
    nnz = sum(g.nnz)
    assert 5.3 < nnz / n < 5.7
    F = O.BN254
    vals = np.concatenate([g.mats[0][2], g.mats[1][2]])
    is_one = (vals == np.array([1, 0, 0, 0], np.uint64)).all(axis=1)
    m1 = acg.to_limbs([F.r - 1])[0]
    is_m1 = (vals == m1).all(axis=1)
    cols = np.concatenate([g.mats[0][1], g.mats[1][1]])
    wire = cols != 0
    assert 0.45 < is_one[wire].mean() < 0.55 and 0.2 < is_m1[wire].mean() < 0.3
    assert g.mats[2][1].tolist() == list(range(1025, 1025 + n))


def test_gate_plan_levels(acg):
    """Levelisation behind the device witness generation (host side of K6): the KAT-1 circuit has two levels (two
    independent products, then the gate that reads both), an unsplit chain adds one level per stage, and gate lists
    that are not single-assignment / define-before-use are refused."""
    c = _kat(acg)
    assert c.plan_stats() == (2, 2)
    mids = [acg.IntermediateWire(i) for i in range(8)]
    sp = acg.ArithCircuit(0, [acg.Split(acg.InputWire(0), mids),
                              acg.Mul(acg.ConstGate(1), acg.unsplit(mids), acg.OutputWire(0)),
                              acg.Equal(acg.OutputWire(0), acg.IntermediateWire(8), acg.OutputWire(1))])
    assert sp.plan_stats() == (3, 1)
    twice = acg.ArithCircuit(0, [acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(1)), acg.IntermediateWire(0)),
                                 acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(0)), acg.IntermediateWire(0))])
    with pytest.raises(acg.AcgError) as e:
        twice.plan_stats()
    assert e.value.code == -6
    early = acg.ArithCircuit(0, [acg.Mul(acg.Var(acg.IntermediateWire(1)), acg.Var(acg.InputWire(1)), acg.IntermediateWire(0)),
                                 acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(0)), acg.IntermediateWire(1))])
    with pytest.raises(acg.AcgError) as e:
        early.plan_stats()
    assert e.value.code == -6
    circuit, _inputs = acg.synth_circuit(0, 2000, 5)
    levels, width = circuit.plan_stats()
    assert 1 <= levels <= 2000 and width >= 1


def test_parallel_lowering_is_deterministic(acg, monkeypatch):
    """arithCircuitToGenQAP in C++ lowers chunks of the gate list on worker threads (ACG_HOST_THREADS); the CSR is
    identical for every thread count, including gates that emit several rows (Equal, Split) at chunk borders."""
    rnd = random.Random(3)
    gates = []
    mid = 0
    for i in range(9000):
        k = rnd.random()
        if k < 0.8 or mid < 3:
            gates.append(acg.Mul(acg.Add(acg.ConstGate(rnd.randrange(1 << 200)), acg.Var(acg.InputWire(i % 5))),
                                 acg.ScalarMul(rnd.randrange(1 << 100), acg.Var(acg.InputWire((i * 7) % 5))),
                                 acg.IntermediateWire(mid)))
            mid += 1
        elif k < 0.9:
            gates.append(acg.Equal(acg.IntermediateWire(mid - 1), acg.IntermediateWire(mid), acg.IntermediateWire(mid + 1)))
            mid += 2
        else:
            outs = [acg.IntermediateWire(mid + j) for j in range(6)]
            gates.append(acg.Split(acg.IntermediateWire(mid - 2), outs))
            mid += 6
    c = acg.ArithCircuit(0, gates)
    ref = None
    for th in ("1", "2", "3", "7", "32"):
        monkeypatch.setenv("ACG_HOST_THREADS", th)
        g = acg.arith_circuit_to_gen_qap(c, None, 1)
        assert g.n_rows == c.num_roots
        snap = [tuple(np.array(a, copy=True) for a in m) for m in g.mats]
        if ref is None:
            ref = snap
        else:
            assert all((a == b).all() for ma, mb in zip(ref, snap) for a, b in zip(ma, mb)), th


def test_tile_stream_build_does_not_depend_on_threads(acg):
    """acg_r1cs_upload builds the tile stream of the tiled kernel on worker threads over contiguous chunks of the tile
    list (both passes: windows / far columns, then the blobs).  The host-only acg_tile_stream_digest hashes everything
    the build produces -- blobs, tile records, far columns, value offsets, the rows left to the long-row path: equal
    for 1, 2, 5 and 16 threads, for every tile geometry, for a dense system (the roomier geometry is chosen), for the
    reference generator's gate mix (long rows) and for a row shard.  The same call validates the stream structurally
    (validate_tile_stream in abi.cu: every entry / operand word inside its legal places of the CTA's shared memory,
    section order, warp records, row permutation, hand-over fields, far columns, each row covered exactly once) and
    fails with ACG_ERR_INTERNAL otherwise -- the builder is checked on CPU for every geometry without a device."""
    systems = {"S": acg.synth_r1cs(0, 1 << 15, 7)[0], "dense": acg.synth_r1cs(0, 1 << 13, 8, True)[0],
               "mix": acg.synth_mixed_r1cs(0, 1 << 14, 5)[0], "bls": acg.synth_r1cs(1, 5000, 3)[0]}
    for name, g in systems.items():
        for variant in range(9):
            d = [g.tile_stream_digest(variant, t) for t in (1, 2, 5, 16)]
            assert all(x == d[0] for x in d), (name, variant, d)
            assert d[0][2] > 0 and d[0][3] > 0
        d = [g.tile_stream_digest(0, t, (1000, g.n_rows - 777)) for t in (1, 3, 8)]
        assert all(x == d[0] for x in d), name
    # the dense system is bound to the roomier geometry when the default is asked for
    assert systems["dense"].tile_stream_digest(0, 1) == systems["dense"].tile_stream_digest(8, 1)
    assert systems["S"].tile_stream_digest(0, 1) != systems["S"].tile_stream_digest(8, 1)
    with pytest.raises(acg.AcgError) as e:
        systems["S"].tile_stream_digest(99, 1)
    assert e.value.code == -1
