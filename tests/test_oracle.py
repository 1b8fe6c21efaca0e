"""Pins the oracle itself: against the outcomes the reference's own tests assert and the KATs of SURVEY.md
8(c) (committed under tests/golden/), Python big-int oracle vs C oracle.  CPU only."""
import random

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import qap_oracle as O
from helpers import FIELDS, csr_from_json, csr_np, gates_oracle, golden, unhex

F = O.BN254


def test_kat6_constants():
    k = golden("kats.json")["kat6"]
    for Fx in (O.BN254, O.BLS12_381):
        g = k[Fx.name]
        assert (Fx.r, Fx.mont_R, Fx.mont_R2, Fx.mont_ninv64) == tuple(int(g[x], 16) for x in ("r", "R", "R2", "ninv64"))
        assert CO.field_constants(Fx.field_id) == (Fx.r, Fx.mont_R, Fx.mont_R2, Fx.mont_ninv64)
        for kk, v in g["roots_of_unity"].items():
            assert Fx.root_of_unity(int(kk)) == int(v, 16) == CO.root_of_unity(Fx.field_id, int(kk))
    # SURVEY KAT-6 literal values
    assert O.BN254.mont_R == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert O.BN254.mont_ninv64 == 0xc2e1f593efffffff
    assert O.BN254.root_of_unity(28) == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert O.BLS12_381.mont_ninv64 == 0xfffffffeffffffff
    assert O.BLS12_381.root_of_unity(32) == 10238227357739495823651030575849232062558860180284477541189508159991286009131


def _kat_circuit():
    gates = [O.Mul(O.Var(O.inw(0)), O.Var(O.inw(1)), O.midw(0)),
             O.Mul(O.Var(O.inw(2)), O.Var(O.inw(3)), O.midw(1)),
             O.Mul(O.Add(O.ConstGate(10), O.Var(O.midw(0))), O.Var(O.midw(1)), O.outw(0))]
    return gates, [[7], [8], [9]], {0: 2, 1: 3, 2: 4, 3: 5}


def test_unit_arithCircuitToQapCorrect():
    """test/Test/QAP.hs:68-75 -> True; values = SURVEY KAT-1."""
    gates, roots, inp = _kat_circuit()
    qap = O.arith_circuit_to_qap(F, roots, gates)
    asg = O.generate_assignment(F, gates, inp)
    assert O.verify_assignment(F, qap, asg)
    k = golden("kats.json")["kat1"]
    h, a, b, c, rem = O.verification_witness_zk(F, 0, 0, 0, qap, asg)
    assert (asg.mids, asg.outputs) == ({0: 6, 1: 20}, {0: 320})
    assert qap.target == unhex(k["target"]) == [F.r - 504, 191, F.r - 24, 1]
    assert a == unhex(k["a"]) == [268, F.r - 73, 5]
    assert c == unhex(k["c"]) == [7916, F.r - 2131, 143]
    assert b == unhex(k["b"]) and b[0] == 353
    assert h == unhex(k["h"]) and h[0] == F.r - 172 and rem == []


def test_unit_arithCircuitToQapNoFalsePositive():
    """test/Test/QAP.hs:77-90 -> False; residuals = SURVEY KAT-2."""
    gates, roots, _ = _kat_circuit()
    qap = O.arith_circuit_to_qap(F, roots, gates)
    bad = O.QapSet(1, {0: 2, 1: 3, 2: 4, 3: 5}, {0: 7, 1: 20}, {0: 320})
    assert not O.verify_assignment(F, qap, bad)
    gq = O.arith_circuit_to_gen_qap(F, roots, gates)
    lay = O.layout_of(bad)
    A, B, C, _ = O.gen_qap_to_csr(F, gq, lay)
    assert O.r1cs_residuals(F, A, B, C, O.witness_vector(F, bad, lay))[0] == [F.r - 1, 0, 20]
    assert O.r1cs_check(F, A, B, C, O.witness_vector(F, bad, lay)) == (2, 0)


def test_bench_and_example_circuit():
    """bench/Circuit.hs:17-35, Example.hs:10-38 (KAT-3): valid under all three builds."""
    g3 = [O.Mul(O.Var(O.inw(0)), O.Var(O.inw(1)), O.midw(0)),
          O.Mul(O.Var(O.midw(0)), O.Add(O.Var(O.inw(0)), O.Var(O.inw(2))), O.outw(0))]
    a3 = O.generate_assignment(F, g3, {0: 7, 1: 5, 2: 4})
    assert (a3.mids[0], a3.outputs[0]) == (35, 385)
    for start, mk in ((0, O.arith_circuit_to_qap_fft), (0, O.arith_circuit_to_qap), (1, O.arith_circuit_to_qap_fft)):
        assert O.verify_assignment(F, mk(F, O.fresh_roots(g3, start), g3), a3)


def test_unit_eqGate():
    """test/Test/Circuit/Arithmetic.hs:154-169."""
    eq = [O.Equal(O.inw(0), O.midw(0), O.outw(0))]
    assert [O.generate_assignment(F, eq, {0: v}).outputs[0] for v in (0, 1, 2, 3)] == [0, 1, 1, 1]
    for v in (0, 5):
        asg = O.generate_assignment(F, eq, {0: v})
        qap = O.arith_circuit_to_qap_fft(F, [[1, 2]], eq)
        assert O.verify_assignment(F, qap, asg)


def test_unit_splitUnsplit():
    """test/Test/Circuit/Arithmetic.hs:171-182 (sampled: the reference sweeps all 2^16 inputs)."""
    nbits = 16
    mids = [O.midw(i) for i in range(nbits)]
    gates = [O.Split(O.inw(0), mids), O.Mul(O.ConstGate(1), O.unsplit(mids), O.outw(0))]
    assert O.valid_arith_circuit(gates)
    rnd = random.Random(1)
    for v in [0, 1, 2, 65535, 32768] + [rnd.randrange(1 << 16) for _ in range(40)]:
        assert O.generate_assignment(F, gates, {0: v}).outputs[0] == v
    asg = O.generate_assignment(F, gates, {0: 0xBEEF})
    qap = O.arith_circuit_to_qap_fft(F, O.fresh_roots(gates, 1), gates)
    assert O.verify_assignment(F, qap, asg)


def test_prop_gateToQapCorrect_like():
    """test/Test/QAP.hs:92-103: single Mul / Equal gate, FFT build, roots [1] / [1,2]."""
    rnd = random.Random(3)
    for _ in range(10):
        nv = rnd.randrange(1, 5)
        def aff(d):
            if d <= 0:
                return O.Var(O.inw(rnd.randrange(nv))) if rnd.randrange(2) else O.ConstGate(rnd.randrange(F.r))
            return O.Add(aff(d - 1), aff(d - 1)) if rnd.randrange(2) else O.ScalarMul(rnd.randrange(F.r), aff(d - 1))
        gate = O.Mul(aff(2), aff(2), O.outw(0)) if rnd.randrange(2) else O.Equal(O.inw(rnd.randrange(nv)), O.midw(0), O.outw(0))
        roots = [1] if gate[0] == "mul" else [1, 2]
        qap = O.gate_to_qap(F, roots, gate)
        for _ in range(3):
            inputs = {i: rnd.randrange(F.r) for i in range(nv)}
            assert O.verify_assignment(F, qap, O.generate_assignment_gate(F, gate, inputs))


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_mixed_circuits_golden(idx):
    """Random Mul/Equal/Split circuits (arbArithCircuit-like): oracle lowering == committed golden, QAP-form
    verification (Lagrange and FFT builds) agrees with the R1CS form, C oracle agrees with Python."""
    case = golden("mixed_circuits.json")[idx]
    Fx = FIELDS[case["field"]]
    gates = gates_oracle(case["gates"])
    assert O.valid_arith_circuit(gates)
    inputs = {int(k): int(v, 16) for k, v in case["inputs"].items()}
    asg = O.generate_assignment(Fx, gates, inputs)
    lay = O.Layout(*case["layout"])
    roots = O.fresh_roots(gates, 1)
    gq = O.arith_circuit_to_gen_qap(Fx, roots, gates, densify=False)
    A, B, C, _ = O.gen_qap_to_csr(Fx, gq, lay)
    for M, name in ((A, "A"), (B, "B"), (C, "C")):
        G = csr_from_json(case[name])
        assert (M.rowptr, M.col, M.val) == (G.rowptr, G.col, G.val)
    w = O.witness_vector(Fx, asg, lay)
    assert w == unhex(case["w"])
    n = len(A.rowptr) - 1
    res = CO.r1cs_eval_check(Fx.field_id, n, lay.n_cols, csr_np(A), csr_np(B), csr_np(C), CO.ints_to_limbs(w), True, 2)
    assert res["n_violations"] == 0
    assert CO.limbs_to_ints(res["Aw"]) == unhex(case["Aw"])
    assert CO.limbs_to_ints(res["Cw"]) == unhex(case["Cw"])
    wb = unhex(case["bad_w"])
    res = CO.r1cs_eval_check(Fx.field_id, n, lay.n_cols, csr_np(A), csr_np(B), csr_np(C), CO.ints_to_limbs(wb))
    assert [res["n_violations"], res["first_bad_row"]] == case["bad_check"]
    if n <= 40:  # reference-shaped QAP check on the small ones (O(m n^2))
        full = O.arith_circuit_to_gen_qap(Fx, roots, gates)
        for mk in (O.create_polynomials, O.create_polynomials_fft):
            assert O.verify_assignment(Fx, mk(Fx, full), asg)


def test_synth_golden_and_c_oracle():
    for case in golden("synth.json"):
        Fx = FIELDS[case["field"]]
        n = case["n"]
        A, B, C, w, lay = O.synth_r1cs(Fx, n, case["seed"], case["dense"])
        for M, name in ((A, "A"), (B, "B"), (C, "C")):
            G = csr_from_json(case[name])
            assert (M.rowptr, M.col, M.val) == (G.rowptr, G.col, G.val)
        assert w[1025:] == unhex(case["w_tail"]) and w[:8] == unhex(case["w_head"])
        res = CO.r1cs_eval_check(Fx.field_id, n, lay.n_cols, csr_np(A), csr_np(B), csr_np(C), CO.ints_to_limbs(w), True)
        assert res["n_violations"] == 0 and CO.limbs_to_ints(res["Bw"]) == unhex(case["Bw"])
        # coset quotient of the C oracle == schoolbook division of the Python oracle
        N = O.next_pow2(n)
        pad = lambda v: np.vstack([v, np.zeros((N - n, 4), np.uint64)])
        a, b, c, h, ok = CO.qap_witness(Fx.field_id, pad(res["Aw"]), pad(res["Bw"]), pad(res["Cw"]))
        assert ok and O.p_norm(Fx, CO.limbs_to_ints(h)) == unhex(case["h"])
        assert O.p_norm(Fx, CO.limbs_to_ints(a)) == unhex(case["a"])


def test_ntt_golden_and_c_oracle():
    for case in golden("ntt.json"):
        Fx = FIELDS[case["field"]]
        v = unhex(case["in"])
        om = Fx.root_of_unity(case["log_n"])
        assert O.ntt(Fx, v, om) == unhex(case["fwd"]) and O.intt(Fx, v, om) == unhex(case["inv"])
        assert CO.limbs_to_ints(CO.ntt(Fx.field_id, CO.ints_to_limbs(v), False, 2)) == unhex(case["fwd"])
        assert CO.limbs_to_ints(CO.ntt(Fx.field_id, CO.ints_to_limbs(v), True, 2)) == unhex(case["inv"])
        # P(w^i) = v_i  (FFT.interpolate's contract, src/QAP.hs:521-523)
        coef = unhex(case["inv"])
        for i in (0, 1, len(v) - 1):
            assert O.p_eval(Fx, coef, pow(om, i, Fx.r)) == v[i]


def test_c_oracle_field_ops_random():
    rnd = random.Random(17)
    for Fx in (O.BN254, O.BLS12_381):
        r = Fx.r
        edge = [0, 1, 2, r - 1, r - 2, Fx.mont_R, (1 << 64) - 1, (1 << 128), r >> 1]
        xs = [rnd.randrange(r) for _ in range(500)] + [e for e in edge for _ in edge]
        ys = [rnd.randrange(r) for _ in range(500)] + [e for _ in edge for e in edge]
        a, b = CO.ints_to_limbs(xs), CO.ints_to_limbs(ys)
        assert CO.limbs_to_ints(CO.fr_binop(Fx.field_id, 0, a, b)) == [(x + y) % r for x, y in zip(xs, ys)]
        assert CO.limbs_to_ints(CO.fr_binop(Fx.field_id, 1, a, b)) == [(x - y) % r for x, y in zip(xs, ys)]
        assert CO.limbs_to_ints(CO.fr_binop(Fx.field_id, 2, a, b)) == [(x * y) % r for x, y in zip(xs, ys)]
        assert CO.limbs_to_ints(CO.fr_binop(Fx.field_id, 3, a[:40], b[:40])) == [pow(x, -1, r) if x else 0 for x in xs[:40]]
        with pytest.raises(ValueError):
            CO.fr_binop(Fx.field_id, 0, CO.ints_to_limbs([r]), CO.ints_to_limbs([1]))


def test_fft_target_conventions():
    """n not a power of two: both plausible fftTargetPoly conventions give the same Bool."""
    g3 = [O.Mul(O.Var(O.inw(0)), O.Var(O.inw(1)), O.midw(0)),
          O.Mul(O.Var(O.midw(0)), O.Var(O.inw(1)), O.midw(1)),
          O.Mul(O.Var(O.midw(1)), O.Add(O.Var(O.inw(0)), O.Var(O.inw(2))), O.outw(0))]
    asg = O.generate_assignment(F, g3, {0: 7, 1: 5, 2: 4})
    bad = asg.copy(); bad.mids[1] = 1
    old = O.FFT_TARGET_FULL_DOMAIN
    try:
        for flag in (False, True):
            O.FFT_TARGET_FULL_DOMAIN = flag
            qap = O.arith_circuit_to_qap_fft(F, O.fresh_roots(g3, 1), g3)
            assert len(qap.target) == (5 if flag else 4)
            assert O.verify_assignment(F, qap, asg) and not O.verify_assignment(F, qap, bad)
    finally:
        O.FFT_TARGET_FULL_DOMAIN = old
