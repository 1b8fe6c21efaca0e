"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI (libacg.so);
expected values come from the oracle (Python big-int / C), the committed golden vectors, or -- at
BASELINE sizes -- size-independent properties.  Integer work: the bar is bit-exact."""
import os
import random

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import qap_oracle as O
from helpers import FIELDS, csr_from_json, gates_acg, gates_oracle, golden, make_genqap, oracle_check, unhex

pytestmark = pytest.mark.gpu
F_R_MINUS_1 = O.BN254.r - 1


# ------------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("fid", [0, 1])
def test_field_ops_device(acg, ctxs, fid):
    F, ctx = FIELDS[fid], ctxs[fid]
    rnd = random.Random(11 + fid)
    r = F.r
    edge = [0, 1, 2, r - 1, r - 2, F.mont_R, (r - F.mont_R) % r, (1 << 32) - 1, (1 << 64) - 1, 1 << 253, r >> 1,
            (r >> 1) + 1, (1 << 224) - 1, 1 << 32, (1 << 96) + 5]
    xs = [rnd.randrange(r) for _ in range(20000)] + [e for e in edge for _ in edge]
    ys = [rnd.randrange(r) for _ in range(20000)] + [e for _ in edge for e in edge]
    a, b = acg.to_limbs(xs), acg.to_limbs(ys)
    assert acg.from_limbs(ctx.fr_binop(0, a, b)) == [(x + y) % r for x, y in zip(xs, ys)]
    assert acg.from_limbs(ctx.fr_binop(1, a, b)) == [(x - y) % r for x, y in zip(xs, ys)]
    assert acg.from_limbs(ctx.fr_binop(2, a, b)) == [(x * y) % r for x, y in zip(xs, ys)]
    inv_in = xs[:300] + edge
    assert acg.from_limbs(ctx.fr_binop(3, acg.to_limbs(inv_in), acg.to_limbs(inv_in))) == \
        [pow(x, -1, r) if x else 0 for x in inv_in]
    # a million random products against the C oracle
    n = 1 << 20
    rng = np.random.default_rng(5)
    big = rng.integers(0, 1 << 63, size=(2, n, 4), dtype=np.uint64)
    big[:, :, 3] &= np.uint64((1 << (r.bit_length() - 192 - 1)) - 1)   # < 2^(bits-1) < r
    assert (ctx.fr_binop(2, big[0], big[1]) == CO.fr_binop(fid, 2, big[0], big[1])).all()
    with pytest.raises(acg.AcgError) as e:
        ctx.fr_binop(0, acg.to_limbs([r]), acg.to_limbs([1]))
    assert e.value.code == -2


# ------------------------------------------------------------------------------------------------ K2
KERNELS = [("rowwise", 0), ("tiled", 0)]   # the tile geometry comes from the `tile_variant` fixture (upload time)


def _select(acg, ctx, kernel, _unused):
    ctx.set_check_kernel(acg.CHECK_ROWWISE if kernel == "rowwise" else acg.CHECK_TILED)


def _reset(acg, ctx):
    ctx.set_check_kernel(acg.CHECK_AUTO)


def _both_kernels(acg, ctx, fn):
    out = []
    for kernel, stages in KERNELS:
        _select(acg, ctx, kernel, stages)
        out.append(fn())
    _reset(acg, ctx)
    assert all(o == out[0] for o in out), out
    return out[0]


def test_reference_unit_tests_on_gpu(acg, ctx_bn):
    """unit_arithCircuitToQapCorrect / NoFalsePositive (test/Test/QAP.hs:68-90), bench + Example circuit,
    through the reference-named API: verify_assignment -> acg_r1cs_check_host."""
    gates = [acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(1)), acg.IntermediateWire(0)),
             acg.Mul(acg.Var(acg.InputWire(2)), acg.Var(acg.InputWire(3)), acg.IntermediateWire(1)),
             acg.Mul(acg.Add(acg.ConstGate(10), acg.Var(acg.IntermediateWire(0))), acg.Var(acg.IntermediateWire(1)),
                     acg.OutputWire(0))]
    c = acg.ArithCircuit(0, gates)
    g = acg.arith_circuit_to_gen_qap(c, [[7], [8], [9]])
    a = acg.generate_assignment(c, {0: 2, 1: 3, 2: 4, 3: 5})
    assert _both_kernels(acg, ctx_bn, lambda: acg.verify_assignment(ctx_bn, g, a)) is True
    a.update(acg.IntermediateWire(0), 7)
    assert _both_kernels(acg, ctx_bn, lambda: acg.verify_assignment(ctx_bn, g, a)) is False
    assert _both_kernels(acg, ctx_bn, lambda: ctx_bn.r1cs_check_host(g, a.to_vector(g.layout))) == (2, 0)
    # bench/Circuit.hs:17-24 + Example.hs: KAT-3, h against the oracle's FFT-built QAP
    c3 = acg.ArithCircuit(0, [acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(1)), acg.IntermediateWire(0)),
                              acg.Mul(acg.Var(acg.IntermediateWire(0)), acg.Add(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(2))),
                                      acg.OutputWire(0))])
    a3 = acg.generate_assignment(c3, {0: 7, 1: 5, 2: 4})
    assert (a3.lookup(acg.IntermediateWire(0)), a3.lookup(acg.OutputWire(0))) == (35, 385)
    k3 = golden("kats.json")["kat3"]["qaps"]
    for start, name in ((0, "bench_fft"), (1, "example_fft")):
        g3 = acg.arith_circuit_to_gen_qap(c3, None, start)
        assert acg.verify_assignment(ctx_bn, g3, a3)
        assert acg.verification_witness(ctx_bn, g3, a3) == unhex(k3[name]["h"])
    a3.update(acg.OutputWire(0), 386)
    assert acg.verification_witness(ctx_bn, acg.arith_circuit_to_gen_qap(c3), a3) is None


def test_eq_and_split_gates_on_gpu(acg, ctx_bn):
    eq = acg.ArithCircuit(0, [acg.Equal(acg.InputWire(0), acg.IntermediateWire(0), acg.OutputWire(0))])
    g = acg.arith_circuit_to_gen_qap(eq, [[1, 2]])
    for v in (0, 1, 2, 3, O.BN254.r - 1):
        a = acg.generate_assignment(eq, {0: v})
        assert acg.verify_assignment(ctx_bn, g, a)
        a.update(acg.OutputWire(0), 1 - a.lookup(acg.OutputWire(0)))
        assert not acg.verify_assignment(ctx_bn, g, a)
    mids = [acg.IntermediateWire(i) for i in range(256)]
    sp = acg.ArithCircuit(0, [acg.Split(acg.InputWire(0), mids), acg.Mul(acg.ConstGate(1), acg.unsplit(mids), acg.OutputWire(0))])
    gs = acg.arith_circuit_to_gen_qap(sp, None, 1)
    assert gs.n_rows == 258
    rnd = random.Random(4)
    for v in (0, 1, O.BN254.r - 1, rnd.randrange(O.BN254.r)):
        a = acg.generate_assignment(sp, {0: v})
        assert a.lookup(acg.OutputWire(0)) == v
        assert _both_kernels(acg, ctx_bn, lambda: acg.verify_assignment(ctx_bn, gs, a)) is True
        a.update(acg.IntermediateWire(200), 1 - a.lookup(acg.IntermediateWire(200)))   # flip one bit
        nv, first = _both_kernels(acg, ctx_bn, lambda: ctx_bn.r1cs_check_host(gs, a.to_vector(gs.layout)))
        assert nv >= 1 and first == 0   # row 0 (the recomposition) breaks; the bit stays boolean


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_mixed_circuits_gpu(acg, ctxs, idx):
    """Golden Mul/Equal/Split circuits (256-entry Split rows, empty C rows): counts, first bad row and
    A.w/B.w/C.w bit-exact, both kernels; coset quotient with deltas vs the big-int oracle."""
    case = golden("mixed_circuits.json")[idx]
    fid = case["field"]
    ctx = ctxs[fid]
    c = acg.ArithCircuit(fid, gates_acg(acg, gates_oracle(case["gates"])))
    g = acg.arith_circuit_to_gen_qap(c, None, 1)
    w = acg.to_limbs(unhex(case["w"]))
    wb = acg.to_limbs(unhex(case["bad_w"]))
    assert _both_kernels(acg, ctx, lambda: ctx.r1cs_check_host(g, w)) == (0, -1)
    assert list(_both_kernels(acg, ctx, lambda: ctx.r1cs_check_host(g, wb))) == case["bad_check"]
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    assert _both_kernels(acg, ctx, lambda: ctx.r1cs_check(m, dw)) == (0, -1)
    for kernel, stages in KERNELS:
        _select(acg, ctx, kernel, stages)
        aw, bw, cw = ctx.r1cs_eval(m, dw)
        assert acg.from_limbs(aw) == unhex(case["Aw"])
        assert acg.from_limbs(bw) == unhex(case["Bw"])
        assert acg.from_limbs(cw) == unhex(case["Cw"])
    _reset(acg, ctx)
    if "qap_delta_3_5_7" in case:
        q = case["qap_delta_3_5_7"]
        bufs, ok = ctx.qap_witness(m, dw, (3, 5, 7))
        assert ok == q["ok"]
        for k in ("a", "b", "c", "h"):
            assert acg.strip(acg.from_limbs(bufs[k])) == unhex(q[k]), k


# ------------------------------------------------------------------------------------------------ per-wire QAP (N4)
def _oracle_sets_as_columns(F, qap_sets, layout):
    """oracle QapSet of polynomials -> list over qapSetToMap columns"""
    n_in, n_mid, n_out = layout
    cols = [qap_sets.constant]
    cols += [qap_sets.inputs.get(i, []) for i in range(n_in)]
    cols += [qap_sets.mids.get(i, []) for i in range(n_mid)]
    cols += [qap_sets.outputs.get(i, []) for i in range(n_out)]
    return [O.p_norm(F, c) for c in cols]


def test_per_wire_qap_matches_oracle(acg, ctx_bn):
    """arithCircuitToQAPFFT / arithCircuitToQAP as QAP values: every wire polynomial and the target equal the big-int
    oracle's; verificationWitnessZk on the QAP value (device scale-and-sum + NTT product) equals the oracle's h for
    deltas (3, 5, 7); a bad witness gives None.  KAT-1 circuit (test/Test/QAP.hs:48-62) extended to 4 gates so that the
    FFT domain is exactly the row count, and the 3-gate original for the padded / Lagrange cases."""
    F = O.BN254
    og = [O.Mul(O.Var(O.inw(0)), O.Var(O.inw(1)), O.midw(0)), O.Mul(O.Var(O.inw(2)), O.Var(O.inw(3)), O.midw(1)),
          O.Mul(O.Add(O.ConstGate(10), O.Var(O.midw(0))), O.Var(O.midw(1)), O.midw(2)),
          O.Mul(O.ScalarMul(F.r - 3, O.Var(O.midw(2))), O.Add(O.Var(O.inw(0)), O.ConstGate(5)), O.outw(0))]
    inputs = {0: 2, 1: 3, 2: 4, 3: 5}
    for gates_o, lagrange_roots in ((og, None), (og[:3], [[7], [8], [9]])):
        c = acg.ArithCircuit(0, gates_acg(acg, gates_o))
        a = acg.generate_assignment(c, inputs)
        oa = O.generate_assignment(F, gates_o, inputs)
        # FFT build, roots 1.. (Example.hs:24)
        roots = O.fresh_roots(gates_o, 1)
        oq = O.arith_circuit_to_qap_fft(F, roots, gates_o)
        q = acg.arith_circuit_to_qap_fft(ctx_bn, c, None, 1)
        for which, osets in (("left", oq.left), ("right", oq.right), ("out", oq.out)):
            want = _oracle_sets_as_columns(F, osets, q.layout)
            assert [q.wire_poly(which, k) for k in range(q.n_cols)] == want, which
        assert q.target == O.p_norm(F, oq.target)
        # h from the device division (acg_qap_verify), for T = X^4 - 1 and for the padded domain's prod (X - w^i)
        oh = O.verification_witness_zk(F, 3, 5, 7, oq, oa)[0]
        assert oh is not None and acg.verification_witness_zk_qap(ctx_bn, 3, 5, 7, q, a) == O.p_norm(F, oh)
        assert acg.verification_witness_zk_qap(ctx_bn, 0, 0, 0, q, a) == O.p_norm(F, O.verification_witness(F, oq, oa))
        assert acg.verify_assignment_qap(ctx_bn, q, a)
        # JSON of the QAP value round-trips (N2)
        from arithmetic_circuits_b200 import json_io as J
        l2, r2, o2, t2 = J.qap_from_json(J.loads(J.dumps(J.qap_to_json(*q.sets(), q.target))))
        assert t2 == q.target and l2[1].get(0, []) == q.wire_poly("left", 1)
        if lagrange_roots is not None:
            oql = O.arith_circuit_to_qap(F, lagrange_roots, gates_o)
            ql = acg.arith_circuit_to_qap(ctx_bn, c, lagrange_roots)
            for which, osets in (("left", oql.left), ("right", oql.right), ("out", oql.out)):
                assert [ql.wire_poly(which, k) for k in range(ql.n_cols)] == _oracle_sets_as_columns(F, osets, ql.layout)
            assert ql.target == O.p_norm(F, oql.target)        # KAT-1: X^3 - 24X^2 + 191X - 504
            assert ql.target == [F.r - 504, 191, F.r - 24, 1]
            assert acg.verify_assignment_qap(ctx_bn, ql, a)
            # KAT-1: h of the Lagrange QAP, divided by prod (X - root) on the device
            assert acg.verification_witness_zk_qap(ctx_bn, 0, 0, 0, ql, a) == unhex(golden("kats.json")["kat1"]["h"])
            ohl = O.verification_witness_zk(F, 3, 5, 7, oql, oa)[0]
            assert ohl is not None and acg.verification_witness_zk_qap(ctx_bn, 3, 5, 7, ql, a) == O.p_norm(F, ohl)
            badl = acg.generate_assignment(c, inputs)
            badl.update(acg.IntermediateWire(0), 7)   # unit_arithCircuitToQapNoFalsePositive, test/Test/QAP.hs:77-90
            assert acg.verification_witness_zk_qap(ctx_bn, 0, 0, 0, ql, badl) is None
            assert not acg.verify_assignment_qap(ctx_bn, ql, badl)
            # an assignment with a wire the QAP does not know verifies as in the reference (combineWithDefaults)
            extra = acg.generate_assignment(c, {**inputs, 9: 123})
            assert acg.verify_assignment_qap(ctx_bn, ql, extra)
        # a bad witness
        bad = acg.generate_assignment(c, inputs)
        bad.update(acg.IntermediateWire(0), 7)
        assert acg.verification_witness_zk_qap(ctx_bn, 3, 5, 7, q, bad) is None
        assert not acg.verify_assignment_qap(ctx_bn, q, bad)


@pytest.mark.parametrize("fid,n_rows,seed", [(0, 1, 1), (0, 2, 2), (0, 33, 3), (0, 64, 4), (1, 100, 5), (0, 700, 6)])
def test_qap_verify_division_vs_oracle(acg, ctxs, fid, n_rows, seed):
    """acg_qap_verify on a Lagrange-built QAP of the synthetic family (arbitrary roots, target prod (X - root)) and on
    its FFT build (padded domain for n not a power of two): h equals the big-int oracle's schoolbook quotRem
    (src/QAP.hs:325-327) with and without delta terms; a tampered witness gives None; the long division crosses
    several 32-coefficient blocks at the larger sizes."""
    F, ctx = FIELDS[fid], ctxs[fid]
    rnd = random.Random(seed)
    circuit, inputs = acg.synth_circuit(fid, n_rows, 1000 + seed)
    a = acg.generate_assignment(circuit, inputs)
    roots = rnd.sample(range(1, 1 << 60), n_rows)
    g = acg.arith_circuit_to_gen_qap(circuit, [[r] for r in roots])
    w = acg.from_limbs(acg.witness_vector(a, g.layout))
    mats = [O.CSR(list(map(int, rp)), list(map(int, col)), acg.from_limbs(val)) for rp, col, val in g.mats]
    abc = [O.csr_matvec(F, M, w) for M in mats]          # rows in ascending-root order
    xs = sorted(r % F.r for r in roots)
    for kind in ("lagrange", "fft"):
        if kind == "lagrange":
            q = acg.create_polynomials_qap(ctx, g)
            target = [1]
            for x in xs:
                target = O.p_mul(F, target, [F.r - x, 1])
            polys = [O.lagrange_interpolate(F, list(zip(xs, v))) for v in abc]
        else:
            q = acg.create_polynomials_fft_qap(ctx, g)
            N = 1
            while N < n_rows:
                N <<= 1
            om = F.root_of_unity(N.bit_length() - 1)
            target = [1]
            for i in range(n_rows):
                target = O.p_mul(F, target, [F.r - pow(om, i, F.r), 1])
            polys = [O.fft_interpolate(F, v + [0] * (N - n_rows)) for v in abc]
        assert q.target == O.p_norm(F, target)
        for d in ((0, 0, 0), (rnd.randrange(F.r), rnd.randrange(F.r), rnd.randrange(F.r))):
            pa, pb, pc = (O.p_add(F, O.p_scale(F, dk, target), pk) for dk, pk in zip(d, polys))
            quo, rem = O.p_quot_rem(F, O.p_sub(F, O.p_mul(F, pa, pb), pc), target)
            assert not rem
            assert acg.verification_witness_zk_qap(ctx, d[0], d[1], d[2], q, a) == O.p_norm(F, quo), (kind, d)
        bad = acg.generate_assignment(circuit, inputs)
        victim = acg.IntermediateWire((n_rows - 1) // 2) if n_rows > 1 else acg.OutputWire(0)
        bad.update(victim, (bad.lookup(victim) + 1) % F.r)
        assert acg.verification_witness_zk_qap(ctx, 0, 0, 0, q, bad) is None
        q.free_device()


# ------------------------------------------------------------------------------------------------ K6
@pytest.mark.parametrize("idx", [0, 1, 2])
def test_device_witness_generation_golden(acg, ctxs, idx):
    """generateAssignment on the device (K6) == the golden witness of the Mul/Equal/Split circuits (which the big-int
    oracle produced) == the sequential C++ fold; and the device witness passes the check without a host round trip."""
    case = golden("mixed_circuits.json")[idx]
    fid = case["field"]
    ctx = ctxs[fid]
    c = acg.ArithCircuit(fid, gates_acg(acg, gates_oracle(case["gates"])))
    g = acg.arith_circuit_to_gen_qap(c, None, 1)
    inputs = {int(k): int(v, 16) for k, v in case["inputs"].items()}
    w_gold = acg.to_limbs(unhex(case["w"]))
    dw, n_levels = ctx.generate_assignment(c, inputs, g.layout)
    assert n_levels >= 1 and len(dw) == g.n_cols
    assert (dw.download() == w_gold).all()
    host = acg.generate_assignment(c, inputs).to_vector(g.layout)
    assert (host == w_gold).all()
    m = ctx.upload_r1cs(g)
    assert _both_kernels(acg, ctx, lambda: ctx.r1cs_check(m, dw)) == (0, -1)


@pytest.mark.parametrize("fid,n,seed,dense", [(0, 300, 11, False), (1, 2000, 12, False), (0, 1500, 13, True)])
def test_device_witness_generation_synth(acg, ctxs, fid, n, seed, dense):
    """S(n, seed) as an ArithCircuit: the levelised device evaluation reproduces the sequential witness exactly."""
    ctx = ctxs[fid]
    circuit, inputs = acg.synth_circuit(fid, n, seed, dense)
    g, w = acg.synth_r1cs(fid, n, seed, dense)
    dw, n_levels = ctx.generate_assignment(circuit, inputs, g.layout)
    assert 1 <= n_levels <= n
    assert (dw.download() == w).all()
    m = ctx.upload_r1cs(g)
    assert ctx.r1cs_check(m, dw) == (0, -1)


def test_device_witness_generation_wide_and_special(acg, ctx_bn):
    """A wide, shallow circuit (grid-barrier mode: one level of 5000 Mul gates, then Equal and Split levels), with
    Equal on zero and non-zero inputs and a 254-bit Split; against the sequential host fold."""
    r = O.BN254.r
    rnd = random.Random(99)
    n = 5000
    gates = [acg.Mul(acg.Add(acg.ConstGate(rnd.randrange(r)), acg.Var(acg.InputWire(i % 7))),
                     acg.ScalarMul(rnd.randrange(r), acg.Var(acg.InputWire((i * 3) % 7))), acg.IntermediateWire(i))
             for i in range(n)]
    # in_5 = 0 makes some products zero: Equal sees both cases
    base = n
    for i in range(64):
        gates.append(acg.Equal(acg.IntermediateWire(i), acg.IntermediateWire(base + 2 * i), acg.IntermediateWire(base + 2 * i + 1)))
    base2 = base + 128
    gates.append(acg.Split(acg.IntermediateWire(3), [acg.IntermediateWire(base2 + i) for i in range(254)]))
    gates.append(acg.Mul(acg.unsplit([acg.IntermediateWire(base2 + i) for i in range(254)]), acg.ConstGate(1), acg.OutputWire(0)))
    c = acg.ArithCircuit(0, gates)
    inputs = {i: rnd.randrange(r) for i in range(7)}
    inputs[5] = 0
    host = acg.generate_assignment(c, inputs)
    dw, n_levels = ctx_bn.generate_assignment(c, inputs)
    assert n_levels == 3
    got = dw.download()
    assert (got == host.to_vector()).all()
    assert host.lookup(acg.OutputWire(0)) == host.lookup(acg.IntermediateWire(3))
    # not single-assignment: refused loudly, the sequential fold stays the faithful path
    bad = acg.ArithCircuit(0, [acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(1)), acg.IntermediateWire(0)),
                               acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(0)), acg.IntermediateWire(0))])
    with pytest.raises(acg.AcgError) as e:
        ctx_bn.generate_assignment(bad, {0: 1, 1: 2})
    assert e.value.code == -6
    # an Equal / Split gate on an input wire the assignment lacks: the reference's lookup panics (src/QAP.hs:445,474), the
    # host fold returns BAD_ARG, and so does the device path (a missing wire counts as 0 only inside a Mul gate's terms)
    eq = acg.ArithCircuit(0, [acg.Equal(acg.InputWire(1), acg.IntermediateWire(0), acg.OutputWire(0))])
    sp = acg.ArithCircuit(0, [acg.Split(acg.InputWire(2), [acg.IntermediateWire(i) for i in range(8)])])
    for circ, need in ((eq, 1), (sp, 2)):
        for gen in (lambda cc, ii: acg.generate_assignment(cc, ii), lambda cc, ii: ctx_bn.generate_assignment(cc, ii)):
            with pytest.raises(acg.AcgError) as e:
                gen(circ, {0: 5})
            assert e.value.code == -1
        dwx, _ = ctx_bn.generate_assignment(circ, {0: 5, need: 3})
        assert (dwx.download() == acg.generate_assignment(circ, {0: 5, need: 3}).to_vector()).all()


@pytest.mark.parametrize("fid,n,seed,dense", [(0, 96, 20260002, False), (0, 64, 7, True), (1, 96, 20260005, False),
                                              (0, 1 << 12, 1, False), (0, 5000, 2, True), (1, (1 << 13) + 3, 3, False)])
def test_synth_parity_small(acg, ctxs, fid, n, seed, dense):
    ctx = ctxs[fid]
    g, w = acg.synth_r1cs(fid, n, seed, dense)
    ref = oracle_check(fid, g, w, True)
    assert ref["n_violations"] == 0
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    distinct = len(np.unique(np.concatenate([x[1] for x in g.mats])))
    assert distinct <= g.n_cols
    assert m.algorithmic_bytes == sum(36 * k for k in g.nnz) + 3 * 4 * (n + 1) + 32 * distinct + 8
    for kernel, stages in KERNELS:
        _select(acg, ctx, kernel, stages)
        assert ctx.r1cs_check(m, dw) == (0, -1)
        aw, bw, cw = ctx.r1cs_eval(m, dw)
        assert (aw == ref["Aw"]).all() and (bw == ref["Bw"]).all() and (cw == ref["Cw"]).all()
    assert ctx.r1cs_check_host(g, w) == (0, -1)     # one-shot path (no preprocessing, untagged row-wise)
    # negative variants: SURVEY 8d (+1 on w[1025 + n/3]) and a few random single-limb flips
    rnd = random.Random(seed)
    for t in [1025 + n // 3] + [rnd.randrange(1, g.n_cols) for _ in range(3)]:
        wb = w.copy()
        wb[t, rnd.randrange(3)] ^= np.uint64(1 << rnd.randrange(60))
        refb = oracle_check(fid, g, wb)
        dw.update(wb)
        for kernel, stages in KERNELS:
            _select(acg, ctx, kernel, stages)
            assert ctx.r1cs_check(m, dw) == (refb["n_violations"], refb["first_bad_row"])
        assert ctx.r1cs_check_host(g, wb) == (refb["n_violations"], refb["first_bad_row"])
    _reset(acg, ctx)
    if n <= 96:  # golden h through the reference-named call
        case = [c for c in golden("synth.json") if (c["field"], c["n"], c["seed"]) == (fid, n, seed)][0]
        dw.update(w)
        bufs, ok = ctx.qap_witness(m, dw)
        assert ok and acg.strip(acg.from_limbs(bufs["h"])) == unhex(case["h"])
        assert acg.strip(acg.from_limbs(bufs["a"])) == unhex(case["a"])


def test_edge_shapes(acg, ctx_bn):
    """Empty rows, a single row, rows longer than a tile (row-wise fallback inside the tiled path),
    tile-boundary sizes, and argument validation."""
    F = O.BN254
    rnd = random.Random(9)
    n_cols = 3000
    w = [1] + [rnd.randrange(F.r) for _ in range(n_cols - 1)]
    specs = {
        "single_row": [3],
        "empty_rows": [0, 0, 5, 0, 0, 0, 2, 0],
        "all_empty": [0] * 9,
        "long_row_2000": [4, 2000, 4, 4, 0, 1500, 3],
        "boundary_255_256_257": [5] * 255 + [6] + [7],
        "tile_pool_edge": [448] * 3 + [447, 1, 449] + [2] * 300,
    }
    for name, lens in specs.items():
        n = len(lens)
        mats = []
        for k in range(3):
            rowptr, col, val = [0], [], []
            for ln in lens:
                ln_k = ln if k < 2 else min(ln, 1)
                for _ in range(ln_k):
                    col.append(rnd.randrange(n_cols))
                    kind = rnd.randrange(4)
                    val.append(1 if kind < 2 else (F.r - 1 if kind == 2 else rnd.randrange(F.r)))
                rowptr.append(len(col))
            mats.append(O.CSR(rowptr, col, val))
        g = make_genqap(acg, 0, n, n_cols, (n_cols - 1, 0, 0), *mats)
        wl = acg.to_limbs(w)
        ref = oracle_check(0, g, wl, True)
        m, dw = ctx_bn.upload_r1cs(g), ctx_bn.upload_witness(wl)
        for kernel, stages in KERNELS:
            _select(acg, ctx_bn, kernel, stages)
            assert ctx_bn.r1cs_check(m, dw) == (ref["n_violations"], ref["first_bad_row"]), name
            aw, bw, cw = ctx_bn.r1cs_eval(m, dw)
            assert (aw == ref["Aw"]).all() and (bw == ref["Bw"]).all() and (cw == ref["Cw"]).all(), name
        _reset(acg, ctx_bn)
        assert ctx_bn.r1cs_check_host(g, wl) == (ref["n_violations"], ref["first_bad_row"]), name
    # validation: column out of range, non-monotone rowptr, non-canonical coefficient / witness
    good = make_genqap(acg, 0, 1, 4, (3, 0, 0), O.CSR([0, 1], [1], [5]), O.CSR([0, 1], [2], [1]), O.CSR([0, 1], [3], [1]))
    assert ctx_bn.r1cs_check_host(good, acg.to_limbs([1, 2, 3, 30])) == (0, -1)
    assert ctx_bn.r1cs_check_host(good, acg.to_limbs([1, 2, 3, 31])) == (1, 0)
    for bad, code in ((make_genqap(acg, 0, 1, 4, (3, 0, 0), O.CSR([0, 1], [4], [5]), O.CSR([0, 1], [2], [1]), O.CSR([0, 1], [3], [1])), -1),
                      (make_genqap(acg, 0, 1, 4, (3, 0, 0), O.CSR([0, 1], [1], [F.r]), O.CSR([0, 1], [2], [1]), O.CSR([0, 1], [3], [1])), -2)):
        with pytest.raises(acg.AcgError) as e:
            ctx_bn.r1cs_check_host(bad, acg.to_limbs([1, 2, 3, 30]))
        assert e.value.code == code
    with pytest.raises(acg.AcgError) as e:
        ctx_bn.r1cs_check_host(good, acg.to_limbs([1, 2, 3, F.r]))
    assert e.value.code == -2


def test_row_shards_cover_the_system(acg, ctx_bn):
    """rows partition across devices; here the shards run one after another on one GPU and the
    per-shard results combine (sum / min) to the full-system answer."""
    from arithmetic_circuits_b200 import sharding
    n = 10000
    g, w = acg.synth_r1cs(0, n, 77)
    w[1025 + 1234, 0] ^= np.uint64(2)
    w[1025 + 8000, 1] ^= np.uint64(4)
    ref = oracle_check(0, g, w)
    dw = ctx_bn.upload_witness(w)
    for world in (2, 4, 8):
        tot, first = 0, None
        for r in range(world):
            rb, re = sharding.row_shard(n, world, r)
            m = ctx_bn.upload_r1cs(g, rb, re)
            nv, fb = ctx_bn.r1cs_check(m, dw)
            tot += nv
            if fb >= 0:
                assert rb <= fb < re
                first = fb if first is None else min(first, fb)
            m.free()
        assert (tot, first) == (ref["n_violations"], ref["first_bad_row"])


def test_full_size_properties(acg, ctx_bn):
    """BASELINE configs[1]: S(2^20, 20260002, BN254).  Honest witness -> 0 violations; tampered -> exactly
    the C oracle's count and first row; both kernels agree; A.w o B.w == C.w recomputed from emitted vectors."""
    n = 1 << 20
    g, w = acg.synth_r1cs(0, n, 20260002)
    m, dw = ctx_bn.upload_r1cs(g), ctx_bn.upload_witness(w)
    for kernel, stages in KERNELS:
        _select(acg, ctx_bn, kernel, stages)
        assert ctx_bn.r1cs_check(m, dw) == (0, -1)
    wb = w.copy()
    wb[1025 + n // 3, 0] += np.uint64(1)       # the SURVEY 8d negative variant
    for t in (1, 1024, 1025 + n - 1):
        wb[t, 2] ^= np.uint64(1 << 17)
    ref = oracle_check(0, g, wb, True, n_threads=8)
    assert ref["n_violations"] > 0
    dw.update(wb)
    for kernel, stages in KERNELS:
        _select(acg, ctx_bn, kernel, stages)
        assert ctx_bn.r1cs_check(m, dw) == (ref["n_violations"], ref["first_bad_row"])
    _reset(acg, ctx_bn)
    assert ctx_bn.r1cs_check_host(g, wb) == (ref["n_violations"], ref["first_bad_row"])
    aw, bw, cw = ctx_bn.r1cs_eval(m, dw)
    assert (aw == ref["Aw"]).all() and (bw == ref["Bw"]).all() and (cw == ref["Cw"]).all()


@pytest.mark.parametrize("fid", [0, 1])
def test_constant_wire_not_one(acg, ctxs, fid):
    """The tiled kernel folds general coefficients on column 0 (the reference's constant wire, w[0] = 1) into
    ready-made terms.  A caller may still pass any w[0]: the result must match the oracle for w[0] = 0, 2, r - 1
    and for a random element (counts, first bad row and the emitted A.w / B.w / C.w vectors)."""
    ctx = ctxs[fid]
    n = 3000
    g, w = acg.synth_r1cs(fid, n, 777 + fid)
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    rng = np.random.default_rng(5)
    rnd = rng.integers(0, 1 << 63, size=4, dtype=np.uint64)
    rnd[3] &= np.uint64((1 << 60) - 1)
    r_minus_1 = acg.to_limbs([acg.field_constants(fid)["modulus"] - 1])[0]
    for w0 in (np.array([0, 0, 0, 0], dtype=np.uint64), np.array([2, 0, 0, 0], dtype=np.uint64), r_minus_1, rnd):
        wb = w.copy()
        wb[0] = w0
        ref = oracle_check(fid, g, wb, True)
        assert ref["n_violations"] > 0
        dw.update(wb)
        for kernel, stages in KERNELS:
            _select(acg, ctx, kernel, stages)
            assert ctx.r1cs_check(m, dw) == (ref["n_violations"], ref["first_bad_row"])
            aw, bw, cw = ctx.r1cs_eval(m, dw)
            assert (aw == ref["Aw"]).all() and (bw == ref["Bw"]).all() and (cw == ref["Cw"]).all()
    dw.update(w)
    _select(acg, ctx, "tiled", 0)
    assert ctx.r1cs_check(m, dw) == (0, -1)
    _reset(acg, ctx)


# ------------------------------------------------------------------------------------------------ K3 / K4
def test_ntt_golden(acg, ctxs):
    for case in golden("ntt.json"):
        ctx = ctxs[case["field"]]
        v = acg.to_limbs(unhex(case["in"]))
        assert acg.from_limbs(ctx.ntt(v, False)) == unhex(case["fwd"])
        assert acg.from_limbs(ctx.ntt(v, True)) == unhex(case["inv"])


@pytest.mark.parametrize("fid,log_n", [(0, 3), (0, 9), (0, 11), (0, 12), (0, 13), (0, 16), (0, 17), (0, 20), (1, 10), (1, 14), (1, 18)])
def test_ntt_vs_c_oracle(acg, ctxs, fid, log_n):
    ctx = ctxs[fid]
    rng = np.random.default_rng(log_n)
    v = rng.integers(0, 1 << 63, size=(1 << log_n, 4), dtype=np.uint64)
    v[:, 3] &= np.uint64((1 << 60) - 1)
    fwd = ctx.ntt(v, False)
    assert (fwd == CO.ntt(fid, v, False, 8)).all()
    assert (ctx.ntt(fwd, True) == v).all()          # round trip
    assert (ctx.ntt(v, True) == CO.ntt(fid, v, True, 8)).all()


def test_ntt_2_22_properties(acg, ctx_bn):
    """BASELINE configs[2] size: round trip, linearity, and a spot check of P(w^i) = v_i by Horner on a
    sparse input (delta at position p -> coefficients are w^(-p*j)/N)."""
    log_n = 22
    n = 1 << log_n
    rng = np.random.default_rng(22)
    v = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    v[:, 3] &= np.uint64((1 << 60) - 1)
    c = ctx_bn.ntt(v, True)
    assert (ctx_bn.ntt(c, False) == v).all()
    F = O.BN254
    d = np.zeros((n, 4), np.uint64)
    p = 123457
    d[p, 0] = 1
    coef = acg.from_limbs(ctx_bn.ntt(d, True)[:5])
    om_inv = pow(F.root_of_unity(log_n), -1, F.r)
    ninv = pow(n, -1, F.r)
    assert coef == [pow(om_inv, p * j, F.r) * ninv % F.r for j in range(5)]
    with pytest.raises(acg.AcgError) as e:   # beyond the 2-adicity of BN254 Fr
        acg.qap._check(acg._lib.lib().acg_ntt(ctx_bn._h, d.ctypes.data_as(__import__("ctypes").c_void_p), 29, 0), ctx_bn)
    assert e.value.code == -6


def test_interpolate_columns(acg, ctx_bn):
    """createPolynomialsFFT's per-wire interpolation (src/QAP.hs:512-525): odd batch counts, padding."""
    F = O.BN254
    rnd = random.Random(8)
    cols = [[rnd.randrange(F.r) for _ in range(ln)] for ln in (1, 2, 3, 5, 8, 13, 16)]
    polys = acg.create_polynomials_fft(ctx_bn, cols)
    for col, p in zip(cols, polys):
        assert p == O.fft_interpolate(F, col + [0] * (16 - len(col)))
    big = np.stack([acg.to_limbs([rnd.randrange(F.r) for _ in range(1 << 10)]) for _ in range(5)])
    out = ctx_bn.interpolate_columns(big)
    for i in range(5):
        assert (out[i] == CO.ntt(0, big[i], True)).all()


@pytest.mark.parametrize("fid,n,delta", [(0, 1000, (0, 0, 0)), (0, 1 << 12, (3, 5, 7)), (1, 3000, (0, 9, 0)), (0, 1, (1, 2, 3))])
def test_qap_witness_vs_c_oracle(acg, ctxs, fid, n, delta):
    ctx = ctxs[fid]
    g, w = acg.synth_r1cs(fid, n, 31 + n)
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    bufs, ok = ctx.qap_witness(m, dw, delta)
    ref = oracle_check(fid, g, w, True)
    N = O.next_pow2(n)
    pad = lambda v: np.vstack([v, np.zeros((N - n, 4), np.uint64)])
    a, b, c, h, rok = CO.qap_witness(fid, pad(ref["Aw"]), pad(ref["Bw"]), pad(ref["Cw"]), delta, 8)
    assert ok and rok
    for k, r in (("a", a), ("b", b), ("c", c), ("h", h)):
        assert (bufs[k] == r).all(), k
    # tampered witness: Nothing
    wb = w.copy()
    wb[g.n_cols - 1, 0] ^= np.uint64(1)
    dw.update(wb)
    assert ctx.qap_witness(m, dw, delta, want=())[1] is False


def test_qap_identity_at_random_point_2_16(acg, ctx_bn):
    """Size-independent property: a(z) * b(z) - c(z) == h(z) * (z^N - 1) at a random z (Schwartz-Zippel), with
    Horner evaluation in Python big ints over the returned coefficient vectors (N = 2^16; the 2^22 case is
    test_config3_qap_witness_2_22, bit-exact against the C oracle)."""
    n = 1 << 16
    F = O.BN254
    g, w = acg.synth_r1cs(0, n, 5)
    m, dw = ctx_bn.upload_r1cs(g), ctx_bn.upload_witness(w)
    bufs, ok = ctx_bn.qap_witness(m, dw, (11, 13, 17))
    assert ok
    z = random.Random(1).randrange(F.r)
    ev = {k: O.p_eval(F, acg.from_limbs(bufs[k]), z) for k in ("a", "b", "c", "h")}
    assert (ev["a"] * ev["b"] - ev["c"]) % F.r == ev["h"] * (pow(z, n, F.r) - 1) % F.r


# ------------------------------------------------------------------------------------------------ K5
def test_lagrange_golden(acg, ctxs):
    for case in golden("lagrange.json"):
        ctx = ctxs[case["field"]]
        polys, target = acg.create_polynomials(ctx, unhex(case["xs"]), [unhex(y) for y in case["ys"]])
        assert polys == [unhex(p) for p in case["polys"]]
        assert target == unhex(case["target"])


def test_lagrange_matches_reference_qap_build(acg, ctx_bn):
    """arithCircuitToQAP (Lagrange build, roots 7,8,9) on the unit-test circuit: target and a column vs KAT-1;
    on roots of unity it coincides with the FFT build."""
    F = O.BN254
    k = golden("kats.json")["kat1"]
    A = csr_from_json(k["A"])
    w = unhex(k["w"])
    aw = O.csr_matvec(F, A, w)
    polys, target = acg.create_polynomials(ctx_bn, [7, 8, 9], [aw])
    assert target == unhex(k["target"]) and polys[0] == unhex(k["a"])
    n = 64
    om = F.root_of_unity(6)
    xs = [pow(om, i, F.r) for i in range(n)]
    rnd = random.Random(6)
    ys = [rnd.randrange(F.r) for _ in range(n)]
    polys, target = acg.create_polynomials(ctx_bn, xs, [ys])
    assert polys[0] == O.fft_interpolate(F, ys) and target == [F.r - 1] + [0] * (n - 1) + [1]
    with pytest.raises(acg.AcgError):
        ctx_bn.lagrange([1, 2, 1], [[1, 2, 3]])
    big_n = 1024
    xs = rnd.sample(range(1, 1 << 40), big_n)
    ys = [rnd.randrange(F.r) for _ in range(big_n)]
    polys, _ = ctx_bn.lagrange(xs, [ys], False)
    for i in (0, 1, 500, big_n - 1):
        assert O.p_eval(F, polys[0], xs[i]) == ys[i]


def test_reference_gate_mix_split_rows_one_launch(acg, ctx_bn, tile_variant):
    """A circuit with the gate mix of the reference's own generator (Mul : Equal : Split = 50 : 10 : 1, 256-bit Split,
    test/Test/Circuit/Arithmetic.hs:77-126): every Split contributes one 256-entry row of general coefficients 2^i.
    However many there are, a check of the default geometry is ONE launch (the tiled kernel's warps claim the long rows
    after their tiles; other geometries: one warp-per-row launch first, then the tiles); counts, first bad row and the
    emitted A.w, B.w, C.w equal the C oracle's -- repeatedly (the claim counter is never reset) and for both forms."""
    n = 1 << 15
    g, w = acg.synth_mixed_r1cs(0, n, 99)
    lens = np.diff(g.mats[0][0].astype(np.int64))
    assert (lens > 8).sum() >= 50
    ref = oracle_check(0, g, w, True)
    assert ref["n_violations"] == 0
    m, dw = ctx_bn.upload_r1cs(g), ctx_bn.upload_witness(w)
    _select(acg, ctx_bn, "tiled", 0)
    for _ in range(3):
        assert ctx_bn.r1cs_check(m, dw) == (0, -1)
        assert ctx_bn.last_timing()["kernel_launches"] == (1 if tile_variant in (0, 8) else 2)
    aw, bw, cw = ctx_bn.r1cs_eval(m, dw)
    assert (aw == ref["Aw"]).all() and (bw == ref["Bw"]).all() and (cw == ref["Cw"]).all()
    other = 2 if tile_variant in (0, 8) else 0   # the other form of the check (geometry is bound at upload)
    ctx_bn.set_tiled_variant(other)
    m2 = ctx_bn.upload_r1cs(g)
    assert ctx_bn.r1cs_check(m2, dw) == (0, -1)
    assert ctx_bn.last_timing()["kernel_launches"] == (2 if other == 2 else 1)
    aw2, bw2, cw2 = ctx_bn.r1cs_eval(m2, dw)
    assert (aw2 == ref["Aw"]).all() and (bw2 == ref["Bw"]).all() and (cw2 == ref["Cw"]).all()
    ctx_bn.set_tiled_variant(tile_variant)
    # flip one output bit of every tenth Split and one Equal/Mul wire
    wb = w.copy()
    long_rows = np.nonzero(lens > 8)[0]
    for r in long_rows[::10]:
        col = int(g.mats[0][1][g.mats[0][0][r] + 17])    # the wire of bit 17 of that Split
        wb[col, 0] ^= np.uint64(1)
    wb[g.n_cols // 2, 0] += np.uint64(1)
    refb = oracle_check(0, g, wb)
    assert refb["n_violations"] > len(long_rows[::10])
    dw.update(wb)
    assert _both_kernels(acg, ctx_bn, lambda: ctx_bn.r1cs_check(m, dw)) == (refb["n_violations"], refb["first_bad_row"])
    assert ctx_bn.r1cs_check(m2, dw) == (refb["n_violations"], refb["first_bad_row"])
    for _ in range(3):
        assert ctx_bn.r1cs_check(m, dw) == (refb["n_violations"], refb["first_bad_row"])
    dw.free()
    m.free()
    m2.free()


@pytest.mark.parametrize("modulus_id", [2, 0, 1])
def test_linear_constraints_bulletproofs(acg, _ctx_bn, modulus_id):
    """checkLinearConstraint (src/Circuit/Bulletproofs.hs:329-338) over the scalar field of secp256k1 (a modulus >= 2^255,
    its own arithmetic on the device) and over the two fields of the main path: a batch of random sparse constraints
    made to hold by solving for the constant, some then broken; counts and first violated index against the big-int
    oracle; edge values (0, n - 1, weights hitting missing wires); a non-canonical weight is rejected."""
    ctx = _ctx_bn
    rnd = random.Random(100 + modulus_id)
    r = acg.modulus_of(modulus_id)
    assert r == (O.SECP256K1_N if modulus_id == 2 else FIELDS[modulus_id].r)
    n_gates, n_in = 300, 40
    edge = [0, 1, r - 1, r - 2, (1 << 255) % r, (r >> 1) + 1]
    val = lambda: rnd.choice(edge) if rnd.random() < 0.15 else rnd.randrange(r)
    asg = {"aL": {i: val() for i in range(n_gates) if rnd.random() < 0.95},      # some wires missing: they count as 0
           "aR": {i: val() for i in range(n_gates)}, "aO": {i: val() for i in range(n_gates)},
           "v": {i: val() for i in range(n_in)}}
    cons = []
    for i in range(1000):
        lc = {k: {rnd.randrange(n_gates): val() for _ in range(rnd.randrange(0, 6))} for k in ("wL", "wR", "wO")}
        lc["wV"] = {rnd.randrange(n_in): val() for _ in range(rnd.randrange(0, 3))}
        dot = lambda wgt, a: sum(c * a.get(ix, 0) for ix, c in wgt.items())
        lc["c"] = (dot(lc["wL"], asg["aL"]) + dot(lc["wR"], asg["aR"]) + dot(lc["wO"], asg["aO"]) - dot(lc["wV"], asg["v"])) % r
        if i % 97 == 5:
            lc["c"] = (lc["c"] + 1 + rnd.randrange(r - 1)) % r       # broken
        cons.append(lc)
    want = [O.check_linear_constraint(r, lc, asg) for lc in cons]
    bad = [i for i, ok in enumerate(want) if not ok]
    assert len(bad) == len([i for i in range(1000) if i % 97 == 5])
    assert acg.check_linear_constraints(ctx, modulus_id, cons, asg) == (len(bad), bad[0])
    good = [lc for lc, ok in zip(cons, want) if ok]
    assert acg.check_linear_constraints(ctx, modulus_id, good, asg) == (0, -1)
    assert acg.check_linear_constraint(ctx, modulus_id, good[0], asg) is True
    assert acg.check_linear_constraint(ctx, modulus_id, cons[bad[0]], asg) is False
    assert acg.check_linear_constraints(ctx, modulus_id, [], asg) == (0, -1)
    # the ABI rejects a weight >= n instead of reducing it
    import ctypes as C
    L = acg._lib.lib()
    rp = np.array([0, 1], np.uint32); col = np.array([0], np.uint32); wv = acg.to_limbs([r])
    zero_rp = np.array([0, 0], np.uint32)
    A = acg.qap.AcgCsr(rp.ctypes.data_as(acg._lib.u32p), col.ctypes.data_as(acg._lib.u32p), wv.ctypes.data_as(acg._lib.u64p), 1)
    B = acg.qap.AcgCsr(zero_rp.ctypes.data_as(acg._lib.u32p), col.ctypes.data_as(acg._lib.u32p), wv.ctypes.data_as(acg._lib.u64p), 0)
    one = acg.to_limbs([1])
    nv, fb = C.c_uint64(), C.c_uint64()
    rc = L.acg_linear_constraints_check(ctx._h, modulus_id, 1, 1, 0, C.byref(A), C.byref(B), acg.qap._ptr(one), acg.qap._ptr(one),
                                        None, C.byref(nv), C.byref(fb))
    assert rc == -2


def test_reference_fixture_replay_gpu(acg, _ctx_bn):
    """tools/replay_fixtures.py with the GPU path switched on, on the committed file in DumpFixtures.hs's format: QAP
    values (Lagrange and FFT builds), verifyAssignment, h and h with deltas from the device equal the file's."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("replay_fixtures", os.path.join(root, "tools", "replay_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(root, "tests", "golden", "reference_format_fixtures.json")) as f:
        fixtures = json.load(f)
    n, failures = mod.replay(fixtures, use_gpu=True)
    assert n >= 50 and not failures, failures


# ------------------------------------------------------------------------------------------------ BASELINE configs at full size
def _default_geometry(acg, ctx):
    ctx.set_tiled_variant(0)
    ctx.set_check_kernel(acg.CHECK_AUTO)


def test_config3_qap_witness_2_22(acg, _ctx_bn):
    """BASELINE configs[2]: S(2^22, 20260003, BN254) -- check + A.w, B.w, C.w + 3 inverse NTTs of 2^22 points + coset
    quotient: h, a, b, c of acg_qap_witness bit-exact against the C oracle's radix-2 NTT / coset division, with delta
    terms; a tampered witness is not divisible."""
    ctx = _ctx_bn
    _default_geometry(acg, ctx)
    n = 1 << 22
    g, w = acg.synth_r1cs(0, n, 20260003)
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    assert ctx.r1cs_check(m, dw) == (0, -1)
    ref = oracle_check(0, g, w, True, n_threads=16)
    aw, bw, cw = ctx.r1cs_eval(m, dw)
    assert (aw == ref["Aw"]).all() and (bw == ref["Bw"]).all() and (cw == ref["Cw"]).all()
    delta = (0x1234567, 3, F_R_MINUS_1)
    bufs, ok = ctx.qap_witness(m, dw, delta)
    a, b, c, h, rok = CO.qap_witness(0, ref["Aw"], ref["Bw"], ref["Cw"], delta, 16)
    assert ok and rok
    for k, want in (("h", h), ("a", a), ("b", b), ("c", c)):
        assert (bufs[k] == want).all(), k
    wb = w.copy()
    wb[1025 + n // 3, 0] += np.uint64(1)
    dw.update(wb)
    _bufs, ok = ctx.qap_witness(m, dw, (0, 0, 0), want=())
    assert not ok
    dw.free()
    m.free()


@pytest.mark.parametrize("variant", [0, 2])
def test_config5_bls12_381_2_20(acg, _ctx_bls, variant):
    """BASELINE configs[4]: S(2^20, 20260005, BLS12-381 Fr) -- honest and tampered witness, both kernels, against the
    C oracle's count and first bad row; emitted A.w, B.w, C.w bit-exact."""
    ctx = _ctx_bls
    ctx.set_tiled_variant(variant)
    n = 1 << 20
    g, w = acg.synth_r1cs(1, n, 20260005)
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    ref = oracle_check(1, g, w, True, n_threads=16)
    assert ref["n_violations"] == 0
    aw, bw, cw = ctx.r1cs_eval(m, dw)
    assert (aw == ref["Aw"]).all() and (bw == ref["Bw"]).all() and (cw == ref["Cw"]).all()
    wb = w.copy()
    wb[1025 + n // 3, 0] += np.uint64(1)
    for t in (7, 1025 + n - 2):
        wb[t, 3] ^= np.uint64(1 << 9)
    refb = oracle_check(1, g, wb, False, n_threads=16)
    assert refb["n_violations"] > 0
    for kernel, stages in KERNELS:
        _select(acg, ctx, kernel, stages)
        dw.update(w)
        assert ctx.r1cs_check(m, dw) == (0, -1)
        dw.update(wb)
        assert ctx.r1cs_check(m, dw) == (refb["n_violations"], refb["first_bad_row"])
    _reset(acg, ctx)
    ctx.set_tiled_variant(0)
    dw.free()
    m.free()


def test_config4_shards_of_2_24(acg, _ctx_bn):
    """BASELINE configs[3]: S(2^24, 20260004, BN254) split in 8 contiguous row blocks, witness replicated -- shards 0 and
    7 run on this GPU: per-shard violation count and first bad (global) row equal the C oracle's on the same shard,
    for the honest witness and for one tampered in both shards' reach."""
    from arithmetic_circuits_b200 import sharding
    ctx = _ctx_bn
    _default_geometry(acg, ctx)
    n = 1 << 24
    dw = None
    for rank in (0, 7):
        rb, re = sharding.row_shard(n, 8, rank)
        g, w = acg.synth_r1cs(0, n, 20260004, rows=(rb, re))
        assert g.n_rows == re - rb == n // 8 and g.n_cols == 1025 + n
        # upload the shard as the row range [rb, re) of the global system (first_bad_row is global)
        a, b, c = g.csr_structs()
        mats = [(np.concatenate([np.zeros(rb, np.uint32), rp, np.full(n - re, rp[-1], np.uint32)]), col, val)
                for rp, col, val in g.mats]
        gg = acg.GenQAP(0, n, g.n_cols, g.layout, mats)
        m = ctx.upload_r1cs(gg, rb, re)
        if dw is None:
            dw = ctx.upload_witness(w)
            wb = w.copy()
            wb[1025 + 12345, 0] += np.uint64(1)        # defined in shard 0, referenced far and wide afterwards
            wb[1025 + n - 5, 1] ^= np.uint64(1 << 20)  # defined in shard 7
        dw.update(w)
        assert ctx.r1cs_check(m, dw) == (0, -1)
        ref = oracle_check(0, g, wb, False, n_threads=16)
        dw.update(wb)
        nv, first = ctx.r1cs_check(m, dw)
        assert nv == ref["n_violations"] > 0
        assert first == rb + ref["first_bad_row"]
        m.free()
    dw.free()


def test_pipelined_witness_updates(acg, _ctx_bn):
    """acg_witness_update_async: two resident vectors, the upload of witness i + 1 (copy stream) overlaps the check of
    witness i; every check still sees exactly its own witness, a non-canonical element is reported by the check that
    reads the vector, and the blocking update leaves a vector untouched when it rejects."""
    import torch
    ctx = _ctx_bn
    _default_geometry(acg, ctx)
    n = 1 << 16
    g, w = acg.synth_r1cs(0, n, 777)
    m = ctx.upload_r1cs(g)
    wits, wants = [], []
    for i in range(6):
        wi = w.copy()
        if i % 3:
            wi[1025 + 1000 * i, 0] ^= np.uint64(1 << i)
        ref = oracle_check(0, g, wi)
        wits.append(torch.from_numpy(wi.view(np.int64)).pin_memory())
        wants.append((ref["n_violations"], ref["first_bad_row"]))
    vecs = [ctx.upload_witness(w), ctx.upload_witness(w)]
    vecs[0].update_async(wits[0].numpy().view(np.uint64))
    got = []
    for i in range(6):
        if i + 1 < 6:
            vecs[(i + 1) & 1].update_async(wits[i + 1].numpy().view(np.uint64))
        got.append(ctx.r1cs_check(m, vecs[i & 1]))
    assert got == wants
    # a non-canonical element: reported by the check of that vector, not earlier and not for the other vector
    bad = w.copy()
    bad[5] = np.array([0xFFFFFFFFFFFFFFFF] * 4, np.uint64)
    vecs[0].update_async(bad)
    vecs[1].update_async(w)
    assert ctx.r1cs_check(m, vecs[1]) == (0, -1)
    with pytest.raises(acg.AcgError) as e:
        ctx.r1cs_check(m, vecs[0])
    assert e.value.code == -2
    vecs[0].update_async(w)
    vecs[0].status()
    assert ctx.r1cs_check(m, vecs[0]) == (0, -1)
    # blocking update: rejected -> unchanged
    with pytest.raises(acg.AcgError) as e:
        vecs[0].update(bad)
    assert e.value.code == -2
    assert (vecs[0].download() == w).all() and ctx.r1cs_check(m, vecs[0]) == (0, -1)
    with pytest.raises(acg.AcgError):
        vecs[0].update_range(bad[:16], 0)
    assert (vecs[0].download() == w).all()
    for v in vecs:
        v.free()
    m.free()


# ------------------------------------------------------------------------------------------------ multi-GPU
def test_peer_exchange_two_ranks():
    """Row shards on two GPUs, result pair all-reduced over peer memory by the check kernel's last CTA: every rank
    sees the oracle's global count / first bad row (tests/peer_exchange_ranks.py under torchrun).  Needs >= 2 GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29537",
                        os.path.join(root, "tests", "peer_exchange_ranks.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "peer exchange ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_overlapping_consecutive_checks(acg, ctx_bn):
    """With acg_ctx_set_overlap_checks(1), back-to-back async checks of the same system are launched as programmatic dependents of each other
    (they overlap on the GPU and share one scratch result pair): every one of them must still deliver the oracle's
    count and first bad row -- for a clean witness, for a tampered one (violation reports from overlapping launches),
    after the witness changes in between (chain broken), and with the overlap switched off."""
    import torch
    n = 1 << 16
    g, w = acg.synth_r1cs(0, n, 31337)
    m, dw = ctx_bn.upload_r1cs(g), ctx_bn.upload_witness(w)
    stream = torch.cuda.current_stream()
    wb = w.copy()
    for t in (1025 + 5, 1025 + n // 2, 1025 + n - 3):
        wb[t, 1] ^= np.uint64(1 << 33)
    ref = oracle_check(0, g, wb)
    want_bad = (ref["n_violations"], ref["first_bad_row"])
    assert want_bad[0] > 0
    _select(acg, ctx_bn, "tiled", 0)
    for overlap in (True, False):
        ctx_bn.set_overlap_checks(overlap)
        for witness, want in ((w, (0, -1)), (wb, want_bad), (w, (0, -1))):
            dw.update(witness)
            results = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(12)]
            for r in results:   # no host synchronisation in between: launches 2..12 chain onto their predecessor
                ctx_bn.r1cs_check_async(m, dw, r.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            got = {(int(r[0].item()), int(r[1].item())) for r in results}
            assert got == {want}, (overlap, got, want)
        # two resident witnesses checked alternately, no host synchronisation: the chain holds across DIFFERENT witness
        # buffers (nothing was enqueued through the context in between) and every launch reports its own witness
        dwa, dwb = ctx_bn.upload_witness(w), ctx_bn.upload_witness(wb)
        results = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(12)]
        for i, r in enumerate(results):
            ctx_bn.r1cs_check_async(m, dwb if i & 1 else dwa, r.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        got = [(int(r[0].item()), int(r[1].item())) for r in results]
        assert got == [want_bad if i & 1 else (0, -1) for i in range(12)], (overlap, got)
        dwa.free()
        dwb.free()
    ctx_bn.set_overlap_checks(False)   # the default: a plain launch per check
    _reset(acg, ctx_bn)


def test_direct_and_ticket_handover_agree(acg, _ctx_bn):
    """kernels.h CheckEpilogue: a plain check hands its result over directly (block 0 opens a gate, violating warps
    update the result themselves, nothing happens at the end of a clean check); overlapped checks go through the
    scratch pair and the last CTA's ticket.  Clean, a few violated rows, violations in every tile: both are what the
    oracle says, repeatedly (the gate ring and the scratch pair are reused from check to check)."""
    ctx = _ctx_bn
    _default_geometry(acg, ctx)
    n = 1 << 18          # 2048 tiles: more than the 740 CTAs of the grid, so the runs are planned
    g, w = acg.synth_r1cs(0, n, 4242)
    m = ctx.upload_r1cs(g)
    few = w.copy()
    few[1025 + n // 5, 0] ^= np.uint64(2)
    few[1025 + n // 2 + 77, 2] ^= np.uint64(1 << 9)
    many = w.copy()
    many[1025:1025 + n // 2:37, 1] ^= np.uint64(1)
    cases = [w, few, many]
    wants = []
    for wi in cases:
        ref = oracle_check(0, g, wi)
        wants.append((ref["n_violations"], ref["first_bad_row"]))
    assert wants[0] == (0, -1) and 0 < wants[1][0] < 16 and wants[2][0] > 2048
    vecs = [ctx.upload_witness(wi) for wi in cases]
    for _ in range(3):
        assert [ctx.r1cs_check(m, v) for v in vecs] == wants
    ctx.set_overlap_checks(True)      # chained checks take the ticket path
    try:
        for v, want in zip(vecs, wants):
            assert [ctx.r1cs_check(m, v) for _ in range(3)] == [want] * 3
    finally:
        ctx.set_overlap_checks(False)
    assert [ctx.r1cs_check(m, v) for v in reversed(vecs)] == list(reversed(wants))
    # more checks than the gate ring has words
    for i in range(600):
        assert ctx.r1cs_check(m, vecs[i % 3]) == wants[i % 3]
    for x in vecs:
        x.free()
    m.free()
