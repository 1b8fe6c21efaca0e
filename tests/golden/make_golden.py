#!/usr/bin/env python3
"""Generates tests/golden/*.json from the Python big-int oracle (oracle/qap_oracle.py), which restates the
reference's algorithms line by line.  The reference itself (Haskell) cannot run in this image, so these
vectors are NOT reference output; they pin (1) the outcomes the reference's own tests assert
(test/Test/QAP.hs:68-90, test/Test/Circuit/Arithmetic.hs:154-182, Example.hs) and (2) the oracle's values,
which SURVEY.md section 8c derived independently (KAT-1..6) -- any drift in oracle, C oracle, C++ host
mirror or CUDA path shows up against these files.

    python tests/golden/make_golden.py        # rewrites the JSON files deterministically
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qap_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
hx = lambda v: "%x" % v
hxl = lambda vs: [hx(v) for v in vs]


def csr_json(M):
    return {"rowptr": M.rowptr, "col": M.col, "val": hxl(M.val)}


def kat_circuits():
    F = O.BN254
    out = {}
    # KAT-1 / KAT-2: test/Test/QAP.hs:48-90
    gates = [O.Mul(O.Var(O.inw(0)), O.Var(O.inw(1)), O.midw(0)),
             O.Mul(O.Var(O.inw(2)), O.Var(O.inw(3)), O.midw(1)),
             O.Mul(O.Add(O.ConstGate(10), O.Var(O.midw(0))), O.Var(O.midw(1)), O.outw(0))]
    roots = [[7], [8], [9]]
    qap = O.arith_circuit_to_qap(F, roots, gates)
    asg = O.generate_assignment(F, gates, {0: 2, 1: 3, 2: 4, 3: 5})
    h, a, b, c, rem = O.verification_witness_zk(F, 0, 0, 0, qap, asg)
    bad = O.QapSet(1, {0: 2, 1: 3, 2: 4, 3: 5}, {0: 7, 1: 20}, {0: 320})
    gq = O.arith_circuit_to_gen_qap(F, roots, gates)
    lay = O.layout_of(asg)
    A, B, C, rts = O.gen_qap_to_csr(F, gq, lay)
    w, wbad = O.witness_vector(F, asg, lay), O.witness_vector(F, bad, lay)
    out["kat1"] = {
        "roots": [7, 8, 9], "inputs": {"0": 2, "1": 3, "2": 4, "3": 5},
        "mids": hxl([asg.mids[0], asg.mids[1]]), "out": hx(asg.outputs[0]),
        "target": hxl(qap.target), "a": hxl(a), "b": hxl(b), "c": hxl(c), "h": hxl(h), "rem": hxl(rem),
        "verify": O.verify_assignment(F, qap, asg),
        "A": csr_json(A), "B": csr_json(B), "C": csr_json(C), "w": hxl(w),
        "residuals": hxl(O.r1cs_residuals(F, A, B, C, w)[0]),
    }
    out["kat2"] = {
        "w": hxl(wbad), "verify": O.verify_assignment(F, qap, bad),
        "residuals": hxl(O.r1cs_residuals(F, A, B, C, wbad)[0]), "check": list(O.r1cs_check(F, A, B, C, wbad)),
    }
    # KAT-3: bench/Circuit.hs:17-24, Example.hs
    g3 = [O.Mul(O.Var(O.inw(0)), O.Var(O.inw(1)), O.midw(0)),
          O.Mul(O.Var(O.midw(0)), O.Add(O.Var(O.inw(0)), O.Var(O.inw(2))), O.outw(0))]
    a3 = O.generate_assignment(F, g3, {0: 7, 1: 5, 2: 4})
    res = {}
    for name, start, mk in (("bench_fft", 0, O.arith_circuit_to_qap_fft), ("bench_lagrange", 0, O.arith_circuit_to_qap),
                            ("example_fft", 1, O.arith_circuit_to_qap_fft)):
        q3 = mk(F, O.fresh_roots(g3, start), g3)
        hh, aa, bb, cc, _ = O.verification_witness_zk(F, 0, 0, 0, q3, a3)
        res[name] = {"verify": hh is not None, "target": hxl(q3.target), "a": hxl(aa), "b": hxl(bb), "c": hxl(cc),
                     "h": hxl(hh)}
    out["kat3"] = {"mid0": hx(a3.mids[0]), "out0": hx(a3.outputs[0]), "qaps": res}
    # KAT-4: unit_eqGate
    eq = [O.Equal(O.inw(0), O.midw(0), O.outw(0))]
    out["kat4"] = {str(v): hx(O.generate_assignment(F, eq, {0: v}).outputs[0]) for v in (0, 1, 2, 3)}
    # KAT-6 constants
    out["kat6"] = {}
    for Fx in (O.BN254, O.BLS12_381):
        out["kat6"][Fx.name] = {"r": hx(Fx.r), "R": hx(Fx.mont_R), "R2": hx(Fx.mont_R2), "ninv64": hx(Fx.mont_ninv64),
                                "roots_of_unity": {str(k): hx(Fx.root_of_unity(k)) for k in (1, 2, 3, Fx.two_adicity)}}
    return out


def mixed_circuit(F, rnd, n_inputs=4, n_gates=12, with_split=True, nbits=256):
    """Random circuit in the spirit of arbArithCircuit (test/Test/Circuit/Arithmetic.hs:77-126):
    Mul / Equal / Split gates over earlier wires."""
    gates, mids = [], []

    def aff(depth):
        if depth <= 0:
            ch = rnd.randrange(3)
            if ch == 0 or (ch == 2 and not mids):
                return O.Var(O.inw(rnd.randrange(n_inputs)))
            if ch == 1:
                return O.ConstGate(rnd.randrange(F.r))
            return O.Var(O.midw(rnd.choice(mids)))
        if rnd.randrange(2):
            return O.ScalarMul(rnd.randrange(F.r), aff(depth - 1))
        return O.Add(aff(depth - 1), aff(depth - 1))
    nxt = 0
    for gi in range(n_gates):
        kind = rnd.choices(("mul", "equal", "split"), (50, 10, 3))[0]
        forced = {2: "zero", 3: "equal_zero", 5: "equal", 7: "split" if with_split else "mul"}.get(gi)
        if kind == "split" and not with_split:
            kind = "mul"
        if forced == "zero":        # a wire that evaluates to 0, so Equal sees both branches
            gates.append(O.Mul(O.ConstGate(0), aff(1), O.midw(nxt)))
            zero_wire = nxt
            mids.append(nxt)
            nxt += 1
            continue
        if forced == "equal_zero":
            gates.append(O.Equal(O.midw(zero_wire), O.midw(nxt), O.midw(nxt + 1)))
            mids += [nxt + 1]
            nxt += 2
            continue
        if forced:
            kind = forced
        if kind == "mul" or not mids:
            gates.append(O.Mul(aff(rnd.randrange(3)), aff(rnd.randrange(3)), O.midw(nxt)))
            mids.append(nxt)
            nxt += 1
        elif kind == "equal":
            gates.append(O.Equal(O.midw(rnd.choice(mids)), O.midw(nxt), O.midw(nxt + 1)))
            mids += [nxt + 1]   # the magic wire is not an output wire (outputWires, Arithmetic.hs:67-71)
            nxt += 2
        else:
            outs = [O.midw(nxt + i) for i in range(nbits)]
            gates.append(O.Split(O.midw(rnd.choice(mids)), outs))
            mids += list(range(nxt, nxt + nbits))
            nxt += nbits
    return gates


def gates_json(gates):
    def aff(c):
        if c[0] == "var":
            return ["var", list(c[1])]
        if c[0] == "const":
            return ["const", hx(c[1])]
        if c[0] == "add":
            return ["add", aff(c[1]), aff(c[2])]
        return ["scalar", hx(c[1]), aff(c[2])]
    out = []
    for g in gates:
        if g[0] == "mul":
            out.append(["mul", aff(g[1]), aff(g[2]), list(g[3])])
        elif g[0] == "equal":
            out.append(["equal", list(g[1]), list(g[2]), list(g[3])])
        else:
            out.append(["split", list(g[1]), [list(o) for o in g[2]]])
    return out


def mixed_cases():
    cases = []
    for F, seed in ((O.BN254, 11), (O.BN254, 12), (O.BLS12_381, 13)):
        rnd = random.Random(seed)
        gates = mixed_circuit(F, rnd, with_split=(seed != 11))   # Split is only satisfiable with >= field-width bits
        inputs = {i: rnd.randrange(F.r) for i in range(4)}
        if seed == 12:
            inputs[0] = 0  # exercises Equal on zero when reachable
        asg = O.generate_assignment(F, gates, inputs)
        roots = O.fresh_roots(gates, 1)
        gq = O.arith_circuit_to_gen_qap(F, roots, gates, densify=False)
        lay = O.layout_of(asg)
        A, B, C, rts = O.gen_qap_to_csr(F, gq, lay)
        w = O.witness_vector(F, asg, lay)
        res, aw, bw, cw = O.r1cs_residuals(F, A, B, C, w)
        assert not any(res)
        # tampered witness: +1 on a wire in the middle
        wb = list(w)
        wb[len(w) // 2] = (wb[len(w) // 2] + 1) % F.r
        case = {"field": F.field_id, "gates": gates_json(gates), "inputs": {str(k): hx(v) for k, v in inputs.items()},
                "layout": [lay.n_in, lay.n_mid, lay.n_out], "w": hxl(w), "A": csr_json(A), "B": csr_json(B),
                "C": csr_json(C), "Aw": hxl(aw), "Bw": hxl(bw), "Cw": hxl(cw),
                "bad_w": hxl(wb), "bad_check": list(O.r1cs_check(F, A, B, C, wb))}
        if len(res) <= 64:
            a, b, c, h, ok = O.qap_witness_ntt(F, aw, bw, cw, 3, 5, 7)
            case["qap_delta_3_5_7"] = {"a": hxl(a), "b": hxl(b), "c": hxl(c), "h": hxl(h), "ok": ok}
        cases.append(case)
    return cases


def synth_cases():
    cases = []
    for F, n, seed, dense in ((O.BN254, 96, 20260002, False), (O.BN254, 64, 7, True), (O.BLS12_381, 96, 20260005, False)):
        A, B, C, w, lay = O.synth_r1cs(F, n, seed, dense)
        res, aw, bw, cw = O.r1cs_residuals(F, A, B, C, w)
        assert not any(res)
        wb = list(w)
        wb[1025 + n // 3] = (wb[1025 + n // 3] + 1) % F.r
        a, b, c, h, ok = O.qap_witness_ntt(F, aw, bw, cw)
        cases.append({"field": F.field_id, "n": n, "seed": seed, "dense": dense, "A": csr_json(A), "B": csr_json(B),
                      "C": csr_json(C), "w_tail": hxl(w[1025:]), "w_head": hxl(w[:8]), "Aw": hxl(aw), "Bw": hxl(bw),
                      "Cw": hxl(cw), "bad_check": list(O.r1cs_check(F, A, B, C, wb)),
                      "h": hxl(h), "a": hxl(a), "ok": ok})
    return cases


def ntt_cases():
    out = []
    for F in (O.BN254, O.BLS12_381):
        rnd = random.Random(99 + F.field_id)
        for log_n in (1, 2, 5, 8):
            v = [rnd.randrange(F.r) for _ in range(1 << log_n)]
            om = F.root_of_unity(log_n)
            out.append({"field": F.field_id, "log_n": log_n, "in": hxl(v), "fwd": hxl(O.ntt(F, v, om)),
                        "inv": hxl(O.intt(F, v, om))})
    return out


def lagrange_cases():
    out = []
    for F in (O.BN254, O.BLS12_381):
        rnd = random.Random(5 + F.field_id)
        for n in (1, 2, 3, 17):
            xs = rnd.sample(range(1, 1000), n) if n < 17 else [rnd.randrange(F.r) for _ in range(n)]
            ys = [[rnd.randrange(F.r) for _ in range(n)] for _ in range(3)]
            polys = [O.lagrange_interpolate(F, list(zip(xs, y))) for y in ys]
            target = [1]
            for x in xs:
                target = O.p_mul(F, target, [(-x) % F.r, 1])
            out.append({"field": F.field_id, "xs": hxl(xs), "ys": [hxl(y) for y in ys], "polys": [hxl(p) for p in polys],
                        "target": hxl(target)})
    return out


def main():
    files = {"kats.json": kat_circuits(), "mixed_circuits.json": mixed_cases(), "synth.json": synth_cases(),
             "ntt.json": ntt_cases(), "lagrange.json": lagrange_cases()}
    for name, obj in files.items():
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f, separators=(",", ":"), sort_keys=True)
            f.write("\n")
        print(name, os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    main()
