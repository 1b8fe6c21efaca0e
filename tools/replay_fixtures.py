#!/usr/bin/env python3
"""Replay reference-held value vectors (tools/DumpFixtures.hs, run on a machine WITH GHC against the unmodified
reference) through this repository and compare every value bit for bit.

    python tools/replay_fixtures.py reference_fixtures.json            # oracle + GPU path (needs a CUDA device)
    python tools/replay_fixtures.py reference_fixtures.json --no-gpu   # CPU only: pins the ORACLE against the reference
    python tools/replay_fixtures.py --emit-from-oracle out.json        # the same file format, produced by the oracle
                                                                       # (self-test of the format and of this tool)

This is the route to lifting the "parity unpinned" caveat of DESIGN.md section 3: the image of this repository has no
Haskell toolchain, so the reference's own values have to come from outside.  What is compared per fixture: the
assignment of generateAssignment, its qapSetToMap vector, every wire polynomial and the target of the QAP value
(arithCircuitToQAP / arithCircuitToQAPFFT), verifyAssignment, verificationWitness (h) and verificationWitnessZk 3 5 7,
plus getRootOfUnity 0..28."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _norm_qapset(j, dec=int):
    """QapSet JSON -> (constant, {i: v}, {i: v}, {i: v}) with plain ints / coefficient lists"""
    from arithmetic_circuits_b200 import json_io as J
    return J.qapset_from_json(j, dec)


def _strip(p):
    p = [int(c) for c in p]
    while p and p[-1] == 0:
        p.pop()
    return p


def oracle_fixture(F, name, gates_json, inputs, roots, lagrange):
    """One fixture object in DumpFixtures.hs's shape, computed by the CPU oracle."""
    from oracle import qap_oracle as O
    from arithmetic_circuits_b200 import json_io as J
    gates = _oracle_gates(gates_json)
    asg = O.generate_assignment(F, gates, inputs)
    qap = O.arith_circuit_to_qap(F, roots, gates) if lagrange else O.arith_circuit_to_qap_fft(F, roots, gates)
    h = O.verification_witness(F, qap, asg)
    hz = O.verification_witness_zk(F, 3, 5, 7, qap, asg)[0]
    enc = lambda qs: J.qapset_to_json(O.p_norm(F, qs.constant), {k: O.p_norm(F, v) for k, v in qs.inputs.items()},
                                      {k: O.p_norm(F, v) for k, v in qs.mids.items()},
                                      {k: O.p_norm(F, v) for k, v in qs.outputs.items()}, enc=list)
    return {"name": name, "circuit": gates_json, "inputs": {str(k): v for k, v in inputs.items()}, "roots": roots,
            "build": "arithCircuitToQAP" if lagrange else "arithCircuitToQAPFFT getRootOfUnity",
            "assignment": J.qapset_to_json(1, asg.inputs, asg.mids, asg.outputs),
            "qap": {"qapInputsLeft": enc(qap.left), "qapInputsRight": enc(qap.right), "qapOutputs": enc(qap.out),
                    "qapTarget": O.p_norm(F, qap.target)},
            "verify": h is not None, "h": None if h is None else O.p_norm(F, h),
            "h_zk_3_5_7": None if hz is None else O.p_norm(F, hz),
            "witness": {str(k): v for k, v in O.qap_set_to_map(asg).items()}}


def _oracle_gates(gates_json):
    """aeson circuit JSON -> oracle gates"""
    from oracle import qap_oracle as O
    wire = lambda j: ({"InputWire": "in", "IntermediateWire": "mid", "OutputWire": "out"}[j["tag"]], int(j["contents"]))

    def aff(j):
        t = j["tag"]
        if t == "Var":
            return O.Var(wire(j["contents"]))
        if t == "ConstGate":
            return O.ConstGate(int(j["contents"]))
        if t == "Add":
            return O.Add(aff(j["contents"][0]), aff(j["contents"][1]))
        return O.ScalarMul(int(j["contents"][0]), aff(j["contents"][1]))
    out = []
    for g in gates_json:
        if g["tag"] == "Mul":
            out.append(O.Mul(aff(g["mulLeft"]), aff(g["mulRight"]), wire(g["mulOutput"])))
        elif g["tag"] == "Equal":
            out.append(O.Equal(wire(g["eqInput"]), wire(g["eqMagic"]), wire(g["eqOutput"])))
        else:
            out.append(O.Split(wire(g["splitInput"]), [wire(o) for o in g["splitOutputs"]]))
    return out


def emit_from_oracle(path):
    from oracle import qap_oracle as O
    F = O.BN254
    W = lambda t, i: {"tag": t, "contents": i}
    var = lambda t, i: {"tag": "Var", "contents": W(t, i)}
    mul = lambda l, r, o: {"tag": "Mul", "mulLeft": l, "mulRight": r, "mulOutput": o}
    kat1 = [mul(var("InputWire", 0), var("InputWire", 1), W("IntermediateWire", 0)),
            mul(var("InputWire", 2), var("InputWire", 3), W("IntermediateWire", 1)),
            mul({"tag": "Add", "contents": [{"tag": "ConstGate", "contents": 10}, var("IntermediateWire", 0)]},
                var("IntermediateWire", 1), W("OutputWire", 0))]
    kat3 = [mul(var("InputWire", 0), var("InputWire", 1), W("IntermediateWire", 0)),
            mul(var("IntermediateWire", 0), {"tag": "Add", "contents": [var("InputWire", 0), var("InputWire", 2)]},
                W("OutputWire", 0))]
    in1, in3 = {0: 2, 1: 3, 2: 4, 3: 5}, {0: 7, 1: 5, 2: 4}
    out = [oracle_fixture(F, "kat1_lagrange_roots_7_8_9", kat1, in1, [[7], [8], [9]], True),
           oracle_fixture(F, "kat1_fft_roots_1_2_3", kat1, in1, [[1], [2], [3]], False),
           oracle_fixture(F, "kat3_bench_fft_roots_0_1", kat3, in3, [[0], [1]], False),
           oracle_fixture(F, "kat3_example_fft_roots_1_2", kat3, in3, [[1], [2]], False),
           oracle_fixture(F, "kat3_lagrange_roots_0_1", kat3, in3, [[0], [1]], True),
           {"name": "roots_of_unity", "getRootOfUnity": [F.root_of_unity(k) for k in range(29)]}]
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    return out


def replay(fixtures, use_gpu):
    import arithmetic_circuits_b200 as acg
    from arithmetic_circuits_b200 import json_io as J
    from oracle import qap_oracle as O
    F = O.BN254
    ctx = acg.Context(acg.BN254_FR, 0) if use_gpu else None
    n_checked, failures = 0, []

    def check(what, got, want):
        nonlocal n_checked
        n_checked += 1
        if got != want:
            failures.append("%s: got %r..., reference %r..." % (what, str(got)[:120], str(want)[:120]))

    for fx in fixtures:
        name = fx["name"]
        if "getRootOfUnity" in fx:
            ref = [int(v) for v in fx["getRootOfUnity"]]
            check(name + " (oracle)", [F.root_of_unity(k) for k in range(len(ref))], ref)
            check(name + " (library)", [acg.get_root_of_unity(0, k) for k in range(len(ref))], ref)
            continue
        inputs = {int(k): int(v) for k, v in fx["inputs"].items()} if isinstance(fx["inputs"], dict) else \
            {int(k): int(v) for k, v in fx["inputs"]}
        roots = [[int(r) for r in per_gate] for per_gate in fx["roots"]]
        lagrange = fx["build"].startswith("arithCircuitToQAP") and "FFT" not in fx["build"]
        ref_asg = _norm_qapset(fx["assignment"])
        ref_qap = J.qap_from_json(fx["qap"])
        ref_h = None if fx["h"] is None else _strip(fx["h"])
        ref_hz = None if fx["h_zk_3_5_7"] is None else _strip(fx["h_zk_3_5_7"])
        wit = fx["witness"]
        ref_w = {int(k): int(v) for k, v in (wit.items() if isinstance(wit, dict) else wit)}
        # ---- the oracle against the reference
        gates = _oracle_gates(fx["circuit"])
        oa = O.generate_assignment(F, gates, inputs)
        check(name + ": oracle generateAssignment", (oa.constant, oa.inputs, oa.mids, oa.outputs), ref_asg)
        check(name + ": oracle qapSetToMap", O.qap_set_to_map(oa), ref_w)
        oq = O.arith_circuit_to_qap(F, roots, gates) if lagrange else O.arith_circuit_to_qap_fft(F, roots, gates)
        norm = lambda qs: (O.p_norm(F, qs.constant), {k: O.p_norm(F, v) for k, v in qs.inputs.items()},
                           {k: O.p_norm(F, v) for k, v in qs.mids.items()}, {k: O.p_norm(F, v) for k, v in qs.outputs.items()})
        check(name + ": oracle QAP value", (norm(oq.left), norm(oq.right), norm(oq.out), O.p_norm(F, oq.target)), ref_qap)
        oh = O.verification_witness(F, oq, oa)
        check(name + ": oracle verifyAssignment", oh is not None, bool(fx["verify"]))
        check(name + ": oracle h", None if oh is None else O.p_norm(F, oh), ref_h)
        ohz = O.verification_witness_zk(F, 3, 5, 7, oq, oa)[0]
        check(name + ": oracle h (deltas 3 5 7)", None if ohz is None else O.p_norm(F, ohz), ref_hz)
        # ---- the library (host mirror + GPU) against the reference
        circuit = J.circuit_from_json(0, fx["circuit"])
        a = acg.generate_assignment(circuit, inputs)
        check(name + ": library generateAssignment", _norm_qapset(J.assignment_to_json(a)), ref_asg)
        if not use_gpu:
            continue
        q = acg.arith_circuit_to_qap(ctx, circuit, roots) if lagrange else acg.arith_circuit_to_qap_fft(ctx, circuit, roots)
        check(name + ": library QAP value", J.qap_from_json(J.qap_to_json(*q.sets(), q.target)), ref_qap)
        check(name + ": library verifyAssignment", acg.verify_assignment_qap(ctx, q, a), bool(fx["verify"]))
        check(name + ": library h", acg.verification_witness_zk_qap(ctx, 0, 0, 0, q, a), ref_h)
        check(name + ": library h (deltas 3 5 7)", acg.verification_witness_zk_qap(ctx, 3, 5, 7, q, a), ref_hz)
    if ctx is not None:
        ctx.close()
    return n_checked, failures


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("fixtures", nargs="?")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--emit-from-oracle", metavar="PATH")
    args = ap.parse_args()
    if args.emit_from_oracle:
        emit_from_oracle(args.emit_from_oracle)
        print("wrote", args.emit_from_oracle)
        if not args.fixtures:
            return 0
    with open(args.fixtures) as f:
        fixtures = json.load(f)
    n, failures = replay(fixtures, not args.no_gpu)
    for msg in failures:
        print("MISMATCH", msg)
    print("%d values compared, %d mismatches" % (n, len(failures)))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
