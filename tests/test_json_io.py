"""JSON interchange: the aeson generic encodings of the reference's types, restated.
Golden strings are written by hand from aeson's `defaultOptions` rules for the derived instances at
src/Circuit/Arithmetic.hs:36,59,150, src/Circuit/Affine.hs:31 and src/QAP.hs:71,79,82-90."""
import json

import numpy as np

from arithmetic_circuits_b200 import json_io as J

R_BN = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _kat_gates(acg):
    # test/Test/QAP.hs:48-62
    return [acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(1)), acg.IntermediateWire(0)),
            acg.Mul(acg.Var(acg.InputWire(2)), acg.Var(acg.InputWire(3)), acg.IntermediateWire(1)),
            acg.Mul(acg.Add(acg.ConstGate(10), acg.Var(acg.IntermediateWire(0))), acg.Var(acg.IntermediateWire(1)),
                    acg.OutputWire(0))]


GOLDEN_GATE2 = ('{"tag":"Mul","mulLeft":{"tag":"Add","contents":[{"tag":"ConstGate","contents":10},'
                '{"tag":"Var","contents":{"tag":"IntermediateWire","contents":0}}]},'
                '"mulRight":{"tag":"Var","contents":{"tag":"IntermediateWire","contents":1}},'
                '"mulOutput":{"tag":"OutputWire","contents":0}}')


def test_gate_encoding_matches_aeson_rules(acg):
    gates = _kat_gates(acg)
    assert J.dumps(J.gate_to_json(gates[2])) == GOLDEN_GATE2
    assert J.dumps(J.gate_to_json(acg.Equal(acg.InputWire(0), acg.IntermediateWire(0), acg.OutputWire(0)))) == \
        ('{"tag":"Equal","eqInput":{"tag":"InputWire","contents":0},"eqMagic":{"tag":"IntermediateWire","contents":0},'
         '"eqOutput":{"tag":"OutputWire","contents":0}}')
    assert J.dumps(J.gate_to_json(acg.Split(acg.InputWire(1), [acg.OutputWire(0), acg.OutputWire(1)]))) == \
        ('{"tag":"Split","splitInput":{"tag":"InputWire","contents":1},"splitOutputs":'
         '[{"tag":"OutputWire","contents":0},{"tag":"OutputWire","contents":1}]}')
    big = R_BN - 1
    s = J.dumps(J.affine_to_json(acg.ScalarMul(big, acg.Var(acg.InputWire(7)))))
    assert s == '{"tag":"ScalarMul","contents":[%d,{"tag":"Var","contents":{"tag":"InputWire","contents":7}}]}' % big


def test_circuit_round_trip_and_semantics(acg):
    gates = _kat_gates(acg) + [acg.Equal(acg.OutputWire(0), acg.IntermediateWire(2), acg.OutputWire(1)),
                               acg.Split(acg.InputWire(0), [acg.IntermediateWire(3 + i) for i in range(8)]),
                               acg.Mul(acg.unsplit([acg.IntermediateWire(3 + i) for i in range(8)]),
                                       acg.ConstGate(1), acg.OutputWire(2))]
    text = J.dumps(J.circuit_to_json(gates))
    back = J.circuit_from_json(acg.BN254_FR, J.loads(text))
    ref = acg.ArithCircuit(acg.BN254_FR, gates)
    assert (back.words == ref.words).all() and back.valid() == ref.valid()
    assert J.dumps(J.circuit_to_json(back.gates)) == text
    # the re-read circuit evaluates like the original (generateAssignment)
    inputs = {0: 200, 1: 3, 2: 4, 3: 5}
    a, b = acg.generate_assignment(ref, inputs), acg.generate_assignment(back, inputs)
    assert (a.to_vector() == b.to_vector()).all()
    assert a.lookup(acg.OutputWire(2)) == 200


def test_qapset_encoding(acg):
    circuit = acg.ArithCircuit(acg.BN254_FR, _kat_gates(acg))
    a = acg.generate_assignment(circuit, {0: 2, 1: 3, 2: 4, 3: 5})
    j = J.assignment_to_json(a)
    assert J.dumps(j) == ('{"qapSetConstant":1,"qapSetInput":{"0":2,"1":3,"2":4,"3":5},'
                          '"qapSetIntermediate":{"0":6,"1":20},"qapSetOutput":{"0":320}}')
    w = J.witness_vector_from_json(J.loads(J.dumps(j)))
    assert (w == a.to_vector()).all()
    # list-of-pairs maps are accepted on input
    j2 = dict(j, qapSetInput=[[0, 2], [1, 3], [2, 4], [3, 5]])
    assert (J.witness_vector_from_json(j2) == w).all()
    # huge residues survive as JSON numbers
    t = J.loads(J.dumps(J.qapset_to_json(1, {0: R_BN - 1}, {}, {})))
    assert J.qapset_from_json(t)[1][0] == R_BN - 1


def test_qap_encoding_strips_trailing_zeros():
    left = ([1, 0, 0], {0: [0, 5, 0]}, {}, {0: []})
    j = J.qap_to_json(left, left, left, [R_BN - 504, 191, R_BN - 24, 1, 0])
    assert j["qapInputsLeft"]["qapSetConstant"] == [1] and j["qapInputsLeft"]["qapSetInput"]["0"] == [0, 5]
    assert j["qapTarget"] == [R_BN - 504, 191, R_BN - 24, 1]
    l2, r2, o2, t2 = J.qap_from_json(json.loads(J.dumps(j)))
    assert l2 == ([1], {0: [0, 5]}, {}, {0: []}) and t2 == [R_BN - 504, 191, R_BN - 24, 1]


def test_json_round_trip_random_circuits(acg):
    """Property: circuit -> aeson JSON text -> circuit is the identity on the marshalled word stream, for random
    Mul / Equal / Split circuits with random affine trees and residues of any size below r."""
    from hypothesis import given, settings, strategies as st

    wires = st.one_of(st.builds(acg.InputWire, st.integers(0, 40)), st.builds(acg.IntermediateWire, st.integers(0, 40)),
                      st.builds(acg.OutputWire, st.integers(0, 40)))
    elems = st.integers(0, R_BN - 1)
    affine = st.recursive(st.one_of(st.builds(acg.Var, wires), st.builds(acg.ConstGate, elems)),
                          lambda inner: st.one_of(st.builds(acg.Add, inner, inner), st.builds(acg.ScalarMul, elems, inner)),
                          max_leaves=6)
    gate = st.one_of(st.builds(acg.Mul, affine, affine, wires), st.builds(acg.Equal, wires, wires, wires),
                     st.builds(acg.Split, wires, st.lists(wires, min_size=1, max_size=5)))

    @settings(max_examples=60, deadline=None)
    @given(st.lists(gate, min_size=1, max_size=8))
    def check(gates):
        text = J.dumps(J.circuit_to_json(gates))
        back = [J.gate_from_json(g) for g in J.loads(text)]
        assert back == [tuple(g) if not isinstance(g, tuple) else g for g in gates] or \
            J.dumps(J.circuit_to_json(back)) == text
        a = acg.ArithCircuit(acg.BN254_FR, gates)
        b = acg.ArithCircuit(acg.BN254_FR, back)
        assert (a.words == b.words).all()

    check()


def test_reference_fixture_replay_cpu():
    """tools/replay_fixtures.py on a file in the exact format tools/DumpFixtures.hs writes (here produced by the oracle:
    tests/golden/reference_format_fixtures.json): the reader understands the aeson encodings of circuit, inputs, roots,
    assignment, QAP value and h, and oracle and host mirror agree with every value in it.  With a file dumped by the real
    reference this is the test that pins value-level parity."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("replay_fixtures", os.path.join(root, "tools", "replay_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(root, "tests", "golden", "reference_format_fixtures.json")) as f:
        fixtures = json.load(f)
    n, failures = mod.replay(fixtures, use_gpu=False)
    assert n >= 30 and not failures, failures
    # a corrupted value is caught
    fixtures[0]["h"][0] += 1
    n, failures = mod.replay(fixtures, use_gpu=False)
    assert len(failures) == 1 and "oracle h" in failures[0]
