"""JSON interchange in the reference's aeson encodings.

Every type on the hot path derives ToJSON / FromJSON generically (aeson `defaultOptions`: sum types as
{"tag": constructor, "contents": payload} -- a list when the constructor has several unnamed fields -- or, for record
constructors, {"tag": constructor, field: value, ...}; single-constructor records as plain objects; newtypes as their
payload; `Map Int v` as an object with decimal string keys):

    Wire             src/Circuit/Arithmetic.hs:32-36     {"tag": "InputWire", "contents": 3}
    AffineCircuit    src/Circuit/Affine.hs:26-31         {"tag": "Add", "contents": [l, r]} / {"tag": "ScalarMul",
                                                         "contents": [f, c]} / {"tag": "ConstGate", "contents": f} /
                                                         {"tag": "Var", "contents": wire}
    Gate             src/Circuit/Arithmetic.hs:44-59     {"tag": "Mul", "mulLeft": .., "mulRight": .., "mulOutput": ..},
                                                         {"tag": "Equal", "eqInput", "eqMagic", "eqOutput"},
                                                         {"tag": "Split", "splitInput", "splitOutputs": [..]}
    ArithCircuit     src/Circuit/Arithmetic.hs:149-150   [gate, ...]                      (newtype over the list)
    QapSet f         src/QAP.hs:66-71                    {"qapSetConstant": f, "qapSetInput": {"0": f, ..},
                                                         "qapSetIntermediate": {..}, "qapSetOutput": {..}}
    QAP f            src/QAP.hs:74-79                    {"qapInputsLeft": QapSet [f], "qapInputsRight": ..,
                                                         "qapOutputs": .., "qapTarget": [f]}
    Prime r          src/QAP.hs:87-90                    the canonical residue `fromP` as a JSON number (arbitrary size)
    VPoly f          src/QAP.hs:82-85                    `unPoly`: little-endian coefficient list, no trailing zeros

This is what lets a machine WITH GHC dump fixtures (ArithCircuit, inputs, QapSet, QAP) that this repository ingests and
checks on the GPU -- the route to pinning value-level parity against real reference output.  The encodings above are
restated from aeson's documented generic defaults; no reference binary could be run here to confirm them byte for byte
("parity unpinned"), so the reader is also lenient where aeson versions differ (it accepts a
2-element list for a Map given as a list of pairs).  Host-side only; nothing here touches the GPU."""
from __future__ import annotations

import json
from typing import Any, Dict, Iterable, List, Mapping, Sequence, Tuple

from . import qap as Q

_WIRE_TAGS = ("InputWire", "IntermediateWire", "OutputWire")


# ---- Wire ------------------------------------------------------------------------------------------------
def wire_to_json(w: int) -> Dict[str, Any]:
    return {"tag": _WIRE_TAGS[w >> 62], "contents": int(w & ((1 << 62) - 1))}


def wire_from_json(j: Mapping[str, Any]) -> int:
    kind = _WIRE_TAGS.index(j["tag"])
    ix = int(j["contents"])
    if ix < 0:
        raise ValueError("negative wire index")
    return (kind << 62) | ix


# ---- AffineCircuit ---------------------------------------------------------------------------------------
def affine_to_json(c) -> Dict[str, Any]:
    """Iterative (unsplit chains are 256 deep, src/Circuit/Arithmetic.hs:238-244)."""
    out: List[Any] = []
    stack: List[Tuple[Any, int]] = [(c, 0)]
    while stack:
        node, state = stack.pop()
        tag = node[0]
        if tag == "var":
            out.append({"tag": "Var", "contents": wire_to_json(node[1])})
        elif tag == "const":
            out.append({"tag": "ConstGate", "contents": int(node[1])})
        elif tag == "add":
            if state == 0:
                stack += [(node, 1), (node[2], 0), (node[1], 0)]
            else:
                r, l = out.pop(), out.pop()
                out.append({"tag": "Add", "contents": [l, r]})
        elif tag == "scalar":
            if state == 0:
                stack += [(node, 1), (node[2], 0)]
            else:
                out.append({"tag": "ScalarMul", "contents": [int(node[1]), out.pop()]})
        else:
            raise ValueError("bad affine node %r" % (tag,))
    return out[0]


def affine_from_json(j: Mapping[str, Any]):
    out: List[Any] = []
    stack: List[Tuple[Any, int]] = [(j, 0)]
    while stack:
        node, state = stack.pop()
        tag = node["tag"]
        if tag == "Var":
            out.append(Q.Var(wire_from_json(node["contents"])))
        elif tag == "ConstGate":
            out.append(Q.ConstGate(int(node["contents"])))
        elif tag == "Add":
            if state == 0:
                l, r = node["contents"]
                stack += [(node, 1), (r, 0), (l, 0)]
            else:
                r, l = out.pop(), out.pop()
                out.append(Q.Add(l, r))
        elif tag == "ScalarMul":
            if state == 0:
                stack += [(node, 1), (node["contents"][1], 0)]
            else:
                out.append(Q.ScalarMul(int(node["contents"][0]), out.pop()))
        else:
            raise ValueError("bad AffineCircuit tag %r" % (tag,))
    return out[0]


# ---- Gate / ArithCircuit ---------------------------------------------------------------------------------
def gate_to_json(g) -> Dict[str, Any]:
    if g[0] == "mul":
        return {"tag": "Mul", "mulLeft": affine_to_json(g[1]), "mulRight": affine_to_json(g[2]),
                "mulOutput": wire_to_json(g[3])}
    if g[0] == "equal":
        return {"tag": "Equal", "eqInput": wire_to_json(g[1]), "eqMagic": wire_to_json(g[2]),
                "eqOutput": wire_to_json(g[3])}
    if g[0] == "split":
        return {"tag": "Split", "splitInput": wire_to_json(g[1]), "splitOutputs": [wire_to_json(w) for w in g[2]]}
    raise ValueError("bad gate %r" % (g[0],))


def gate_from_json(j: Mapping[str, Any]):
    tag = j["tag"]
    if tag == "Mul":
        return Q.Mul(affine_from_json(j["mulLeft"]), affine_from_json(j["mulRight"]), wire_from_json(j["mulOutput"]))
    if tag == "Equal":
        return Q.Equal(wire_from_json(j["eqInput"]), wire_from_json(j["eqMagic"]), wire_from_json(j["eqOutput"]))
    if tag == "Split":
        return Q.Split(wire_from_json(j["splitInput"]), [wire_from_json(w) for w in j["splitOutputs"]])
    raise ValueError("bad Gate tag %r" % (tag,))


def circuit_to_json(gates: Sequence) -> List[Dict[str, Any]]:
    """ArithCircuit f -> JSON value (a list: the newtype is transparent)."""
    return [gate_to_json(g) for g in gates]


def circuit_from_json(field: int, j: Iterable[Mapping[str, Any]]) -> "Q.ArithCircuit":
    return Q.ArithCircuit(field, [gate_from_json(g) for g in j])


# ---- QapSet ----------------------------------------------------------------------------------------------
def _int_map_to_json(m: Mapping[int, Any], enc) -> Dict[str, Any]:
    return {str(int(k)): enc(m[k]) for k in sorted(m)}


def _int_map_from_json(j, dec) -> Dict[int, Any]:
    if isinstance(j, Mapping):
        return {int(k): dec(v) for k, v in j.items()}
    return {int(k): dec(v) for k, v in j}  # list-of-pairs form


def qapset_to_json(constant, inputs: Mapping[int, Any], intermediates: Mapping[int, Any], outputs: Mapping[int, Any],
                   enc=int) -> Dict[str, Any]:
    return {"qapSetConstant": enc(constant), "qapSetInput": _int_map_to_json(inputs, enc),
            "qapSetIntermediate": _int_map_to_json(intermediates, enc), "qapSetOutput": _int_map_to_json(outputs, enc)}


def qapset_from_json(j: Mapping[str, Any], dec=int):
    """-> (constant, inputs, intermediates, outputs)"""
    return (dec(j["qapSetConstant"]), _int_map_from_json(j["qapSetInput"], dec),
            _int_map_from_json(j["qapSetIntermediate"], dec), _int_map_from_json(j["qapSetOutput"], dec))


def assignment_to_json(a: "Q.QapSet") -> Dict[str, Any]:
    """The witness QapSet f of generateAssignment (src/QAP.hs:597-603)."""
    n_in, n_mid, n_out = a.dims()
    sets: List[Dict[int, int]] = []
    for kind, n in ((0, n_in), (1, n_mid), (2, n_out)):
        d = {}
        for i in range(n):
            v = a.lookup((kind << 62) | i)
            if v is not None:
                d[i] = v
        sets.append(d)
    return qapset_to_json(1, sets[0], sets[1], sets[2])


def witness_vector_from_json(j: Mapping[str, Any], layout: Tuple[int, int, int] = None):
    """QapSet f (JSON) -> dense witness vector in qapSetToMap order (src/QAP.hs:605-620) as (n_cols, 4) limbs.
    layout = (num inputs, num intermediates, num outputs); default: max key + 1 of each map, as the reference does."""
    const, ins, mids, outs = qapset_from_json(j)
    n_in, n_mid, n_out = layout if layout is not None else tuple((max(m) + 1 if m else 0) for m in (ins, mids, outs))
    vals = [0] * (1 + n_in + n_mid + n_out)
    vals[0] = const
    for base, m in ((1, ins), (1 + n_in, mids), (1 + n_in + n_mid, outs)):
        for k, v in m.items():
            vals[base + k] = v
    return Q.to_limbs(vals)


# ---- QAP -------------------------------------------------------------------------------------------------
def _poly_enc(p: Sequence[int]) -> List[int]:
    return [int(c) for c in Q.strip(p)]


def _poly_dec(j: Sequence[Any]) -> List[int]:
    return Q.strip([int(c) for c in j])


def qap_to_json(left, right, out, target) -> Dict[str, Any]:
    """left / right / out = (constant poly, {i: poly}, {i: poly}, {i: poly}); polys little-endian coefficient lists."""
    return {"qapInputsLeft": qapset_to_json(*left, enc=_poly_enc), "qapInputsRight": qapset_to_json(*right, enc=_poly_enc),
            "qapOutputs": qapset_to_json(*out, enc=_poly_enc), "qapTarget": _poly_enc(target)}


def qap_from_json(j: Mapping[str, Any]):
    return (qapset_from_json(j["qapInputsLeft"], _poly_dec), qapset_from_json(j["qapInputsRight"], _poly_dec),
            qapset_from_json(j["qapOutputs"], _poly_dec), _poly_dec(j["qapTarget"]))


def dumps(value) -> str:
    """Compact text as aeson's `encode` writes it (no spaces); integers of any size are plain JSON numbers."""
    return json.dumps(value, separators=(",", ":"))


def loads(text: str):
    return json.loads(text)
