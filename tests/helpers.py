"""Shared test helpers: golden loading, JSON circuit -> oracle form / product form."""
import json
import os

import numpy as np

from oracle import qap_oracle as O
from oracle import c_oracle as CO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = {0: O.BN254, 1: O.BLS12_381}


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def unhex(xs):
    return [int(x, 16) for x in xs]


def csr_from_json(j):
    return O.CSR(list(j["rowptr"]), list(j["col"]), unhex(j["val"]))


def csr_np(M):
    """oracle CSR -> numpy triple for the C oracle / product"""
    return (np.array(M.rowptr, np.uint32), np.array(M.col, np.uint32), CO.ints_to_limbs(M.val))


def gates_oracle(jg):
    def aff(c):
        if c[0] == "var":
            return O.Var(tuple(c[1]))
        if c[0] == "const":
            return O.ConstGate(int(c[1], 16))
        if c[0] == "add":
            return O.Add(aff(c[1]), aff(c[2]))
        return O.ScalarMul(int(c[1], 16), aff(c[2]))
    out = []
    for g in jg:
        if g[0] == "mul":
            out.append(O.Mul(aff(g[1]), aff(g[2]), tuple(g[3])))
        elif g[0] == "equal":
            out.append(O.Equal(tuple(g[1]), tuple(g[2]), tuple(g[3])))
        else:
            out.append(O.Split(tuple(g[1]), [tuple(o) for o in g[2]]))
    return out


def gates_acg(acg, gates):
    """oracle-form gates -> product-form gates"""
    wk = {"in": acg.InputWire, "mid": acg.IntermediateWire, "out": acg.OutputWire}
    W = lambda w: wk[w[0]](w[1])

    def aff(c):
        stack, out = [(c, False)], []
        while stack:
            node, done = stack.pop()
            if node[0] == "var":
                out.append(acg.Var(W(node[1])))
            elif node[0] == "const":
                out.append(acg.ConstGate(node[1]))
            elif node[0] == "add":
                if done:
                    r = out.pop(); l = out.pop()
                    out.append(acg.Add(l, r))
                else:
                    stack += [(node, True), (node[2], False), (node[1], False)]
            else:
                if done:
                    out.append(acg.ScalarMul(node[1], out.pop()))
                else:
                    stack += [(node, True), (node[2], False)]
        return out[0]
    res = []
    for g in gates:
        if g[0] == "mul":
            res.append(acg.Mul(aff(g[1]), aff(g[2]), W(g[3])))
        elif g[0] == "equal":
            res.append(acg.Equal(W(g[1]), W(g[2]), W(g[3])))
        else:
            res.append(acg.Split(W(g[1]), [W(o) for o in g[2]]))
    return res


def genqap_equals_csr(acg, g, A, B, C):
    for (rp, col, val), M in zip(g.mats, (A, B, C)):
        assert rp.tolist() == M.rowptr
        assert col.tolist() == M.col
        assert acg.from_limbs(val) == M.val


def make_genqap(acg, field, n_rows, n_cols, layout, A, B, C):
    """Build a product GenQAP from oracle CSR matrices (keeps numpy arrays alive)."""
    mats = [csr_np(M) for M in (A, B, C)]
    return acg.GenQAP(field, n_rows, n_cols, layout, mats)


def oracle_check(field, g, w, want_vectors=False, n_threads=4):
    """C-oracle check of a product GenQAP + witness (numpy limbs)."""
    A, B, C = [(m[0], m[1], m[2]) for m in g.mats]
    return CO.r1cs_eval_check(field, g.n_rows, g.n_cols, A, B, C, w, want_vectors, n_threads)
