"""The DEVICE field algorithms of csrc/fr.cuh (split even/odd accumulator CIOS, carry chains) executed on
the CPU against an emulated carry flag, compared with Python big integers.  This is the same instruction
sequence ptxas compiles for sm_100a, so arithmetic bugs surface without a GPU."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import qap_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "arithmetic-circuits_b200")


@pytest.fixture(scope="module")
def hostlib():
    out = os.path.join(PKG, "_build", "libacg_hosttest.so")
    src = os.path.join(PKG, "csrc", "host", "fr_host_test.cpp")
    deps = [src, os.path.join(PKG, "csrc", "fr.cuh"), os.path.join(PKG, "csrc", "fr_constants.inc")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out, src])
    return C.CDLL(out)


def _run(lib, acg, fid, op, xs, ys):
    a, b = acg.to_limbs(xs), acg.to_limbs(ys)
    o = np.empty_like(a)
    rc = lib.acg_hosttest_binop(fid, op, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                o.ctypes.data_as(C.c_void_p), C.c_uint64(len(xs)))
    assert rc == 0
    return acg.from_limbs(o)


@pytest.mark.parametrize("F", [O.BN254, O.BLS12_381], ids=lambda f: f.name)
def test_device_field_algorithms_on_host(hostlib, acg, F):
    rnd = random.Random(5)
    r, R = F.r, 1 << 256
    edge = [0, 1, 2, r - 1, r - 2, F.mont_R, (r - F.mont_R) % r, (1 << 32) - 1, (1 << 64) - 1, 1 << 253, r >> 1,
            (r >> 1) + 1, (1 << 224) - 1, (1 << 32), (1 << 96) + 5]
    xs = [rnd.randrange(r) for _ in range(5000)] + [e for e in edge for _ in edge]
    ys = [rnd.randrange(r) for _ in range(5000)] + [e for _ in edge for e in edge]
    fid = F.field_id
    assert _run(hostlib, acg, fid, 0, xs, ys) == [(x + y) % r for x, y in zip(xs, ys)]
    assert _run(hostlib, acg, fid, 1, xs, ys) == [(x - y) % r for x, y in zip(xs, ys)]
    assert _run(hostlib, acg, fid, 2, xs, ys) == [(x * y) % r for x, y in zip(xs, ys)]
    Rinv = pow(R, -1, r)
    assert _run(hostlib, acg, fid, 6, xs, ys) == [(x * y * Rinv) % r for x, y in zip(xs, ys)]
    assert _run(hostlib, acg, fid, 4, xs, ys) == [(x * R) % r for x in xs]
    assert _run(hostlib, acg, fid, 5, xs, ys) == [(-x) % r for x in xs]
    # operand contract the tiled kernel's lazy row sums rely on: vector operand <= p, scalar operand any
    # 256-bit value; fr_add on [0, p] inputs stays in [0, p]
    lim = R - 1
    wide = [rnd.randrange(R) for _ in range(3000)] + [lim, lim - 1, r, 2 * r, R - r, 0] * 3
    vec = [rnd.randrange(r) for _ in range(3000)] + [r, r - 1, 0, r, 1, r] * 3
    assert _run(hostlib, acg, fid, 7, vec, wide) == [(x * y * Rinv) % r for x, y in zip(vec, wide)]
    got = _run(hostlib, acg, fid, 8, vec, [r] * len(vec))
    assert all(g % r == v % r and g <= r for g, v in zip(got, vec))
    inv_in = xs[:50] + edge
    assert _run(hostlib, acg, fid, 3, inv_in, inv_in) == [pow(x, -1, r) if x else 0 for x in inv_in]


def test_host_field_matches_device_algorithms(hostlib, acg):
    """host/fr_host.hpp (64-bit limbs, used for per-gate host logic) agrees with fr.cuh through the public
    entry points: witness of a squaring chain."""
    F = O.BN254
    gates = [acg.Mul(acg.Var(acg.InputWire(0)), acg.Var(acg.InputWire(0)), acg.IntermediateWire(0))]
    for i in range(1, 40):
        gates.append(acg.Mul(acg.Var(acg.IntermediateWire(i - 1)), acg.Add(acg.Var(acg.IntermediateWire(i - 1)), acg.ConstGate(i)),
                             acg.IntermediateWire(i)))
    c = acg.ArithCircuit(0, gates)
    x = 0x1234567890abcdef1234567890abcdef1234567890abcdef
    a = acg.generate_assignment(c, {0: x})
    v = x * x % F.r
    for i in range(1, 40):
        v = v * (v + i) % F.r
    assert a.lookup(acg.IntermediateWire(39)) == v
