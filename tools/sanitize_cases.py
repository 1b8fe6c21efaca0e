#!/usr/bin/env python3
"""The small configurations run under compute-sanitizer (tools/r02_sanitize.sh): every kernel family of the path on
2^12 .. 2^16 rows, including the 12-deep chain of overlapping (programmatic dependent) check launches, the warp-per-row
long-row kernel, device witness generation, the NTT / coset quotient pipeline, Lagrange and the QAP division."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import arithmetic_circuits_b200 as acg  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = acg.Context(acg.BN254_FR, 0)
stream = torch.cuda.current_stream()


def k2(n, variant):
    ctx.set_tiled_variant(variant)
    g, w = acg.synth_r1cs(0, n, 4242 + n)
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    for kern in (acg.CHECK_TILED, acg.CHECK_ROWWISE):
        ctx.set_check_kernel(kern)
        assert ctx.r1cs_check(m, dw) == (0, -1)
    ctx.set_check_kernel(acg.CHECK_TILED)
    wb = w.copy()
    wb[1025 + n // 3, 0] += np.uint64(1)
    dwb = ctx.upload_witness(wb)
    bad = ctx.r1cs_check(m, dwb)
    assert bad[0] > 0
    # violated rows in (nearly) every tile: every CTA reports through the gate of the direct hand-over at once
    wm = w.copy()
    wm[1025:1025 + n // 2:19, 1] ^= np.uint64(1)
    dwm = ctx.upload_witness(wm)
    many = ctx.r1cs_check(m, dwm)
    ctx.set_check_kernel(acg.CHECK_ROWWISE)
    assert many[0] > n // 128 and ctx.r1cs_check(m, dwm) == many
    ctx.set_check_kernel(acg.CHECK_TILED)
    assert ctx.r1cs_check(m, dwm) == many and ctx.r1cs_check(m, dw) == (0, -1)
    dwm.free()
    # 12 back-to-back launches, overlapping (PDL chain), alternating the two witnesses
    ctx.set_overlap_checks(True)
    res = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(12)]
    for i, r in enumerate(res):
        ctx.r1cs_check_async(m, dwb if i & 1 else dw, r.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    ctx.set_overlap_checks(False)
    assert [int(r[0]) for r in res] == [bad[0] if i & 1 else 0 for i in range(12)]
    a, b, c = ctx.r1cs_eval(m, dw)
    for x in (dw, dwb, m):
        x.free()


if which in ("all", "k2"):
    for n, v in ((1 << 12, 0), (1 << 16, 0), (1 << 17, 0), (1 << 14, 2), (1 << 13, 7), (1 << 13, 8)):   # 2^17: planned runs
        k2(n, v)
    print("k2 ok")
if which in ("all", "mix"):
    ctx.set_tiled_variant(0)
    g, w = acg.synth_mixed_r1cs(0, 1 << 13, 5)
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    assert ctx.r1cs_check(m, dw) == (0, -1)
    ctx.r1cs_eval(m, dw)
    # pipelined witness updates (copy stream)
    dw2 = ctx.upload_witness(w)
    for i in range(4):
        (dw2 if i & 1 else dw).update_async(w)
        assert ctx.r1cs_check(m, dw2 if i & 1 else dw) == (0, -1)
    print("mix / long rows / async update ok")
if which in ("all", "qap"):
    g, w = acg.synth_r1cs(0, 1 << 12, 9)
    m, dw = ctx.upload_r1cs(g), ctx.upload_witness(w)
    bufs, ok = ctx.qap_witness(m, dw, (3, 5, 7))
    assert ok
    circuit, inputs = acg.synth_circuit(0, 300, 11)
    dwg, _ = ctx.generate_assignment(circuit, inputs)
    gq = acg.arith_circuit_to_gen_qap(circuit, [[1000 + 7 * i] for i in range(300)])
    a = acg.generate_assignment(circuit, inputs)
    for q in (acg.create_polynomials_qap(ctx, gq), acg.create_polynomials_fft_qap(ctx, gq)):
        assert acg.verification_witness_zk_qap(ctx, 1, 2, 3, q, a) is not None
        q.free_device()
    print("qap / ntt / lagrange / division / witness generation ok")
ctx.close()
