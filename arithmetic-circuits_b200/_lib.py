"""ctypes binding of include/acg.h (libacg.so).  The library is the product; this file only declares
prototypes.  Loading fails loudly when libacg.so has not been built -- there is no Python or CPU
fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacg.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p


class AcgCsr(C.Structure):
    _fields_ = [("rowptr", u32p), ("col", u32p), ("val", u64p), ("nnz", C.c_uint64)]


class AcgTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32)]


# name -> (restype, argtypes); mirrors include/acg.h one to one
PROTOTYPES = {
    "acg_abi_version": (C.c_int, []),
    "acg_strerror": (C.c_char_p, [C.c_int]),
    "acg_last_error": (C.c_char_p, [vp]),
    "acg_ctx_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(vp)]),
    "acg_ctx_destroy": (None, [vp]),
    "acg_ctx_set_check_kernel": (C.c_int, [vp, C.c_int]),
    "acg_ctx_set_overlap_checks": (C.c_int, [vp, C.c_int]),
    "acg_ctx_set_tiled_variant": (C.c_int, [vp, C.c_int]),
    "acg_r1cs_stream_bytes": (C.c_uint64, [vp]),
    "acg_r1cs_set_row_offset": (C.c_int, [vp, C.c_uint64]),
    "acg_tile_stream_digest": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(AcgCsr), C.POINTER(AcgCsr),
                                        C.POINTER(AcgCsr), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "acg_last_timing": (C.c_int, [vp, C.POINTER(AcgTiming)]),
    "acg_kernel_launch_count": (C.c_uint64, [vp]),
    "acg_profile_begin": (C.c_int, [vp, C.c_uint32]),
    "acg_profile_end": (C.c_int, [vp, C.POINTER(C.c_float), C.c_uint32, u32p]),
    "acg_field_constants": (C.c_int, [C.c_int, u64p, u64p, u64p, u64p, u32p]),
    "acg_root_of_unity": (C.c_int, [C.c_int, C.c_uint32, u64p]),
    "acg_r1cs_upload": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(AcgCsr), C.POINTER(AcgCsr),
                                  C.POINTER(AcgCsr), C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "acg_r1cs_free": (None, [vp]),
    "acg_r1cs_algorithmic_bytes": (C.c_uint64, [vp]),
    "acg_witness_upload": (C.c_int, [vp, vp, C.c_uint32, C.POINTER(vp)]),
    "acg_witness_update": (C.c_int, [vp, vp, vp, C.c_uint32]),
    "acg_witness_update_range": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32]),
    "acg_witness_update_async": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32]),
    "acg_vec_status": (C.c_int, [vp, vp]),
    "acg_vec_stream_wait": (C.c_int, [vp, vp, vp]),
    "acg_vec_free": (None, [vp]),
    "acg_vec_len": (C.c_uint32, [vp]),
    "acg_vec_device_ptr": (vp, [vp]),
    "acg_peer_create": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(vp), vp]),
    "acg_peer_connect": (C.c_int, [vp, vp, vp]),
    "acg_peer_free": (None, [vp]),
    "acg_r1cs_check_async_allreduce": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "acg_poly_combine": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp]),
    "acg_qap_upload": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.POINTER(vp)]),
    "acg_qap_free": (None, [vp]),
    "acg_qap_quotient_len": (C.c_uint32, [vp]),
    "acg_qap_verify": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32, u32p, C.POINTER(C.c_int)]),
    "acg_fft_target": (C.c_int, [vp, C.c_uint32, vp]),
    "acg_circuit_plan_stats": (C.c_int, [vp, u32p, u32p]),
    "acg_vec_download": (C.c_int, [vp, vp, vp, C.c_uint32]),
    "acg_generate_assignment_device": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.POINTER(vp), u32p]),
    "acg_r1cs_check": (C.c_int, [vp, vp, vp, u64p, u64p]),
    "acg_r1cs_check_async": (C.c_int, [vp, vp, vp, vp, vp]),
    "acg_r1cs_check_host": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(AcgCsr), C.POINTER(AcgCsr),
                                      C.POINTER(AcgCsr), vp, u64p, u64p]),
    "acg_r1cs_eval": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "acg_ntt": (C.c_int, [vp, vp, C.c_uint32, C.c_int]),
    "acg_ntt_device": (C.c_int, [vp, vp, C.c_uint32, C.c_int, vp]),
    "acg_interpolate_columns": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32]),
    "acg_qap_witness": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int)]),
    "acg_lagrange": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp]),
    "acg_fr_binop": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_uint64]),
    "acg_linear_constraints_check": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(AcgCsr),
                                               C.POINTER(AcgCsr), vp, vp, vp, u64p, u64p]),
    "acg_circuit_parse": (C.c_int, [C.c_int, vp, C.c_uint64, C.POINTER(vp)]),
    "acg_circuit_free": (None, [vp]),
    "acg_circuit_num_gates": (C.c_uint64, [vp]),
    "acg_circuit_valid": (C.c_int, [vp]),
    "acg_circuit_num_roots": (C.c_uint64, [vp]),
    "acg_generate_assignment": (C.c_int, [vp, vp, vp, C.c_uint32, C.POINTER(vp)]),
    "acg_assignment_free": (None, [vp]),
    "acg_assignment_dims": (C.c_int, [vp, u32p, u32p, u32p]),
    "acg_assignment_lookup": (C.c_int, [vp, C.c_uint64, u64p]),
    "acg_assignment_update": (C.c_int, [vp, C.c_uint64, u64p]),
    "acg_assignment_to_vector": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "acg_circuit_to_r1cs": (C.c_int, [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "acg_r1cs_host_free": (None, [vp]),
    "acg_r1cs_host_dims": (C.c_int, [vp, u32p, u32p, u32p, u32p, u32p]),
    "acg_r1cs_host_csr": (C.c_int, [vp, C.c_int, C.POINTER(AcgCsr)]),
    "acg_r1cs_host_roots": (u64p, [vp]),
    "acg_synth_r1cs": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(vp), C.POINTER(u64p)]),
    "acg_synth_r1cs_rows": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(vp),
                                      C.POINTER(u64p)]),
    "acg_synth_circuit_words": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(u64p), u64p,
                                          C.POINTER(u32p), C.POINTER(u64p), u32p]),
    "acg_free": (None, [vp]),
}

_lib = None


class AcgLibraryMissing(RuntimeError):
    pass


def lib():
    """The loaded libacg.so.  Raises AcgLibraryMissing if it was not built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AcgLibraryMissing(
                "%s not found: build it with `python arithmetic-circuits_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)  # AttributeError here means the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib
