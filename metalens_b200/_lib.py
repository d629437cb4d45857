"""ctypes binding of libmetalens_b200.so (include/metalens_b200.h).

There is NO fallback: if the shared library is missing or a call fails, the
product path raises.  Build the library with ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C metalens_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmetalens_b200.so")


class MetalensB200Error(RuntimeError):
    pass


_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every symbol declared in include/metalens_b200.h
SIGNATURES = {
    "mlb_version": (C.c_int, []),
    "mlb_last_error": (C.c_char_p, []),
    "mlb_device_caps": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "mlb_launch_count": (C.c_longlong, []),
    "mlb_twiddle_build": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                    C.c_void_p, C.c_int, C.c_void_p]),
    "mlb_cgemm_tn": (C.c_int, [_PP, C.c_int, C.c_void_p, C.c_int, _PP, C.c_int,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mlb_fold": (C.c_int, [_PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                           _PP, C.c_int, C.c_int, C.c_void_p]),
    "mlb_tf32_split": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mlb_twiddle_tf32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p]),
    "mlb_cgemm_tc": (C.c_int, [_PP, _PP, C.c_int, _PP, _PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               _PP, _PP, C.c_int, C.c_int, C.c_void_p]),
    "mlb_cgemm_tc_split": (C.c_int, [_PP, _PP, C.c_int, _PP, _PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     _PP, _PP, C.c_int, C.c_int, _PP, C.c_int, C.c_void_p]),
    "mlb_fft_twiddle": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_fft_max_length": (C.c_int, []),
    "mlb_fft_tune": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "mlb_fft_rows_can_transpose": (C.c_int, [C.c_int]),
    "mlb_fft_rows": (C.c_int, [_PP, C.c_int, _PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mlb_fft_mixed_plan": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_fft_mixed_compiled": (C.c_int, [C.c_int, C.c_int]),
    "mlb_fft_cols": (C.c_int, [_PP, C.c_int, _PP, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                               C.c_void_p]),
    "mlb_fft_cols_power_blocks": (C.c_int, [C.c_int, C.c_int]),
    "mlb_fft_cols_power": (C.c_int, [_PP, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, _PP, C.c_int, C.c_void_p]),
    "mlb_fft_cols_power_total": (C.c_int, [_PP, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "mlb_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "mlb_get_option": (C.c_int, [C.c_char_p]),
    "mlb_czt_chirps": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "mlb_czt_pointwise": (C.c_int, [_PP, C.c_int, _PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mlb_ff_epilogue_blocks": (C.c_int, [C.c_int, C.c_int]),
    "mlb_ff_epilogue": (C.c_int, [_PP, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_cone_power": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                 C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlb_sum_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "mlb_struct_sizes": (C.c_int, [C.c_void_p]),
    "mlb_nearfield_blocks": (C.c_int, [C.c_int, C.c_int]),
    "mlb_nearfield_tune": (C.c_int, [C.c_int]),
    "mlb_nearfield_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mlb_nearfield_ring_table_floats": (C.c_longlong, [C.c_void_p]),
    "mlb_nearfield_assemble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_nearfield_assemble_ties": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_void_p]),
    "mlb_nearfield_fixup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_table_pack_build": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlb_table_eval": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "mlb_hex_count": (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_hex_fill": (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double,
                               C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_cells_bin": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    # 8e: multi-GPU exchange steps (peer memory over NVLink; NCCL wrappers)
    "mlb_peer_flag_words": (C.c_int, []),
    "mlb_peer_state_words": (C.c_int, []),
    "mlb_fft_rows_scatter": (C.c_int, [_PP, C.c_int, _PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]),
    "mlb_peer_barrier": (C.c_int, [_PP, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_peer_allgather": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, _PP, C.c_longlong, C.c_longlong,
                                     C.c_void_p, _PP, C.c_int, C.c_int, _PP, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p]),
    "mlb_peer_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlb_peer_wait_sum": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "mlb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "mlb_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "mlb_comm_destroy": (C.c_int, [C.c_void_p]),
    "mlb_allgather_P": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mlb_allgather_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mlb_allreduce_scalar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
}

_lib = None


def load():
    """Load the C-ABI library once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MetalensB200Error(
            "metalens_b200: %s not found. The CUDA extension is the product; there is no CPU "
            "fallback. Build it with `make -C metalens_b200/csrc` (needs nvcc, sm_100a)." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the .so disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().mlb_last_error().decode("utf-8", "replace")
        raise MetalensB200Error("%s failed (rc=%d): %s" % (what, rc, msg))


def ptr_array(tensors):
    """void*[len] of device pointers (host array, as the C-ABI's h_ arguments expect); entries are
    tensors or raw integer addresses (peer-mapped buffers)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t if isinstance(t, int) else t.data_ptr()
    return C.cast(arr, _PP), arr
