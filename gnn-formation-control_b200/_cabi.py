"""ctypes binding of libgfc.so (C ABI declared in include/gfc.h).

The reference's only FFI is ``ct.CDLL(remoteApi.so)`` with int32 status returns
(sim.py:21); this binding follows the same convention: every entry point returns
an int, 0 = ok, and :func:`check` turns anything else into a ``RuntimeError``
carrying ``gfc_last_error()``.  There is no CPU fallback: if the library has not
been built the import fails loudly.
"""
import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GFC_LIB") or os.path.join(_HERE, "libgfc.so")   # GFC_LIB: an instrumented build (tools/)

GFC_OK, GFC_ERR_BAD_ARG, GFC_ERR_UNSUPPORTED, GFC_ERR_WORKSPACE, GFC_ERR_CUDA, GFC_ERR_TIMEOUT = range(6)
GSO_BINARY_LE, GSO_SYM_NORM_LT, GSO_BINARY_LT = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU = 0, 1, 2
PREC_FP32_3XTF32, PREC_TF32, PREC_F16 = 0, 1, 2
PREC_FLAG_BINARY_GSO = 0x100
PREC_FLAG_SHARED_GSO = 0x200
PATH_TILE, PATH_WORKSPACE, PATH_TCGEN05_WIDE = 1, 2, 3
OPT_SKIP_GRAD_REDUCE = 1
OPT_DISABLE_TCGEN05 = 2
OPT_WIDE_FLUSH_EVERY = 3
OPT_PDL = 4
OPT_CSR_FUSED = 5
OPT_WIDE_NO_PREFETCH = 6
OPT_DP_TIMEOUT_MS = 7
OPT_WIDE_MASK_HANDOVER = 8
OPT_WIDE_FWD_MASK = 9
OPT_CSR_STAGE_IDX = 10

GSO_MODES = {"binary_le": GSO_BINARY_LE, "sym_norm_lt": GSO_SYM_NORM_LT, "binary_lt": GSO_BINARY_LT}
ACTIVATIONS = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "leaky_relu": ACT_LEAKY_RELU}
PRECISIONS = {"fp32": PREC_FP32_3XTF32, "3xtf32": PREC_FP32_3XTF32, "tf32": PREC_TF32, "f16": PREC_F16}

_p, _i, _f, _d, _sz, _i64 = ct.c_void_p, ct.c_int, ct.c_float, ct.c_double, ct.c_size_t, ct.c_int64

# name -> (restype, argtypes); must list every symbol include/gfc.h declares
SIGNATURES = {
    "gfc_version": (_i, []),
    "gfc_last_error": (ct.c_char_p, []),
    "gfc_last_launch_count": (_i, []),
    "gfc_set_option": (_i, [_i, _i]),
    "gfc_last_path": (_i, []),
    "gfc_set_debug_clock_buffer": (_i, [_p, _sz]),
    "gfc_device_info": (_i, [ct.POINTER(_i)] * 4),
    "gfc_use_stats": (_i, [_p]),
    "gfc_use_mask": (_i, [_p, _sz]),
    "gfc_mask_filled": (_i, []),
    "gfc_filter_mask_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "gfc_gso_build": (_i, [_p, _i, _i, _d, _i, _p, _p, _p]),
    "gfc_filter_workspace_bytes": (_sz, [_i] * 7),
    "gfc_filter_path": (_i, [_i] * 7),
    "gfc_tile_plan_info": (_i, [_i] * 7 + [ct.POINTER(_i)]),
    "gfc_filter_fwd": (_i, [_p, _p, _p, _p, _p] + [_i] * 6 + [_i, _f, _i, _p, _sz, _p]),
    "gfc_filter_fwd_pos": (_i, [_p, _p, _d, _i, _p, _p, _p] + [_i] * 5 + [_i, _f, _i, _p, _sz, _p]),
    "gfc_filter_fwd_pos_nm": (_i, [_p, _p, _d, _i, _p, _p, _p] + [_i] * 5 + [_i, _f, _i, _p, _sz, _p]),
    "gfc_filter_bwd": (_i, [_p] * 8 + [_i] * 6 + [_i, _f, _i, _p, _sz, _p]),
    "gfc_filter_bwd_pos": (_i, [_p, _p, _d, _i] + [_p] * 6 + [_i] * 5 + [_i, _f, _i, _p, _sz, _p]),
    "gfc_dp_exchange_bytes": (_sz, [_i, _i]),
    "gfc_dp_signal_bytes": (_sz, [_i, _i]),
    "gfc_filter_bwd_pos_dp": (_i, [_p, _p, _d, _i] + [_p] * 5 + [_i] * 5 + [_i, _f, _i, _p, _sz, _p, _p, _i, _i, _f, _p]),
    "gfc_filter_bwd_dp": (_i, [_p] * 7 + [_i] * 6 + [_i, _f, _i, _p, _sz, _p, _p, _i, _i, _f, _p]),
    "gfc_dp_allreduce": (_i, [_p, _p, _i, _p, _p, _i, _i, _f, _p]),
    "gfc_dp_status": (_i, [_p, _i, ct.POINTER(_i), _p]),
    "gfc_csr_count": (_i, [_p, _i, _i, _d, _i, _p, _p]),
    "gfc_csr_scan": (_i, [_p, _i, _i, _p, _p]),
    "gfc_csr_fill": (_i, [_p, _i, _i, _d, _i, _p, _i64, _p, _p, _p]),
    "gfc_csr_build": (_i, [_p, _i, _i, _d, _i, _p, _i64, _p, _p, _p, _p]),
    "gfc_filter_csr_workspace_bytes": (_sz, [_i] * 6),
    "gfc_filter_csr_fwd": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p] + [_i] * 5 + [_i, _f, _i, _p, _sz, _p]),
    "gfc_filter_csr_bwd": (_i, [_p] * 7 + [_i64] + [_p] * 6 + [_i] * 5 + [_i, _f, _i, _p, _sz, _p]),
}


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libgfc.so not built at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gnn-formation-control_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = ct.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
if os.environ.get("GFC_PDL") is not None:   # A/B switches (profiling aids)
    lib.gfc_set_option(OPT_PDL, int(os.environ["GFC_PDL"]))
if os.environ.get("GFC_DISABLE_TCGEN05") is not None:
    lib.gfc_set_option(OPT_DISABLE_TCGEN05, int(os.environ["GFC_DISABLE_TCGEN05"]))
for _kv in filter(None, os.environ.get("GFC_SET_OPTIONS", "").split(",")):   # e.g. GFC_SET_OPTIONS="10=0,9=0": any gfc_set_option key
    lib.gfc_set_option(int(_kv.split("=")[0]), int(_kv.split("=")[1]))
if os.environ.get("GFC_WIDE_MASK_HANDOVER") is not None:
    lib.gfc_set_option(OPT_WIDE_MASK_HANDOVER, int(os.environ["GFC_WIDE_MASK_HANDOVER"]))


class GfcError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib.gfc_last_error()
        super().__init__("%s failed with status %d: %s" % (where, code, msg.decode() if msg else ""))


def check(rc, where):
    if rc != GFC_OK:
        raise GfcError(rc, where)


def ptr(t):
    """raw device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else ct.c_void_p(t.data_ptr())


def version():
    return lib.gfc_version()


def tile_plan(B, N, G, F, K, backward=False, from_positions=False):
    out = (_i * 12)()
    check(lib.gfc_tile_plan_info(B, N, G, F, K, int(backward), int(from_positions), out), "gfc_tile_plan_info")
    keys = ["ok", "graphs_per_tile", "rows", "rows_padded", "ntiles", "grid", "smem_bytes",
            "taps_in_smem", "dh_in_registers", "nb_dh", "nparts", "ldz"]
    return dict(zip(keys, list(out)))


def last_launch_count():
    return lib.gfc_last_launch_count()


def last_path():
    return lib.gfc_last_path()
