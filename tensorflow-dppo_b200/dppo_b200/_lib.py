"""ctypes loader and prototypes for libdppo_b200.so (include/dppo_b200.h)."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG, "lib", "libdppo_b200.so")

MAX_HIDDEN = 8
NUM_LOSS_SLOTS = 8
NCCL_UNIQUE_ID_BYTES = 128
PEER_HANDLE_BYTES = 128
GAE_REFERENCE, GAE_STANDARD = 0, 1
GEMM_FP32_SIMT, GEMM_TF32X3, GEMM_BF16, GEMM_FP16X2 = 0, 1, 2, 3
PD_DIAG_GAUSSIAN, PD_CATEGORICAL = 0, 1
VLOSS_PER_SAMPLE, VLOSS_LITERAL_PAIRWISE = 0, 1


class LibraryMissing(RuntimeError):
    pass


class DppoConfig(C.Structure):
    _fields_ = [
        ("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("n_hidden", C.c_int32),
        ("hidden", C.c_int32 * MAX_HIDDEN), ("max_rows", C.c_int32), ("device", C.c_int32),
        ("gemm_mode", C.c_int32), ("entcoeff", C.c_float), ("vcoeff", C.c_float), ("clip_param", C.c_float),
        ("pd_type", C.c_int32), ("value_loss_mode", C.c_int32),
    ]


_p = C.c_void_p
_PROTOS = {
    # name: (restype, argtypes)
    "dppo_last_error": (C.c_char_p, []),
    "dppo_version": (C.c_int, []),
    "dppo_param_count": (C.c_int64, [C.POINTER(DppoConfig)]),
    "dppo_param_offsets": (C.c_int, [C.POINTER(DppoConfig), C.POINTER(C.c_int64), C.c_int]),
    "dppo_grad_buffer_len": (C.c_int64, [C.POINTER(DppoConfig)]),
    "dppo_workspace_bytes": (C.c_size_t, [C.POINTER(DppoConfig)]),
    "dppo_create": (C.c_int, [C.POINTER(DppoConfig), _p, C.c_size_t, C.POINTER(_p)]),
    "dppo_destroy": (C.c_int, [_p]),
    "dppo_gae": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                           _p, _p, _p]),
    "dppo_forward": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, _p]),
    "dppo_loss_grad": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, C.c_int, C.c_float, _p, _p, _p]),
    "dppo_comm_unique_id": (C.c_int, [_p]),
    "dppo_comm_init": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "dppo_comm_destroy": (C.c_int, [_p]),
    "dppo_comm_peer_export": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "dppo_comm_peer_import": (C.c_int, [_p, _p]),
    "dppo_allreduce_avg": (C.c_int, [_p, _p, C.c_int64, _p]),
    "dppo_adam_apply": (C.c_int, [_p, _p, _p, _p, _p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float,
                                  C.c_float, _p]),
    "dppo_sync_old": (C.c_int, [_p, _p, _p, _p]),
    "dppo_act": (C.c_int, [_p, _p, _p, C.c_int, C.c_uint64, C.c_uint64, _p, _p, _p]),
    "dppo_update_host": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                   C.c_float, _p, _p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _p, _p, _p]),
    "dppo_update": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_float,
                              _p, _p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _p, _p]),
    "dppo_categorical": (C.c_int, [_p, C.c_int, _p, _p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _p, _p]),
    "dppo_update_epochs": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_float, _p, _p, C.c_int, C.c_float,
                                     C.c_float, C.c_float, C.c_float, _p, _p, _p]),
    "dppo_minibatch_perm": (C.c_int, [_p, C.c_int, C.c_uint64, C.c_uint64, _p, _p]),
    "dppo_diag_gaussian": (C.c_int, [_p, C.c_int, _p, _p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _p, _p]),
    "dppo_train_host": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, _p, _p, _p, C.c_int, C.c_float, _p, _p, C.c_int, C.c_float,
                                  C.c_float, C.c_float, C.c_float, _p, _p, _p]),
    "dppo_save_state": (C.c_int, [_p, C.c_char_p, _p, _p, _p, _p, C.c_int64, _p]),
    "dppo_load_state": (C.c_int, [_p, C.c_char_p, _p, _p, _p, _p, C.POINTER(C.c_int64), _p]),
    "dppo_profile_reset": (C.c_int, [_p, C.c_int]),
    "dppo_profile_read": (C.c_int, [_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int]),
    "dppo_dbg_gemm": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p, C.c_int, _p, _p]),
}
EXPORTED = tuple(_PROTOS)
PROFILE_KINDS = ("gae", "gemm_nn", "gemm_nt", "gemm_tn", "loss", "finalize", "misc", "adam", "allreduce", "tc_gemm")

_lib = None


def load_library(path: str = None):
    """dlopen libdppo_b200.so and attach prototypes.  Raises LibraryMissing -- never falls back."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("DPPO_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise LibraryMissing(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  dppo_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def last_error(lib) -> str:
    return lib.dppo_last_error().decode("utf-8", "replace")
