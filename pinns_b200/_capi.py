"""ctypes binding of libpinn_b200.so (the C ABI declared in include/pinn_b200.h).

There is no CPU fallback: if the shared library has not been built (see
``__graft_entry__.build`` / ``pinns_b200/csrc/Makefile``) importing this module
raises, and every compute entry point fails loudly without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

PINN_B200_ABI_VERSION = 1
PINN_MAX_LAYERS = 16

PDE_BURGERS, PDE_EULER = 0, 1
LOSS_V1_INF_L2, LOSS_V2_INF_ADMM, LOSS_V3_L1SQ, LOSS_V4_MSE, LOSS_V5_ADMM = 1, 2, 3, 4, 5
PATH_AUTO, PATH_GENERIC, PATH_FUSED, PATH_TENSOR = 0, 1, 2, 3
NSUMS = 8
SUM_DATA, SUM_RES, SUM_ABSF, SUM_MISFIT, SUM_F2 = 0, 1, 2, 3, 4

# PINN_B200_LIB: measurement knob (A/B runs of differently compiled kernels); the default is the in-tree build
LIB_PATH = os.environ.get("PINN_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpinn_b200.so")


class PinnConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_layers", C.c_int32),
        ("layers", C.c_int32 * PINN_MAX_LAYERS),
        ("pde", C.c_int32),
        ("loss", C.c_int32),
        ("lb", C.c_double * 2),
        ("ub", C.c_double * 2),
        ("lambda1", C.c_float),
        ("lambda2", C.c_float),
        ("rho", C.c_float),
        ("trainable_lambda", C.c_int32),
        ("device", C.c_int32),
        ("path", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


_H = C.c_void_p
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); mirrors include/pinn_b200.h one to one
PROTOTYPES = {
    "pinn_create": (C.c_int, [C.POINTER(PinnConfig), C.POINTER(_H)]),
    "pinn_destroy": (C.c_int, [_H]),
    "pinn_last_error": (C.c_char_p, [_H]),
    "pinn_set_stream": (C.c_int, [_H, C.c_void_p]),
    "pinn_synchronize": (C.c_int, [_H]),
    "pinn_num_params": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "pinn_packed_len": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "pinn_kernel_path": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "pinn_launch_count": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "pinn_set_params": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "pinn_get_params": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "pinn_set_lambda": (C.c_int, [_H, C.c_float, C.c_float]),
    "pinn_get_lambda": (C.c_int, [_H, _fp, _fp]),
    "pinn_set_data": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "pinn_set_collocation": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int]),
    "pinn_feed_collocation": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int64]),
    "pinn_comm_export": (C.c_int, [_H, C.c_void_p]),
    "pinn_comm_attach": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    "pinn_comm_detach": (C.c_int, [_H]),
    "pinn_comm_status": (C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pinn_sample_collocation": (C.c_int, [_H, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64]),
    "pinn_sample_lhs": (C.c_int, [_H, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int64]),
    "pinn_resampled_epochs": (C.c_int, [_H, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64]),
    "pinn_get_collocation": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "pinn_set_data_weight": (C.c_int, [_H, C.c_float]),
    "pinn_loss_grad_device": (C.c_int, [_H]),
    "pinn_packed_ptr": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "pinn_l1_pass1": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "pinn_loss_grad": (C.c_int, [_H, C.POINTER(C.c_double), C.c_void_p]),
    "pinn_loss_value": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "pinn_admm_misfit": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "pinn_adam_config": (C.c_int, [_H, C.c_float, C.c_float, C.c_float, C.c_float]),
    "pinn_adam_apply": (C.c_int, [_H]),
    "pinn_adam_steps": (C.c_int, [_H, C.c_int64]),
    "pinn_adam_reset": (C.c_int, [_H]),
    "pinn_predict": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]),
    "pinn_admm_init": (C.c_int, [_H]),
    "pinn_admm_update": (C.c_int, [_H, C.c_int]),
    "pinn_admm_adam_step": (C.c_int, [_H, C.c_int]),
    "pinn_admm_get_state": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int]),
    "pinn_admm_set_state": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int]),
    "pinn_kernel_timing": (C.c_int, [_H, C.c_int]),
    "pinn_kernel_time": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pinn_measure_fma_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C pinns_b200/csrc`).  pinns_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load_library()


class PinnError(RuntimeError):
    pass


def check(rc: int, handle=None, what: str = "") -> None:
    if rc != 0:
        msg = lib.pinn_last_error(handle)
        raise PinnError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
