"""ctypes binding of ``libgpras_b200.so`` (the C ABI declared in ``include/gpras_b200.h``).

There is deliberately no fallback: if the shared library is missing or no CUDA device is visible,
every compute call raises.  ``SYMBOLS`` lists every exported entry point with its signature; the
CPU test-suite checks that the library exports each of them.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libgpras_b200.so"
ABI_VERSION = 4  # == GPRAS_B200_ABI_VERSION in include/gpras_b200.h

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
vp = C.c_void_p

# name -> (restype, argtypes)
SYMBOLS = {
    "gpras_abi_version": (C.c_int, []),
    "gpras_last_error": (C.c_char_p, []),
    "gpras_device_count": (C.c_int, []),
    "gpras_gp_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpras_gp_destroy": (C.c_int, [vp]),
    "gpras_gp_set_stream": (C.c_int, [vp, vp]),
    "gpras_gp_set_data": (C.c_int, [vp, vp, vp, C.c_int]),
    "gpras_gp_lml_grad": (C.c_int, [vp, vp, vp, vp]),
    "gpras_gp_lml_grad_host": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "gpras_gp_lml_grad_enqueue": (C.c_int, [vp, vp, C.c_int]),
    "gpras_gp_lml_grad_fetch": (C.c_int, [vp, vp, vp]),
    "gpras_gp_condition": (C.c_int, [vp, vp]),
    "gpras_gp_predict": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int]),
    "gpras_gp_set_cell_map": (C.c_int, [vp, vp, vp, C.c_int]),
    "gpras_gp_predict_cells": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_long]),
    "gpras_gp_cell_pitch": (C.c_long, [vp]),
    "gpras_gp_get_matrix": (C.c_int, [vp, C.c_int, vp]),
    "gpras_gp_set_blocking_wait": (C.c_int, [vp, C.c_int]),
    "gpras_gp_last_launches": (C.c_int, [vp]),
    "gpras_gp_last_stage_ms": (C.c_int, [vp, vp]),
    "gpras_gp_set_stage_timing": (C.c_int, [vp, C.c_int]),
    "gpras_sgpr_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpras_sgpr_destroy": (C.c_int, [vp]),
    "gpras_sgpr_set_data": (C.c_int, [vp, vp, vp, C.c_int]),
    "gpras_sgpr_elbo_grad": (C.c_int, [vp, vp, vp, C.c_double, vp, vp, vp]),
    "gpras_sgpr_elbo_grad_enqueue": (C.c_int, [vp, vp, vp, C.c_double, C.c_int]),
    "gpras_sgpr_elbo_grad_fetch": (C.c_int, [vp, vp, vp, vp]),
    "gpras_sgpr_condition": (C.c_int, [vp, vp, vp, C.c_double]),
    "gpras_sgpr_predict": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "gpras_sgpr_last_launches": (C.c_int, [vp]),
    "gpras_sgpr_batch_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpras_sgpr_batch_destroy": (C.c_int, [vp]),
    "gpras_sgpr_batch_set_data": (C.c_int, [vp, vp, vp]),
    "gpras_sgpr_batch_elbo_grad": (C.c_int, [vp, vp, vp, C.c_double, vp, vp, vp, vp]),
    "gpras_sgpr_batch_adam": (
        C.c_int,
        [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, vp, vp, vp],
    ),
    "gpras_sgpr_batch_train": (
        C.c_int,
        [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, vp, vp, vp],
    ),
    "gpras_sgpr_batch_condition": (C.c_int, [vp, vp, vp, C.c_double, vp]),
    "gpras_sgpr_batch_predict": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "gpras_sgpr_batch_last_launches": (C.c_int, [vp]),
    "gpras_metrics_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_long]),
    "gpras_metrics_destroy": (C.c_int, [vp]),
    "gpras_metrics_set_elevations": (C.c_int, [vp, vp, vp]),
    "gpras_metrics_reset": (C.c_int, [vp, C.c_double]),
    "gpras_metrics_update": (C.c_int, [vp, vp, C.c_long, vp, C.c_long, vp, C.c_long, C.c_int, C.c_int]),
    "gpras_gp_predict_metrics": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, C.c_long, C.c_int, vp, vp]),
    "gpras_metrics_update_modes": (C.c_int, [vp, vp, vp, C.c_long, C.c_int, vp, C.c_long, vp, vp, C.c_long, C.c_int]),
    "gpras_metrics_finalize": (C.c_int, [vp, C.c_double, vp, vp, vp]),
    "gpras_metrics_timesteps": (C.c_long, [vp]),
    "gpras_metrics_last_launches": (C.c_int, [vp]),
    "gpras_metrics_fidelity": (C.c_int, [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp]),
    "gpras_pre_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_double]),
    "gpras_pre_destroy": (C.c_int, [vp]),
    "gpras_pre_fit": (C.c_int, [vp, vp, C.c_long, C.c_int, C.c_int, vp, vp, C.c_int, C.c_double, C.c_int]),
    "gpras_pre_set_modes": (C.c_int, [vp, C.c_int]),
    "gpras_pre_set_state": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]),
    "gpras_pre_get": (C.c_int, [vp, C.c_int, vp]),
    "gpras_pre_modes": (C.c_int, [vp]),
    "gpras_pre_eigen_count": (C.c_int, [vp]),
    "gpras_pre_iterations": (C.c_int, [vp]),
    "gpras_pre_transform": (C.c_int, [vp, vp, C.c_long, C.c_int, C.c_int, vp]),
    "gpras_pre_reverse": (C.c_int, [vp, vp, vp, C.c_int, vp, vp]),
    "gpras_pre_reverse_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_long]),
    "gpras_pre_cell_pitch": (C.c_long, [vp]),
    "gpras_pre_reverse_metrics": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_long, C.c_int]),
    "gpras_pre_trim": (C.c_int, [vp]),
    "gpras_pre_last_launches": (C.c_int, [vp]),
    "gpras_pre_last_stage_ms": (C.c_int, [vp, vp]),
    "gpras_dsyev128": (C.c_int, [vp, vp, vp, vp]),
    "gpras_kmeans_lloyd": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_double, vp, vp, vp]),
    "gpras_dgemm_tiles": (
        C.c_int,
        [vp, C.c_int, C.c_int, C.c_int, vp, C.c_long, vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double],
    ),
    "gpras_dpotrf": (C.c_int, [vp, vp, C.c_long, vp, C.c_long, C.c_int, vp, vp]),
    "gpras_dtrtri": (C.c_int, [vp, vp, C.c_long, vp, C.c_long, vp, C.c_long, C.c_int]),
    "gpras_dlauum": (C.c_int, [vp, vp, C.c_long, vp, C.c_long, C.c_int]),
}

KERNEL_IDS = {"RBF": 0, "Matern12": 1, "Matern32": 2, "Matern52": 3, "Exponential": 4}

_lib = None


class GprasError(RuntimeError):
    """Negative status from the C ABI (bad argument, CUDA failure, wrong state)."""


class NotPositiveDefiniteError(np.linalg.LinAlgError):
    """Positive status: the covariance matrix lost positive definiteness at the given pivot."""


def load() -> C.CDLL:
    """Load the in-tree shared library, binding every declared symbol.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise GprasError(
            f"{LIB_PATH} is missing: build it with `python -m gpras_b200.build` "
            "(gpras_b200 has no CPU fallback; the CUDA library is the only compute path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == 0:
        return
    msg = load().gpras_last_error().decode("utf-8", "replace")
    if status > 0:
        raise NotPositiveDefiniteError(f"{msg} (first failing pivot {status})")
    raise GprasError(f"gpras_b200 C ABI error {status}: {msg}")


def ptr(a) -> int:
    """Raw address of a C-contiguous float64 numpy array or a torch tensor."""
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
            raise ValueError("expected a C-contiguous float64 array")
        return a.ctypes.data
    return a.data_ptr()
