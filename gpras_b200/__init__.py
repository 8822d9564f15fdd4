"""gpras_b200: B200-native (sm_100a) Gaussian-process-regression hot path of fema-ffrd/gpras.

``gpras_b200.gpr`` mirrors ``gpras.gpr`` (``GPRAS``, ``KERNEL_FACTORY``, ``OPTIMIZERS`` ...); the numerics run in
hand-written CUDA behind ``libgpras_b200.so`` (C ABI in ``include/gpras_b200.h``).  No CPU fallback exists.
"""

from .gpr import GPRAS, KERNEL_FACTORY, OPTIMIZERS, InductionInitializerType, KernelType, OptimizerType  # noqa: F401

__all__ = ["GPRAS", "KERNEL_FACTORY", "OPTIMIZERS", "KernelType", "OptimizerType", "InductionInitializerType"]
