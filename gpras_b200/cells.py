"""Modes -> mesh-cells expansion folded into one dense map (the step after ``GPRAS.predict`` in
``production/analysis/pipeline.py:260-261``).

``PreProcessor.reverse_transform`` (``gpras/preprocess.py:1052-1094``) computes, on wet cells,
``((m * x_std + x_mean) @ eofs) / weights + input_mean`` for the mean and ``v @ (diag(x_std) eofs / weights)^2``
for the variance, then scatters into the full cell vector (dry cells = elevation or 0, variance 0).  Folding the
scales and the scatter into ``e_mean`` (P x C, zero columns on dry cells) and ``bias`` (C) turns both into plain
GEMMs over all C cells, which the device runs on the DMMA engine with the bias fused into the epilogue:

    cell_mean = mode_mean @ e_mean + bias          cell_var = mode_var @ (e_mean ** 2)
"""

from __future__ import annotations

import numpy as np


def fold_cell_map(eofs, x_mean, x_std, weights, input_mean, dry_indices, elevations, depth: bool = False):
    """Return (e_mean (P, C), bias (C,)) equivalent to the reference's reverse transform."""
    eofs = np.asarray(eofs, np.float64)
    dry = np.asarray(dry_indices, bool)
    wet = ~dry
    p, c = eofs.shape[0], dry.shape[0]
    w = np.ones(eofs.shape[1]) if weights is None else np.asarray(weights, np.float64)
    e_wet = (np.asarray(x_std, np.float64)[:, None] * eofs) / w[None, :]
    e_mean = np.zeros((p, c))
    e_mean[:, wet] = e_wet
    bias = np.zeros(c)
    bias[wet] = (np.asarray(x_mean, np.float64) @ eofs) / w + np.asarray(input_mean, np.float64)
    if not depth:
        bias[dry] = np.asarray(elevations, np.float64)[dry]
    return e_mean, bias
