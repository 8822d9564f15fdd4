"""Modes -> mesh-cells expansion of predicted mean and variance (NumPy, FP64).

Test infrastructure only (see ``oracle/__init__.py``).  Restates
``PreProcessor.reverse_transform`` / ``_linear_transform_for_var``
(``gpras/preprocess.py:1052-1094``) as used at ``production/analysis/pipeline.py:261``:

    wet cells:  mean_c = ((m * x_std + x_mean) @ eofs) / weights + input_mean
                var_c  = v @ (x_std[:, None] * eofs / weights) ** 2
    dry cells:  mean_c = elevation (``wse`` / ``velocity``) or 0 (``depth``);  var_c = 0
"""

from __future__ import annotations

import numpy as np


def reverse_transform(mean, var, eofs, x_mean, x_std, weights, input_mean, dry_indices, elevations, depth=False):
    mean = np.asarray(mean, np.float64)
    t = mean.shape[0]
    c = dry_indices.shape[0]
    wet = ~dry_indices
    out_m = np.empty((t, c))
    out_m[:, dry_indices] = 0.0 if depth else elevations[dry_indices]
    cells = (mean * x_std + x_mean) @ eofs
    if weights is not None:
        cells = cells / weights
    out_m[:, wet] = cells + input_mean
    if var is None:
        return out_m
    a = x_std[:, None] * eofs
    if weights is not None:
        a = a / weights[None, :]
    out_v = np.zeros((t, c))
    out_v[:, wet] = np.asarray(var, np.float64) @ (a * a)
    return out_m, out_v
