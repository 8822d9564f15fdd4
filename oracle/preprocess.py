"""CPU restatement of ``PreProcessor.fit / transform / wse_2_depth`` and ``compute_norths_rule`` (NumPy, FP64).

Test infrastructure only (see ``oracle/__init__.py``).  PINNED: ``tests/golden/preprocess_reference.npz`` holds the
outputs of the reference's own, unmodified ``PreProcessor`` (``gpras/preprocess.py:866-1162``) run in the build
container by ``tests/golden/make_golden_reference.py``; ``tests/test_oracle.py`` checks every function here against
them.

What the reference computes
---------------------------
``fit`` (``preprocess.py:947-1007``): classify cells as always dry / always flooded / transitional from the column
max / min depth against ``wet_threshold`` (``:1096-1132``); drop always-dry cells; centre by the column mean;
multiply by the cell weights; PCA (scikit-learn ``IncrementalPCA`` -- with fewer samples than ``5 * cells`` it is one
SVD of the re-centred matrix, right singular vectors sign-normalised so that the largest-magnitude entry of every
component is positive, ``explained_variance = s^2 / (n - 1)``); keep ``spatial_mode_count`` components (North's rule
when not given, ``:1323-1353``); mean and population std of the retained scores.
``transform`` (``:1009-1039``): ``((x[:, wet] - input_mean) * weights) @ eofs.T``, standardised.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def wse_to_depth(x, elevations):
    """``PreProcessor.wse_2_depth`` (``preprocess.py:1041-1045``)."""
    return np.maximum(np.asarray(x, np.float64) - elevations, 0.0)


def classify(max_depth, min_depth, wet_threshold):
    """``_classify_depths`` (``preprocess.py:1127-1132``): 0 = unset ('' in the reference), 1 = AD, 2 = TF, 3 = AF."""
    cls = np.zeros(max_depth.shape, np.int8)
    cls[max_depth < wet_threshold] = 1
    cls[max_depth > wet_threshold] = 2
    cls[min_depth > wet_threshold] = 3
    return cls


def norths_rule(eigenvalues, n_samples) -> int:
    """``compute_norths_rule`` (``preprocess.py:1323-1353``)."""
    ev = np.asarray(eigenvalues, np.float64)
    ev = ev[ev > 1.0]
    if ev.size == 0:
        return 0
    gap = np.abs(np.diff(ev))
    err = np.sqrt(2.0 / n_samples) * ev[:-1]
    hit = np.nonzero(gap <= err)[0]
    if hit.size == 0 or hit[0] == 0:
        return int(ev.size)
    return int(hit[0])


@dataclass
class Fitted:
    wet_class: np.ndarray    # (C,) int8, see classify()
    dry: np.ndarray          # (C,) bool
    input_mean: np.ndarray   # (C_wet,)
    weights: np.ndarray      # (C_wet,)
    eofs: np.ndarray         # (P, C_wet)
    eigenvalues: np.ndarray  # (min(N, C_wet),)
    x_mean: np.ndarray       # (P,)
    x_std: np.ndarray        # (P,)
    modes: int


def fit(x, elevations, weights, spatial_mode_count=None, wet_threshold=0.03, hydraulic_parameter="wse") -> Fitted:
    x = np.array(x, np.float64)
    if hydraulic_parameter == "depth":
        x = wse_to_depth(x, elevations)
        cls = classify(x.max(axis=0), x.min(axis=0), wet_threshold)
    elif hydraulic_parameter == "wse":
        cls = classify(x.max(axis=0) - elevations, x.min(axis=0) - elevations, wet_threshold)
    else:
        cls = np.full(x.shape[1], 2, np.int8)
    dry = cls == 1
    xw = x[:, ~dry]
    mu = xw.mean(axis=0)
    w = np.asarray(weights, np.float64)[~dry]
    a = (xw - mu) * w
    n = a.shape[0]
    a = a - a.mean(axis=0)  # IncrementalPCA re-centres its (single) batch
    _, s, vt = np.linalg.svd(a, full_matrices=False)
    sign = np.sign(vt[np.arange(vt.shape[0]), np.argmax(np.abs(vt), axis=1)])
    sign[sign == 0] = 1.0
    vt = vt * sign[:, None]
    ev = s * s / (n - 1)
    p = norths_rule(ev, n) if spatial_mode_count is None else int(spatial_mode_count)
    eofs = vt[:p]
    z = ((xw - mu) * w) @ eofs.T
    return Fitted(cls, dry, mu, w, eofs, ev, z.mean(axis=0), z.std(axis=0), p)


def transform(f: Fitted, x, elevations=None, hydraulic_parameter="wse"):
    x = np.asarray(x, np.float64)
    if hydraulic_parameter == "depth":
        x = wse_to_depth(x, elevations)
    z = ((x[:, ~f.dry] - f.input_mean) * f.weights) @ f.eofs.T
    return (z - f.x_mean) / f.x_std
