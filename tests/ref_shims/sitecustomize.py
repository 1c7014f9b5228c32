# numpy >= 2 removed np.Inf, which the reference's gaussian_constraints.py:65 still uses.
import numpy as _np
if not hasattr(_np, "Inf"):
    _np.Inf = _np.inf
