"""Oracle (TEST INFRASTRUCTURE): the percentile-based score distribution of ``src/retrievers/hybrid.py:391-398``, with the
same pandas calls (drop the zeros and every occurrence of the two smallest distinct scores, then
``Series.quantile(np.linspace(0, 1, N + 1))``).  Pinned by ``oracle/make_golden.py::golden_distribution``."""
from __future__ import annotations

import numpy as np
import pandas as pd


def percentile_distribution(scores: np.ndarray, n_points: int) -> np.ndarray:
    s = pd.Series(np.asarray(scores, dtype=np.float64))
    kept = s[(s != 0.0) & (~s.isin(s.drop_duplicates().nsmallest(2)))]
    return kept.quantile(np.linspace(0, 1, n_points + 1)).to_numpy()
