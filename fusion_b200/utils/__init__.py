"""Mirrors of the ``src/utils`` entry points on the hot path (north-star spelling: "src/utils fusion entry points")."""
from ..retrievers.hybrid import Aggregator, Ranker  # noqa: F401
