"""``src/utils``-side spelling of the fusion entry points (the reference keeps them in src/retrievers/hybrid.py)."""
from ..retrievers.hybrid import Aggregator  # noqa: F401

fuse = Aggregator.fuse
convert2dict = Aggregator.convert2dict
transform_scores = Aggregator.transform_scores
weight_scores = Aggregator.weight_scores
aggregate_scores = Aggregator.aggregate_scores
