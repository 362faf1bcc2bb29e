"""fusion_b200: B200-native retrieval scoring + rank fusion behind the Python entry points of maastrichtlawtech/fusion.

Layers: ``csrc/`` (CUDA kernels + C ABI, include/fusion_b200.h) -> ``ops`` (tensor API) -> ``index`` (device index
containers) -> ``retrievers`` / ``utils`` (drop-in mirrors of the reference's signatures) -> ``sharding`` (multi-GPU).
"""
__version__ = "0.1.0"
