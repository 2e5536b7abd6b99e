"""smartqsm_b200 — B200-native (sm_100a) drop-in for SmartQSM's ``spconv-contraction`` skeletonisation stage.

Only the pieces that path needs live here: ``csrc/`` (CUDA kernels + the C ABI of
``libsqsm_b200.so``), ``_lib`` (ctypes binding), ``core_algorithms`` (mirror of the reference's
plug-in interface), ``preprocess`` (the cloud preparation in front of the path), ``sharding``
(tree-sharded multi-GPU driver) and ``synthetic`` (seeded inputs).
There is no CPU fallback: creating an engine without the built library or without an sm_100 GPU
raises.
"""
__version__ = "0.1.0"


def __getattr__(name):
    if name in ("SpconvBasedContraction", "CoreAlgorithmBase", "registry", "register_into", "load_checkpoint"):
        from . import core_algorithms
        return getattr(core_algorithms, name)
    if name in ("prepare_cloud", "PreparedCloud"):
        from . import preprocess
        return getattr(preprocess, name)
    if name == "Engine":
        from ._lib import Engine
        return Engine
    raise AttributeError(name)
