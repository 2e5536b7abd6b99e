"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/sqsm_b200.h, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sqsm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sqsm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from smartqsm_b200 import _lib
    from smartqsm_b200.build import build_library
    build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sqsm_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "python binding and header disagree"
    assert lib.sqsm_abi_version() == 1


def test_struct_layouts_match_header():
    from smartqsm_b200 import _lib
    lib = _lib.load_library()                     # sizeof() as compiled into the library against the ctypes declarations
    for which, struct in enumerate((_lib.SqsmConfig, _lib.SqsmIterStats, _lib.SqsmPrepConfig, _lib.SqsmPrepStats)):
        assert lib.sqsm_sizeof(which) == ctypes.sizeof(struct), struct.__name__
    assert lib.sqsm_sizeof(99) == -1
    assert ctypes.sizeof(_lib.SqsmConfig) == 16 * 4
    assert ctypes.sizeof(_lib.SqsmIterStats) == 14 * 8 + 8 + 4 * 4 + 5 * 4 + 4   # 14 int64, double, 4 int32, 5 float, tail pad


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from smartqsm_b200._lib import Engine
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback|sm_100"):
        Engine()
    from smartqsm_b200.core_algorithms import SpconvBasedContraction
    with pytest.raises(RuntimeError, match="no CPU path"):
        SpconvBasedContraction(ckpt_path="x.npz", batch_size=16, max_iteration=10, device="cpu")


def test_plugin_interface_mirrors_reference():
    import inspect
    from smartqsm_b200.core_algorithms import CoreAlgorithmBase, SpconvBasedContraction, registry
    sig = inspect.signature(SpconvBasedContraction.__init__)
    for name in ("ckpt_path", "batch_size", "max_iteration", "device", "voxel_size", "cube_size", "buffer_size",
                 "verbose", "monitored_by_chamfer_distance"):
        assert name in sig.parameters
    assert sig.parameters["voxel_size"].default == 0.01 and sig.parameters["cube_size"].default == 4.0
    assert sig.parameters["buffer_size"].default == 0.4 and sig.parameters["device"].default == "cuda"
    assert issubclass(SpconvBasedContraction, CoreAlgorithmBase)
    assert registry["thinning"]["spconv_based_contraction"] is SpconvBasedContraction
    for m in ("input", "get_pipeline", "output"):
        assert callable(getattr(SpconvBasedContraction, m))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: importing every product module must not pull it in, and no product source
    mentions it (bench.py may only use it on the cpu_baseline / --impl reference legs)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); "
            "import smartqsm_b200, smartqsm_b200._lib, smartqsm_b200.core_algorithms, smartqsm_b200.preprocess, "
            "smartqsm_b200.sharding, smartqsm_b200.synthetic, smartqsm_b200.checkpoints, smartqsm_b200.build; "
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; "
            "assert not bad, bad") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "smartqsm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
