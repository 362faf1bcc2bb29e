"""The driver's "does it build" check: __graft_entry__.build() must compile the library for sm_100a (nvcc cross-compiles
without a GPU), load it and find every C-ABI symbol at the ABI version the Python layer expects."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_entry_point_runs_on_cpu():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    entry = importlib.import_module("__graft_entry__")
    entry.build()
    from fusion_b200 import _lib, build
    assert build.is_current()
    assert _lib.load().fz_abi_version() == _lib.ABI_VERSION
    assert callable(entry.smoke)
