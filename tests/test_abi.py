"""The C-ABI shared library loads without a GPU and exports every symbol include/fusion_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "fusion_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(fz_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from fusion_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.fz_abi_version() == _lib.ABI_VERSION == 2


def test_argument_errors_are_reported_without_a_gpu():
    from fusion_b200 import _lib
    lib = _lib.load()
    rc = lib.fz_merge_topk_f32(None, None, 1, 1, 1, 1, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.fz_last_error()
    rc = lib.fz_dense_topk(None, None, None, None, 1, 10, 100, 1, ctypes.c_float(0), 0, 64, 4, None, None, None, None, 0, None)
    assert rc == -1
    with pytest.raises(_lib.FusionB200Error):
        _lib.check(rc, "fz_dense_topk")


def test_postings_struct_matches_header_layout():
    from fusion_b200 import _lib
    # fz_postings_t: 10 pointers, int64 dense_stride, 6 x int32, int64 n_docs
    assert ctypes.sizeof(_lib.Postings) == 10 * 8 + 8 + 6 * 4 + 8
    assert _lib.Postings.n_docs.offset == 10 * 8 + 8 + 6 * 4 and _lib.Postings.dense_stride.offset == 80


def test_shard_sync_struct_matches_header_layout():
    from fusion_b200 import _lib
    # fz_shard_sync_t: hook, user, exchange pointers, int32 n_shards, int32 floor_rank, int64 sched_docs
    assert ctypes.sizeof(_lib.ShardSync) == 3 * 8 + 8 + 8
    assert _lib.ShardSync.n_shards.offset == 24 and _lib.ShardSync.floor_rank.offset == 28
    assert _lib.ShardSync.sched_docs.offset == 32


def test_header_is_plain_c_and_layouts_match_ctypes(tmp_path):
    """include/fusion_b200.h must be consumable from C (cgo / JNI / ctypes-style bindings): compile it with gcc -std=c99
    and compare the struct layouts the compiler sees with the ctypes mirrors."""
    import shutil
    import subprocess
    from fusion_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr_check.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fusion_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(fz_shard_sync_t), offsetof(fz_shard_sync_t, n_shards),\n'
                   '  offsetof(fz_shard_sync_t, floor_rank), offsetof(fz_shard_sync_t, sched_docs), sizeof(fz_postings_t),\n'
                   '  offsetof(fz_postings_t, n_docs), offsetof(fz_postings_t, dense_stride), sizeof(fz_splade_head_t),\n'
                   '  offsetof(fz_splade_head_t, head_dim), offsetof(fz_splade_head_t, flags), sizeof(fz_build_plan_t),\n'
                   '  offsetof(fz_build_plan_t, n_tiled), offsetof(fz_build_plan_t, n_coarse)); return 0; }\n')
    exe = tmp_path / "hdr_check"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(_lib.ShardSync), _lib.ShardSync.n_shards.offset, _lib.ShardSync.floor_rank.offset,
                   _lib.ShardSync.sched_docs.offset, ctypes.sizeof(_lib.Postings), _lib.Postings.n_docs.offset,
                   _lib.Postings.dense_stride.offset, ctypes.sizeof(_lib.SpladeHead), _lib.SpladeHead.head_dim.offset,
                   _lib.SpladeHead.flags.offset, ctypes.sizeof(_lib.BuildPlan), _lib.BuildPlan.n_tiled.offset,
                   _lib.BuildPlan.n_coarse.offset]


def test_a_plain_c_host_links_and_calls_the_library(tmp_path):
    """What a cgo / JNI shim does: a C99 program includes the header, links libfusion_b200.so and calls entry points - here
    the ones that need no GPU (ABI version, an argument error with its message, a workspace-size query)."""
    import shutil
    import subprocess
    from fusion_b200 import _lib, build
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lib_path = build.build()
    src = tmp_path / "host.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "fusion_b200.h"\n'
                   'int main(void) {\n'
                   '  fz_build_plan_t plan; memset(&plan, 0, sizeof plan);\n'
                   '  int rc = fz_build_postings_plan(NULL, NULL, 0, 0, 1024, 512, 1 << 20, NULL, NULL, &plan, NULL, 0, NULL);\n'
                   '  printf("%d %d %zu %s\\n", fz_abi_version(), rc, fz_build_postings_workspace_bytes(1000), fz_last_error());\n'
                   '  return 0; }\n')
    exe = tmp_path / "host"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", os.path.dirname(lib_path), "-lfusion_b200", "-Wl,-rpath," + os.path.dirname(lib_path)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split(None, 3)
    assert int(out[0]) == _lib.ABI_VERSION
    assert int(out[1]) == -1 and "fz_build_postings_plan" in out[3]
    assert int(out[2]) > 0
