"""The C-ABI library loads and exports every symbol include/jvgpu.h declares; struct layouts match
between the header (compiled with gcc) and the ctypes mirrors.  No compute calls: runs without a GPU."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "jvgpu.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"^JV_API\s+[\w\s\*]+?\b(jv_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("jv_index_create", "jv_index_destroy", "jv_search_batch", "jv_exact_topk", "jv_pq_encode",
                 "jv_merge_topk", "jv_last_error", "jv_version"):
        assert must in syms


def test_library_exports_every_declared_symbol(jv):
    lib = jv.native.load()
    syms = declared_symbols()
    assert set(syms) == set(jv.native.SYMBOLS), "ctypes table and header disagree"
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.jv_version() == (0 << 16) | 1
    assert lib.jv_last_error() is not None


def test_struct_layouts_match_header(jv, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "jvgpu.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(jv_index_desc), sizeof(jv_search_params),"
        " sizeof(jv_query_stats), sizeof(jv_batch_timing), offsetof(jv_index_desc, adjacency), offsetof(jv_index_desc, pq_codes),"
        " offsetof(jv_search_params, accept_bits), offsetof(jv_index_desc, flags));"
        "printf(\"%zu %zu %d\\n\", offsetof(jv_index_desc, nvq_m), offsetof(jv_index_desc, nvq_global_mean), JV_INDEX_DESC_SIZE_V1);return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    n = jv.native
    assert [int(x) for x in out] == [
        C.sizeof(n.IndexDesc), C.sizeof(n.SearchParams), C.sizeof(n.QueryStats), C.sizeof(n.BatchTiming),
        n.IndexDesc.adjacency.offset, n.IndexDesc.pq_codes.offset, n.SearchParams.accept_bits.offset, n.IndexDesc.flags.offset,
        n.IndexDesc.nvq_m.offset, n.IndexDesc.nvq_global_mean.offset, n.IndexDesc.nvq_m.offset]  # V1 layout ends where nvq_m starts


def test_oracle_desc_mirror_matches(jv, oracle):
    assert C.sizeof(oracle.IndexDesc) == C.sizeof(jv.native.IndexDesc)
    assert [f[0] for f in oracle.IndexDesc._fields_] == [f[0] for f in jv.native.IndexDesc._fields_]


def test_no_cpu_fallback_without_gpu(jv):
    """On a box without a CUDA device every compute entry point must fail loudly (JV_ERR_CUDA)."""
    import numpy as np
    lib = jv.native.load()
    cnt = C.c_int32(0)
    st = lib.jv_device_count(C.addressof(cnt))
    if st == 0 and cnt.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(jv.native.JVectorNativeError) as ei:
        jv.pq_encode(np.zeros((4, 8), np.float32), 2, 4, np.zeros(32, np.float32))
    assert ei.value.status == jv.native.ERR_CUDA
    with pytest.raises(jv.native.JVectorNativeError):
        jv.GpuIndex(0, np.zeros((4, 8), np.float32), np.full((4, 2), -1, np.int32), 0)


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pkg = ROOT / "opensearch-jvector_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        text = f.read_text()
        assert "oracle" not in text.replace("oracle's", "").replace("oracle ", "").replace("the oracle", "") or \
            not re.search(r"(import|include|CDLL|dlopen)[^\n]*oracle", text), f"{f} references oracle/"
        assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f"{f} imports oracle"
        assert "libjvoracle" not in text, f"{f} loads the oracle library"
