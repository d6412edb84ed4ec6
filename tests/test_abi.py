"""CPU-side checks of the drop-in boundary: libtsg.so loads and exports every symbol that
include/tsg.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tsg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "two-stage-gnn_b200", "libtsg.so"))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libtsg.so does not export {n}"
    lib.tsg_abi_version.restype = ctypes.c_int
    assert lib.tsg_abi_version() == 1


def test_python_prototypes_cover_header():
    from tsg import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_workspace_queries_are_pure_host():
    from tsg import _lib
    assert _lib.lib.tsg_csr_build_workspace_bytes(1000, 100) > 0
    assert _lib.lib.tsg_topk_workspace_bytes(1000, 10) > 0
    assert _lib.lib.tsg_filter_adj_workspace_bytes(5000) > 0
    assert _lib.lib.tsg_triplet_workspace_bytes(10, 30, 32) > 0


def test_no_cpu_fallback():
    import pytest
    import torch
    from tsg import ops
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(RuntimeError):
        ops.build_csr(ops.EdgeList.from_edge_index(ei), 2)
