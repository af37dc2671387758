"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares, and the
ctypes table in motifs_cabi.py covers the same set (no compute calls — no GPU here)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^[A-Za-z_][\w\s\*]*?\b([A-Za-z_]\w*)\s*\(", src, flags=re.M):
            names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    import motifs_cabi
    assert os.path.exists(motifs_cabi.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(motifs_cabi.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 20
    missing = [n for n in decl if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    import motifs_cabi
    decl = declared_symbols()
    assert set(motifs_cabi.SIGNATURES) == decl, set(motifs_cabi.SIGNATURES) ^ decl
    lib = motifs_cabi.load()
    assert lib.mb200_abi_version() == 1
    assert lib.mb200_compiled_arch() == 100


def test_operators_refuse_cpu_tensors():
    import pytest
    import torch
    import motifs_cabi
    from lib.fpn.roi_align.functions.roi_align import RoIAlignFunction
    from lib.fpn.nms.functions.nms import apply_nms
    with pytest.raises(motifs_cabi.MotifsB200Error):
        RoIAlignFunction(7, 7, 1 / 16)(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5))
    with pytest.raises(motifs_cabi.MotifsB200Error):
        apply_nms(torch.zeros(3), torch.zeros(3, 4))
