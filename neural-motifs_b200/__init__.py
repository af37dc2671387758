"""neural-motifs_b200 — B200-native hot path of rowanz/neural-motifs.

The directory name carries a hyphen (it mirrors the reference repo's name), so it is loaded by
path: `__graft_entry__.load_package()` / tests' conftest put this directory on sys.path, after
which the reference's own import paths work unchanged (`from lib.rel_model import RelModel`,
`from lib.fpn.roi_align.functions.roi_align import RoIAlignFunction`, `from config import ...`).
"""
import os
import sys

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
if PACKAGE_DIR not in sys.path:
    sys.path.insert(0, PACKAGE_DIR)

import motifs_cabi  # noqa: E402  (ctypes binding of csrc/libmotifs_b200.so)
