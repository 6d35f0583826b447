"""The C-ABI library loads and exports every symbol include/ci_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tfcausalimpact_b200 import build
    build.build()
    hdr = open(os.path.join(ROOT, 'include', 'ci_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b(ci_[a-z_0-9]+)\s*\(', hdr))
    assert len(names) >= 18
    lib = ctypes.CDLL(os.path.join(ROOT, 'tfcausalimpact_b200', 'libci_b200.so'))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.ci_version() >= 100
    assert lib.ci_num_params(50, 7) == 155 and lib.ci_num_params(0, 1) == 2 and lib.ci_num_params(1, 1) == 7
