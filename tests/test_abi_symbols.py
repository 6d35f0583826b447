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


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call (error code + message, never a crash)."""
    from tfcausalimpact_b200 import _native as nt
    lib = nt.lib()
    L = nt.make_layout(1, 1, 10, 2, 0, 1, 1)
    bad = nt.CiLayout(0, 1, 10, 2, 0, 1, 1, 12)
    assert lib.ci_eval_workspace_bytes(ctypes.byref(bad)) == 0
    assert lib.ci_kalman_logp_grad(ctypes.byref(bad), None, None, None, None, None, ctypes.c_size_t(0), None) == -1
    assert b'non-positive' in lib.ci_last_error_string()
    assert lib.ci_kalman_logp_grad(ctypes.byref(L), None, None, None, None, None, ctypes.c_size_t(0), None) == -1
    assert b'NULL' in lib.ci_last_error_string()
    misaligned = nt.CiLayout(1, 1, 10, 2, 1, 1, 1, 13)
    assert lib.ci_pack_design(ctypes.byref(misaligned), None, None, None, None, None) == -1
    assert b'Tpad' in lib.ci_last_error_string()
    r_without_s = nt.CiLayout(1, 1, 10, 2, 0, 1, 3, 12)
    assert lib.ci_series_stats(ctypes.byref(r_without_s), None, ctypes.c_float(0.01), None, None) == -1
    huge_s = nt.make_layout(1, 1, 10, 2, 0, 400, 1)
    assert lib.ci_eval_workspace_bytes(ctypes.byref(huge_s)) > 0
    assert lib.ci_fstate_floats(ctypes.byref(nt.make_layout(1, 1, 10, 2, 0, 7, 1))) == 7 + 7 * 8
    assert lib.ci_fstate_floats(ctypes.byref(nt.make_layout(1, 1, 10, 2, 0, 52, 1))) == 53 + 2 * 53 * 53
    # second Seasonal component (custom `model=`): latent = level + slope + 4 + 52 effects, 2 drift scales
    two = nt.make_layout(1, 1, 10, 2, 3, 4, 1, flags=nt.FLAG_LLT | nt.FLAG_DENSE_REG, S2=52, r2=1)
    assert lib.ci_num_params_layout(ctypes.byref(two)) == 1 + 2 + 3 + 2 == nt.num_params(3, 4, two.flags, 52)
    assert lib.ci_fstate_floats(ctypes.byref(two)) == 58 + 2 * 58 * 58
    second_without_first = nt.make_layout(1, 1, 10, 2, 0, 1, 1, S2=4, r2=1)
    assert lib.ci_series_stats(ctypes.byref(second_without_first), None, ctypes.c_float(0.01), None, None) == -1
    assert b'second seasonal' in lib.ci_last_error_string()
