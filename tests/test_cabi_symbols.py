"""The C-ABI library builds, loads, and exports every symbol include/pygpr_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pygpr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:gpr|grbcm)_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from pygpr_b200.csrc import build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 14, names
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/pygpr_b200.h is not exported" % n


def test_ctypes_signatures_cover_header():
    from pygpr_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.gpr_abi_version() == _lib.ABI_VERSION


def test_bad_arguments_are_status_codes_not_crashes():
    """Argument validation happens before any CUDA call, so it is testable without a device."""
    from pygpr_b200 import _lib
    lib = _lib.load()
    rc = lib.gpr_se_scale(None, 0, None, 0, 1, 1, 1, 128, 4, None, None, None)
    assert rc == -1 and b"bad argument" in lib.gpr_last_error()
    rc = lib.gpr_potrf_batched(None, 0, None, None, None, 0, None, 1, 100, 100, 0, None, None, None, None, None, None, None, None)
    assert rc == -1


def test_no_oracle_or_cpu_fallback_in_product():
    """The product package must not import the oracle (test infrastructure) or carry a CPU compute path."""
    pkg = os.path.join(ROOT, "pygpr_b200")
    lapack_call = re.compile(r"\b(torch|tc)\.(linalg\.)?(cholesky\w*|solve_triangular|inv|solve|lu\w*)\s*\(")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), fn
            code = "\n".join(line.split("#")[0] for line in src.splitlines())
            code = re.sub(r'"""(.|\n)*?"""', "", code)
            assert not lapack_call.search(code), (fn, lapack_call.search(code).group(0))


def test_engine_refuses_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pygpr_b200 import engine
    with pytest.raises(RuntimeError):
        engine.device()


def test_comm_entry_points_without_a_gpu():
    """gpr_comm_*: exported, argument errors are reported before NCCL is touched (no compute, no GPU)."""
    import pytest
    from pygpr_b200 import _lib
    lib = _lib.load()
    assert lib.gpr_comm_unique_id_bytes() == 128
    with pytest.raises(_lib.GprError):
        _lib.call("gpr_comm_unique_id", None)
    with pytest.raises(_lib.GprError):
        _lib.call("gpr_comm_init", None, 2, 0, None)
    with pytest.raises(_lib.GprError):
        _lib.call("gpr_comm_allreduce_f64", None, None, 0, None)
    _lib.call("gpr_comm_destroy", None)           # destroying nothing is not an error
