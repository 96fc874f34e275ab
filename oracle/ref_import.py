"""Import the UNMODIFIED reference (build container only) — TEST INFRASTRUCTURE.

/root/reference is read-only and exists only in the build container (never on the GPU box), so
this module is used solely by ``oracle/gen_golden.py`` and by CPU tests that are skipped when the
reference tree is absent.  ``gr_bcm.py:4`` imports matplotlib (not installed, never used), so an
empty in-process stub is seeded into ``sys.modules`` before the import.  Importing the reference
flips torch's process-wide default dtype to float64 (covar.py:5); callers that care restore it.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("PYGPR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "PyGPR"))


def load():
    """Return the reference package module ``PyGPR`` (raises if the tree is absent)."""
    if not available():
        raise ImportError("reference tree not present at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import PyGPR  # noqa: E402
    return PyGPR
