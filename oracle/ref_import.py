"""Import the UNMODIFIED reference (build container only) — TEST INFRASTRUCTURE.

/root/reference is read-only and exists only in the build container (never on the GPU box); ``__graft_entry__.build()``
installs an unmodified copy into the git-ignored ``baseline/_ref`` (``pip install --target``), which does travel.  Used by
``oracle/gen_golden.py``, by CPU tests that are skipped when no reference tree is present, and by ``oracle/cpu_arm.py``
(bench.py's reference arm and cpu_baseline leg).  ``gr_bcm.py:4`` imports matplotlib (not installed, never used), so an
empty in-process stub is seeded into ``sys.modules`` before the import.  Importing the reference
flips torch's process-wide default dtype to float64 (covar.py:5); callers that care restore it.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# where the unmodified reference may live: an explicit override, the read-only mount of the build container, or the copy
# installed by __graft_entry__.build() into the git-ignored baseline/_ref (pip install --target; it travels to the GPU box)
_CANDIDATES = [os.environ.get("PYGPR_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]


def root():
    for c in _CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "PyGPR")):
            return c
    return None


REF_ROOT = root() or "/root/reference"


def available() -> bool:
    return root() is not None


def load():
    """Return the reference package module ``PyGPR`` (raises if the tree is absent)."""
    REF_ROOT = root()
    if REF_ROOT is None:
        raise ImportError("reference tree not present (looked in %s)" % [c for c in _CANDIDATES if c])
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
