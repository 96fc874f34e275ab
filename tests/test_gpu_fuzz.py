"""Randomised parity sweep (tools/fuzz_parity.py): ragged sizes, 2..9 experts, aligned / unaligned communication sets, shared /
per-expert hyper-parameters, every quantity of the path against the CPU oracle at relative 1e-8."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed,form", [(11, "subst"), (12, "auto")])
def test_random_shapes_against_oracle(seed, form):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), str(seed), "16"], capture_output=True,
                         text=True, timeout=600, cwd=ROOT, env=dict(os.environ, PYGPR_PREDICT_FORM=form))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "fuzz_parity: ok" in res.stdout


def test_repeated_runs_are_bit_identical():
    """tools/determinism_check.py: fit + predict + gradient repeated (dataflow / two-stream Cholesky, column-owner solve, shared and
    general paths) reproduce bit for bit — the library has no atomically-ordered sums, so any difference would be a race."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "determinism_check.py"), "6"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "determinism_check: ok" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_random_spd_matrices_all_schedules():
    """tools/fuzz_potrf.py: in-place factorisation + inverse of random SPD batches (ragged orders, 1..12 matrices) under every
    schedule of the Cholesky against torch.linalg on the same GPU."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_potrf.py"), "3", "8"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "fuzz_potrf: ok" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
