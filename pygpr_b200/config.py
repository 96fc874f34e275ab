"""Developer A/B knobs of the host layer and of the library's schedules (not part of the reference's API).

Every knob takes its initial value from the environment ONCE, at import; afterwards it is a plain attribute
(``options.share = False``) or, for the two schedule overrides that live in the library, a setter call.

    PYGPR_GRBCM_SHARE=0        options.share        treat GRBCM experts as unrelated matrices (no shared leading block)
    PYGPR_GRBCM_FUSED_FIT=0    options.fused_fit    fit the communication expert in its own launch
    PYGPR_PREDICT_CHUNK=n      options.predict_chunk  test points per prediction batch of GRBCM.predict
    PYGPR_PREDICT_FORM=...     options.predict_form   auto | subst | gemm  (engine.predict)
    PYGPR_EMULATE_FP64=s       options.emulate_slices  GRBCM predictions through the emulated-FP64 GEMM with s digit planes (0 = off)
    PYGPR_EMULATE_CHUNK=e      options.emulate_chunk   experts per batched emulated product
    PYGPR_EMULATE_FP64_GRAD=s  options.emulate_grad_slices  K^-1 of the NLML gradient through the emulated-FP64 GEMM (0 = off)
    PYGPR_EMULATE_FP64_INVERSE=s  options.emulate_inverse_slices  L^-T assembled from segment inverses by emulated-FP64 products (0 = off)
    PYGPR_EMULATE_FP64_FIT=s   options.emulate_fit_slices   blocked Cholesky with emulated-FP64 panel solves / trailing updates (0 = off)
    PYGPR_POTRF_MODE=...       set_potrf_mode()     auto | columns | columns1 | dataflow
    PYGPR_TRSV_MODE=...        set_trsv_mode()      auto | single
"""
from __future__ import annotations

import os
from dataclasses import dataclass

from . import _lib


@dataclass
class Options:
    share: bool = os.environ.get("PYGPR_GRBCM_SHARE", "1") != "0"
    fused_fit: bool = os.environ.get("PYGPR_GRBCM_FUSED_FIT", "1") != "0"
    predict_chunk: int = int(os.environ.get("PYGPR_PREDICT_CHUNK", "16384"))
    predict_form: str = os.environ.get("PYGPR_PREDICT_FORM", "auto")
    # the data-path all-reduces: "torch" = torch.distributed (NCCL), "cabi" = the library's own NCCL wrapper (gpr_comm_*)
    comm: str = os.environ.get("PYGPR_COMM", "torch")
    # opt-in: the local experts' prediction products on the INT8 tensor cores with this many digit planes per operand
    # (0 = off, the FP64 DMMA path; 6 = FP64 grade, ~1e-11 on V; 3-4 = the FP32-grade mode).  engine.predict_shared_emulated
    emulate_slices: int = int(os.environ.get("PYGPR_EMULATE_FP64", "0"))
    emulate_chunk: int = int(os.environ.get("PYGPR_EMULATE_CHUNK", "8"))
    # below this (padded) order the emulated prediction and gradient stay off even when opted in: at n = 1000 its staging costs more than the
    # product it replaces (config 1: 0.38 vs 0.28 ms); from 2048 up it wins.  Tests set 0 to exercise it on small shapes.
    emulate_min_order: int = int(os.environ.get("PYGPR_EMULATE_MIN_ORDER", "2048"))
    # opt-in: K^-1 = L^-T L^-1 of the NLML gradient on the INT8 tensor cores with this many digit planes (0 = off; 7 keeps the
    # gradient at ~1e-10 of the FP64 path).  engine.lml_grad_emulated
    emulate_grad_slices: int = int(os.environ.get("PYGPR_EMULATE_FP64_GRAD", "0"))
    # opt-in: assemble L^-T (needed by the emulated prediction and gradient) from FP64 inverses of diagonal segments and
    # emulated-FP64 products between them, with this many digit planes (0 = the plain FP64 blocked trtri).  engine._inverse_factors
    emulate_inverse_slices: int = int(os.environ.get("PYGPR_EMULATE_FP64_INVERSE", "0"))
    # opt-in: the factorisation itself as a right-looking blocked Cholesky whose panel solves and trailing updates are emulated-FP64
    # products (diagonal blocks of `emulate_fit_panel` rows stay FP64; 0 = 2048 for up to 8 matrices, 1024 beyond), this many digit
    # planes (0 = the FP64 DMMA Cholesky)
    emulate_fit_slices: int = int(os.environ.get("PYGPR_EMULATE_FP64_FIT", "0"))
    emulate_fit_panel: int = int(os.environ.get("PYGPR_EMULATE_FIT_PANEL", "0"))
    emulate_fit_chunk: int = int(os.environ.get("PYGPR_EMULATE_FIT_CHUNK", "32"))
    # the FP64 work on the next diagonal block runs on a high-priority side stream under the current panel's trailing update (0 = in line)
    emulate_fit_lookahead: bool = os.environ.get("PYGPR_EMULATE_FIT_LOOKAHEAD", "1") != "0"


options = Options()


def set_potrf_mode(mode: str) -> None:
    _lib.call("gpr_set_option", b"potrf_mode", mode.encode())


def set_trsv_mode(mode: str) -> None:
    _lib.call("gpr_set_option", b"trsv_mode", mode.encode())


def set_ozaki_bk(bk: int) -> None:
    """k-block of the emulated-FP64 GEMM: 0 = default (64 B swizzled rows), 32 / 64 = forced (the 32 B variant is an A/B knob)."""
    _lib.call("gpr_set_option", b"ozaki_bk", str(int(bk)).encode())
