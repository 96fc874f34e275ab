"""Host-side driver of the CUDA hot path: thin Python over the C ABI (include/pygpr_b200.h).

Everything here works on contiguous float64 CUDA tensors with an explicit leading expert
dimension ``B``; the reference-shaped API (covar.py / gpr.py / loss.py / gr_bcm.py) sits on top.
Torch is used for device memory and streams only; all arithmetic of the path runs in the
library's sm_100a kernels.  No CPU fallback exists: a missing library or device raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import TILE

F64 = torch.float64
JITTER = 1e-7  # gpr.py:68, loss.py:38,63,96
INFO_STALLED = -2 ** 31  # info[] sentinel of an aborted launch (GPR_INFO_STALLED in the library)

# number of library kernels launched so far (bench.py reports the delta over its timed region)
LAUNCHES = 0


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pygpr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _dev(t: torch.Tensor) -> torch.Tensor:
    """float64, contiguous, on the current CUDA device."""
    return t.detach().to(device=device(), dtype=F64).contiguous()


def round_up(n: int, mult: int) -> int:
    return (n + mult - 1) // mult * mult


def scale_inputs(x: torch.Tensor, hp: torch.Tensor, B: int, npad: Optional[int] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """a = x*ls zero-padded to [B, npad, dpad]; nrm = |a|^2 [B, npad]   (covar.py:137-140, 106).

    x: [B, n, d] or [n, d] / [1, n, d] (shared by all experts); hp: [B, P] with ls = hp[:, 1:1+d].
    """
    x = x.reshape(-1, x.shape[-2], x.shape[-1])
    n, d = x.shape[-2], x.shape[-1]
    assert x.shape[0] in (1, B)
    npad = round_up(n, TILE) if npad is None else npad
    dpad = round_up(d, 4)
    a = torch.empty(B, npad, dpad, dtype=F64, device=x.device)
    nrm = torch.empty(B, npad, dtype=F64, device=x.device)
    x_bstride = 0 if x.shape[0] == 1 else n * d
    hp_stride = 0 if hp.shape[0] == 1 else hp.shape[-1]
    _lib.call("gpr_se_scale", _p(x), x_bstride, _p(hp), hp_stride, B, n, d, npad, dpad, _p(a), _p(nrm), _stream())
    _count(1)
    return a, nrm


def se_kernel_matrix(hp: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Dense K_se [B, n, n] (no noise, no jitter) — Squared_exponential.kernel, covar.py:129-150."""
    B, n = x.shape[0], x.shape[1]
    a, nrm = scale_inputs(x, hp, B)
    K = torch.empty(B, n, n, dtype=F64, device=x.device)
    _lib.call("gpr_se_build", _p(a), _p(nrm), _p(hp), hp.shape[-1], None, B, n, a.shape[1], a.shape[2],
              _p(K), n, 0, 0, _stream())
    _count(1)
    return K


def se_cross_matrix(hp: torch.Tensor, x: torch.Tensor, xp: torch.Tensor) -> torch.Tensor:
    """Dense K* [B, m, n] — cross branch of Squared_exponential.kernel, covar.py:151-160.

    Exactly one of x / xp may carry fewer batch entries than hp (broadcast), as in the reference.
    """
    B = hp.shape[0]
    n, m = x.shape[-2], xp.shape[-2]
    a, nrm = scale_inputs(x, hp, B)
    as_, nrms = scale_inputs(xp, hp, B)
    Ks = torch.empty(B, m, n, dtype=F64, device=hp.device)
    _lib.call("gpr_se_cross", _p(as_), _p(nrms), as_.shape[1], _p(a), _p(nrm), a.shape[1], _p(hp), hp.shape[-1],
              B, m, n, a.shape[2], _p(Ks), _stream())
    _count(1)
    return Ks


def se_kernel_grad(hp: torch.Tensor, x: torch.Tensor, Kse: torch.Tensor) -> torch.Tensor:
    """dK [B, d+1, n, n] for Squared_exponential.kernel_and_grad (covar.py:184-199); API parity only."""
    B, n, d = x.shape
    dK = torch.empty(B, d + 1, n, n, dtype=F64, device=x.device)
    _lib.call("gpr_se_kernel_grad", _p(Kse), _p(x), _p(hp), hp.shape[-1], B, n, d, _p(dK), _stream())
    return dK


@dataclass
class FitState:
    """Device-resident result of a batched fit (the reference's krnchd / wt, gpr.py:57-59, plus extras)."""
    B: int
    n: int
    npad: int
    d: int
    dpad: int
    has_noise: bool
    hp: torch.Tensor          # [B, P]
    a: torch.Tensor           # [B, npad, dpad] scaled inputs
    nrm: torch.Tensor         # [B, npad]
    L: torch.Tensor           # [B, npad, npad] lower factor (blocks above the block diagonal undefined)
    invdiag: torch.Tensor     # [B, npad/128, 128, 128]
    z: torch.Tensor           # [B, npad]  L^-1 y
    alpha: Optional[torch.Tensor]  # [B, npad]  K^-1 y
    logdet_half: torch.Tensor  # [B]
    zz: torch.Tensor          # [B]
    info: torch.Tensor        # [B] int32
    prior: torch.Tensor       # [B] sig^2 + sig_n^2
    shared_from: Optional["FitState"] = None   # communication expert whose factor is the leading block of every expert here
    Linv: Optional[torch.Tensor] = None        # [B, npad, npad] L^-1 (tril_inverse), cached for the GEMM-form prediction
    Kinv: Optional[torch.Tensor] = None        # [B, n, n] K^-1 (kinv), cached for the shared-block gradient
    shared_upper: bool = False                 # the copied leading block also carries the communication expert's L_g^-T
    # state of the opt-in emulated forms (ensure_inverse_factors / _emulated_experts)
    upper_is_inverse: bool = False             # the strict block-upper part of L holds L^-T (in place)
    plane_cache: Optional[dict] = None         # digit planes of the inverse rows per expert chunk, kept while HBM allows
    sig2_max: Optional[float] = None           # max sig^2 over the experts: the bound of K*'s digit split, read back once

    def nlml(self) -> torch.Tensor:
        """NLML = 1/2 y^T alpha + sum log L_ii + n/2 log 2 pi   (loss.py:49-51); y^T alpha = z.z"""
        return 0.5 * self.zz + self.logdet_half + 0.5 * self.n * math.log(2.0 * math.pi)

    def record_stream(self, stream) -> None:
        """The state was produced on another stream: tell the caching allocator that `stream` uses its buffers too."""
        for t in (self.hp, self.a, self.nrm, self.L, self.invdiag, self.z, self.alpha, self.logdet_half, self.zz, self.info, self.prior):
            if t is not None:
                t.record_stream(stream)

    def check(self) -> None:
        """Mirror tc.cholesky's RuntimeError on a non-PD matrix (gpr.py:69)."""
        bad = torch.nonzero(self.info)
        if bad.numel():
            b = int(bad[0, 0])
            if int(self.info[b]) == INFO_STALLED:            # the hang guard of the dataflow launch fired: not a numerical verdict
                raise _lib.GprError("factorisation aborted: a tile task waited longer than the library's time limit for its "
                                    "dependencies (a scheduling stall on a wedged or heavily shared GPU); the matrices were NOT "
                                    "found to be non-positive-definite — retry")
            raise torch.linalg.LinAlgError(
                "cholesky: expert %d: the leading minor of order %d is not positive-definite" % (b, int(self.info[b])))


def split_hp(hp: torch.Tensor, d: int, has_noise: bool, jitter: float = JITTER):
    """diag_add = sig_n^2 + jitter (covar.py:238 + gpr.py:68), prior = sig^2 + sig_n^2 (gpr.py:98)."""
    sig2 = hp[:, 0] * hp[:, 0]
    if has_noise:
        sn2 = hp[:, d + 1] * hp[:, d + 1]
        return (sn2 + jitter).contiguous(), (sig2 + sn2).contiguous()
    return torch.full_like(sig2, jitter), sig2.contiguous()


def fused_dim_ok(d: int) -> bool:
    """Whether the on-chip kernel-tile generation fits: the two [128, dpad] feature tiles must sit behind the 128x128 tile in
    the 227 KB of shared memory (factorisation and gradient epilogue are the tightest: d <= 44).  Larger d takes the dense path."""
    dpad = round_up(d, 4)
    sld = dpad + 4 if dpad % 8 == 0 else dpad
    feat = (2 * TILE * sld + 2 * TILE) * 8
    return TILE * 132 * 8 + feat <= 226 * 1024


def can_share(shared: Optional[FitState], n: int, d: int, has_noise: bool) -> bool:
    """Shape conditions for the shared leading block (the caller vouches that rows 0:ng of every expert are the
    communication set and that all hyper-parameter rows equal the communication expert's)."""
    return (shared is not None and shared.B == 1 and shared.alpha is not None and shared.n >= TILE and shared.n < n
            and shared.d == d and shared.has_noise == has_noise)


def shared_rows(shared: FitState) -> int:
    """Rows of the communication expert that are shared with the local experts: its whole 128-row blocks."""
    return (shared.n // TILE) * TILE


def fit(x: torch.Tensor, y: torch.Tensor, hp: torch.Tensor, has_noise: bool, want_alpha: bool = True,
        jitter: float = JITTER, L_out: Optional[torch.Tensor] = None, ws_tag: str = "",
        shared: Optional[FitState] = None) -> FitState:
    """Batched Exact_GP.update (gpr.py:65-74) for cov = Compose([SE, WN]) or SE alone.

    K is never materialised: tiles are generated inside the factorisation (gen=1).
    x:[B,n,d], y:[B,n], hp:[B,P] — CUDA float64 contiguous.

    shared: the fitted communication expert when every expert's first ``shared.n`` rows are its training set and every
    hyper-parameter row equals its own (GRBCM with one shared vector): its factor is then the leading block of every
    expert's factor and is reused instead of recomputed (gpr_potrf_batched_shared).
    """
    B, n, d = x.shape
    assert hp.shape == (B, d + 1 + int(has_noise)), (hp.shape, B, d, has_noise)
    from .config import options
    if options.emulate_fit_slices > 0 and can_fit_emulated(n, shared) and fit_emulated_pays(B, n):
        return fit_emulated(x, y, hp, has_noise, options.emulate_fit_slices, want_alpha, jitter, L_out, shared,
                            options.emulate_fit_panel, options.emulate_fit_chunk)
    npad = round_up(n, TILE)
    nt = npad // TILE
    dev = x.device
    a, nrm = scale_inputs(x, hp, B, npad)
    dpad = a.shape[2]
    diag_add, prior = split_hp(hp, d, has_noise, jitter)
    L = torch.empty(B, npad, npad, dtype=F64, device=dev) if L_out is None else L_out
    invdiag = torch.empty(B, nt, TILE, TILE, dtype=F64, device=dev)
    ypad = torch.zeros(B, npad, dtype=F64, device=dev)
    ypad[:, :n] = y
    z = torch.empty(B, npad, dtype=F64, device=dev)
    logdet_half = torch.empty(B, dtype=F64, device=dev)
    zz = torch.empty(B, dtype=F64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    work = _work("potrf" + ws_tag, int(_lib.load().gpr_potrf_work_bytes(B, npad)), dev)   # ws_tag: a fit on another stream
    if shared is not None:
        assert can_share(shared, n, d, has_noise)
        ng, gt = shared_rows(shared), shared.n // TILE
        z[:, :ng] = shared.z[0, :ng]                      # the forward solve, log-determinant and y^T K^-1 y continue
        if ng == shared.n:                                # from the communication expert's values (its first ng rows)
            logdet_half[:] = shared.logdet_half[0]
            zz[:] = shared.zz[0]
        else:
            part = torch.empty(2, dtype=F64, device=dev)
            _lib.call("gpr_factor_partials", _p(shared.L), shared.npad, _p(shared.z), ng, _p(part), _stream())
            logdet_half[:] = part[0]
            zz[:] = part[1]
        info[:] = torch.where((shared.info[0] > 0) & (shared.info[0] <= ng), shared.info[0], torch.zeros_like(shared.info[0]))
        _lib.call("gpr_potrf_batched_shared", _p(L), _p(a), _p(nrm), _p(hp), hp.shape[-1], _p(diag_add), B, n, npad, dpad,
                  _p(invdiag), _p(ypad), _p(z), _p(logdet_half), _p(zz), _p(info), _p(work), _stream(), ng,
                  _p(shared.L), shared.npad, _p(shared.invdiag))
        # the leading block itself, for the kernels that walk the whole factor (backward solve, gradient, full covariance);
        # if L_g^-T was already formed in the communication factor's block-upper part (kinv / tril_inverse) it comes along
        shared_upper = shared.Kinv is not None or shared.Linv is not None
        _lib.call("gpr_copy_leading_block", _p(L), B, npad, _p(shared.L), shared.npad, ng, int(not shared_upper), _stream())
        invdiag[:, :gt] = shared.invdiag[0, :gt]
        _count(2)
    else:
        _lib.call("gpr_potrf_batched", _p(L), 1, _p(a), _p(nrm), _p(hp), hp.shape[-1], _p(diag_add), B, n, npad, dpad,
                  _p(invdiag), _p(ypad), _p(z), _p(logdet_half), _p(zz), _p(info), _p(work), _stream())
        _count(2 * nt - 1)
    alpha = None
    if want_alpha:
        alpha = torch.empty(B, npad, dtype=F64, device=dev)
        twork = _work("trsv" + ws_tag, int(_lib.load().gpr_trsv_work_bytes(B)), dev)
        _lib.call("gpr_trsv_backward", _p(L), _p(invdiag), B, npad, _p(z), _p(alpha), _p(twork), _stream())
        _count(1)
    st = FitState(B, n, npad, d, dpad, has_noise, hp, a, nrm, L, invdiag, z, alpha, logdet_half, zz, info, prior, shared)
    st.shared_upper = shared is not None and shared_upper
    return st


_side_streams = {}


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    """One high-priority stream per device for latency chains that should run under a tensor-core-bound launch on the current one."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _side_streams[key]


def can_fit_emulated(n: int, shared: Optional[FitState]) -> bool:
    """Shapes fit_emulated handles: any order from 512 up (padded to whole 128-row blocks like engine.fit); a shared communication
    block must itself be a whole number of blocks."""
    return n >= 4 * TILE and (shared is None or (shared.n % TILE == 0 and shared.B == 1 and shared.n < n))


def fit_emulated_pays(B: int, n: int) -> bool:
    """Where the emulated factorisation beats the FP64 DMMA one (tools/fit_panel_sweep.py, profiles/r02_fit_panel_sweep.jsonl): one
    matrix from order 8192 up (order 4096: 4.8 vs 4.3 ms — the FP64 chain of diagonal tiles is the bound either way), several
    matrices from order 4096 up."""
    return n >= 8192 or (B >= 4 and n >= 4096)


def fit_emulated(x: torch.Tensor, y: torch.Tensor, hp: torch.Tensor, has_noise: bool, slices: int = 7, want_alpha: bool = True,
                 jitter: float = JITTER, L_out: Optional[torch.Tensor] = None, shared: Optional[FitState] = None,
                 panel: int = 0, chunk: int = 32) -> FitState:
    """engine.fit (Exact_GP.update, gpr.py:65-74) with the O(n^3) work on the INT8 tensor cores (opt-in): a
    right-looking blocked Cholesky whose panel solves and trailing updates are emulated-FP64 GEMMs.

    K is built in place (gpr_se_build, lower tiles).  Per panel of `panel` rows, for a chunk of experts at a time:
      D    = the panel's diagonal block, factorised in FP64 by the library's own batched Cholesky (gpr_potrf_batched on a contiguous
             copy; its fused forward solve gives z for these rows, its log-determinant / z.z parts are accumulated),
      P    = K[below, panel] inv(L_DD)^T            one emulated GEMM against X = L_DD^-1 (gpr_tril_inverse), written in place,
      K[below, below] -= P P^T                      one emulated GEMM (alpha = -1, beta = 1, lower tiles only),
      y[below] -= P z_panel                         gpr_gemv_sub (the forward solve's right-hand side).
    With a shared communication block (shared = its fitted state, tile aligned) the first "panel" is the whole shared block:
    its factor is copied, P = K_lg L_g^-T against the communication expert's cached L_g^-1 (one B operand for all experts).
    ~98 % of the flops run as GEMMs; the factor, z, log-determinant and info are those of engine.fit to ~1e-11."""
    B, n, d = x.shape
    assert hp.shape == (B, d + 1 + int(has_noise)) and can_fit_emulated(n, shared)
    lib = _lib.load()
    npad, dev = round_up(n, TILE), x.device
    nt = npad // TILE
    if panel <= 0:
        panel = 2048 if B <= 8 else 1024                   # few matrices: fewer, larger FP64 diagonal blocks (measured, see fit_emulated_pays)
    a, nrm = scale_inputs(x, hp, B, npad)
    dpad = a.shape[2]
    diag_add, prior = split_hp(hp, d, has_noise, jitter)
    L = torch.empty(B, npad, npad, dtype=F64, device=dev) if L_out is None else L_out
    invdiag = torch.empty(B, nt, TILE, TILE, dtype=F64, device=dev)
    yrem = torch.zeros(B, npad, dtype=F64, device=dev)         # right-hand side, updated panel by panel (zero on the identity padding)
    yrem[:, :n] = y
    z = torch.empty(B, npad, dtype=F64, device=dev)
    logdet_half = torch.zeros(B, dtype=F64, device=dev)
    zz = torch.zeros(B, dtype=F64, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    _lib.call("gpr_se_build", _p(a), _p(nrm), _p(hp), hp.shape[-1], _p(diag_add), B, n, npad, dpad, _p(L), npad, 1, 1, _stream())
    _count(1)
    r_start, pXg, shared_upper = 0, None, False
    if shared is not None:
        ng = shared.n
        _lib.call("gpr_copy_leading_block", _p(L), B, npad, _p(shared.L), shared.npad, ng, 1, _stream())
        invdiag[:, : ng // TILE] = shared.invdiag[0, : ng // TILE]
        z[:, :ng] = shared.z[0, :ng]
        logdet_half[:] = shared.logdet_half[0]
        zz[:] = shared.zz[0]
        info[:] = shared.info[0]
        pXg = ozaki_split(tril_inverse(shared)[0, :ng, :ng].contiguous(), slices)
        r_start = ng
    nn = npad * npad

    def panel_update(e0, ne, r0, r1, r2, pX, zsrc, z_stride, b_shared, ready):
        """rows below [r0, r1): P = K[below, r0:r1] X^T, K[below, below] -= P P^T, y[below] -= P z[r0:r1].  The trailing update goes in
        two products: first the columns [r1, r2) of the NEXT panel (after it — and the right-hand side update — the next diagonal
        block and its part of y are final: `ready` is recorded), then the rest."""
        h, hR, h2 = r1 - r0, npad - r1, r2 - r1
        win = L[e0, r1:, r0:]                                  # [hR, .] window of the first expert of the chunk; the others nn apart
        pA = ozaki_split(win, slices, batch=ne, batch_stride=nn, rows=hR, k=h)
        ozaki_gemm_nt(pA, pX, out=L[e0:e0 + ne, r1:, r0:r1], batch=ne, b_lower=0, b_shared=b_shared)
        pP = ozaki_split(win, slices, batch=ne, batch_stride=nn, rows=hR, k=h)
        ozaki_gemm_nt(pP, pP, out=L[e0:e0 + ne, r1:, r1:r2], batch=ne, beta=1.0, alpha=-1.0, c_lower=True, a_win=(0, hR), b_win=(0, h2))
        _lib.call("gpr_gemv_sub", _p(win), npad, nn, ne, hR, h, _p(zsrc), z_stride, _p(yrem[e0:e0 + ne, r1:]), yrem.stride(0), _stream())
        _count(1)
        ready.record(main)
        if r2 < npad:
            ozaki_gemm_nt(pP, pP, out=L[e0:e0 + ne, r2:, r2:], batch=ne, beta=1.0, alpha=-1.0, c_lower=True,
                          a_win=(h2, hR - h2), b_win=(h2, hR - h2))

    def factor_diag(e0, ne, r0, r1):
        """FP64: the diagonal block [r0, r1) of the chunk's experts (+ its part of the forward solve), its inverse as digit planes for the
        panel solve.  Runs on the side stream, under the previous panel's trailing update."""
        h = r1 - r0
        D = L[e0:e0 + ne, r0:r1, r0:r1].contiguous()
        yp = yrem[e0:e0 + ne, r0:r1].contiguous()
        invd, zp, ld, zzp, inf = potrf_inplace(D, yp)
        L[e0:e0 + ne, r0:r1, r0:r1] = D
        invdiag[e0:e0 + ne, r0 // TILE:r1 // TILE] = invd
        z[e0:e0 + ne, r0:r1] = zp
        logdet_half[e0:e0 + ne] += ld
        zz[e0:e0 + ne] += zzp
        cur = info[e0:e0 + ne]
        info[e0:e0 + ne] = torch.where((cur == 0) & (inf > 0), inf + r0, cur)
        pX = None
        if r1 < npad:
            X = torch.empty(ne, h, h, dtype=F64, device=dev)
            scratch = _work("emf_potri", int(lib.gpr_potri_scratch_bytes(ne, h)), dev)
            _lib.call("gpr_tril_inverse", _p(D), _p(invd), ne, h, _p(X), _p(scratch), _stream())
            _count(3)
            pX = ozaki_split(X.view(ne * h, h), slices)
            for t in (pX.planes, pX.scale):                    # allocated on the side stream, read by the main one
                t.record_stream(main)
        zp.record_stream(main)
        return pX, zp

    # experts per chunk: what the free HBM holds of the transient digit planes (two sets of the widest [below x panel] window)
    widest = (npad - r_start) * max(r_start, min(panel, npad - r_start))      # elements of the widest [below x panel] window
    hmax = max(r_start if shared is None else 0, min(panel, npad - r_start))  # diagonal blocks inverted per chunk (not the shared one)
    per = 2 * slices * widest + hmax * hmax * (16 + slices) + (1 << 20)
    free, _ = torch.cuda.mem_get_info(dev)
    E = int(max(1, min(chunk, B, (free - (2 << 30)) // per)))
    # Look-ahead over panels: the FP64 work on diagonal block p+1 (a latency chain that occupies a few SMs) runs on a high-priority side
    # stream while the tensor cores finish panel p's trailing update on the main one.
    from .config import options
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if options.emulate_fit_lookahead else main
    for e0 in range(0, B, E):
        ne = min(e0 + E, B) - e0
        ready, done = torch.cuda.Event(), torch.cuda.Event()
        if shared is not None:
            panel_update(e0, ne, 0, r_start, min(r_start + panel, npad), pXg, z[e0:e0 + ne], z.stride(0), True, ready)
        else:
            ready.record(main)
        r0 = r_start
        r1 = min(r0 + panel, npad)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            pX, zp = factor_diag(e0, ne, r0, r1)
            done.record(side)
        while True:
            main.wait_event(done)
            if r1 >= npad:
                break
            r2 = min(r1 + panel, npad)
            ready, done = torch.cuda.Event(), torch.cuda.Event()
            panel_update(e0, ne, r0, r1, r2, pX, zp, zp.stride(0), False, ready)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                pX, zp = factor_diag(e0, ne, r1, r2)
                done.record(side)
            r0, r1 = r1, r2
    alpha = None
    if want_alpha:
        alpha = torch.empty(B, npad, dtype=F64, device=dev)
        twork = _work("trsv", int(lib.gpr_trsv_work_bytes(B)), dev)
        _lib.call("gpr_trsv_backward", _p(L), _p(invdiag), B, npad, _p(z), _p(alpha), _p(twork), _stream())
        _count(1)
    return FitState(B, n, npad, d, dpad, has_noise, hp, a, nrm, L, invdiag, z, alpha, logdet_half, zz, info, prior, shared)


def fit_grbcm(xg: torch.Tensor, yg: torch.Tensor, hp_g: torch.Tensor, x: torch.Tensor, y: torch.Tensor, hp: torch.Tensor,
              has_noise: bool, jitter: float = JITTER, L_out: Optional[torch.Tensor] = None) -> Tuple[FitState, FitState]:
    """GRBCM's two factorisations (gr_bcm.py:28-29 -> gpr.py:65-74) in ONE launch, for the shared-hyper-parameter case:
    the communication expert on (xg [1,ng,d], yg [1,ng], hp_g [1,P]) and the B local experts x [B,n,d] whose first ng rows
    are xg and whose hp rows all equal hp_g (the caller vouches for both; ng % 128 == 0).  Equivalent to
    ``st_g = fit(xg, yg, hp_g); st_l = fit(x, y, hp, shared=st_g)`` but the communication expert's serial chain of block
    columns runs under the local experts' tiles (gpr_potrf_grbcm).  Returns (st_g, st_l)."""
    B, n, d = x.shape
    ng = xg.shape[1]
    assert xg.shape[0] == 1 and ng >= TILE and ng < n and hp.shape == (B, d + 1 + int(has_noise)) and hp_g.shape[0] == 1
    dev = x.device
    npad, gt, ngp = round_up(n, TILE), ng // TILE, round_up(ng, TILE)
    ngs = gt * TILE                                       # shared rows: the whole 128-row blocks of the communication set
    nt = npad // TILE
    lib = _lib.load()

    def buffers(Bk, nk, npk):
        return (torch.empty(Bk, npk // TILE, TILE, TILE, dtype=F64, device=dev), torch.zeros(Bk, npk, dtype=F64, device=dev),
                torch.empty(Bk, npk, dtype=F64, device=dev), torch.empty(Bk, dtype=F64, device=dev),
                torch.empty(Bk, dtype=F64, device=dev), torch.empty(Bk, dtype=torch.int32, device=dev))

    a_g, nrm_g = scale_inputs(xg, hp_g, 1, ngp)
    a, nrm = scale_inputs(x, hp, B, npad)
    dpad = a.shape[2]
    dadd_g, prior_g = split_hp(hp_g, d, has_noise, jitter)
    dadd, prior = split_hp(hp, d, has_noise, jitter)
    Lg = torch.empty(1, ngp, ngp, dtype=F64, device=dev)
    L = torch.empty(B, npad, npad, dtype=F64, device=dev) if L_out is None else L_out
    inv_g, yp_g, z_g, ld_g, zz_g, info_g = buffers(1, ng, ngp)
    inv, yp, z, ld, zz, info = buffers(B, n, npad)
    yp_g[:, :ng] = yg
    yp[:, :n] = y
    work = _work("potrf_grbcm", int(lib.gpr_potrf_grbcm_work_bytes(B, npad, ngp)), dev)
    _lib.call("gpr_potrf_grbcm", _p(L), _p(a), _p(nrm), _p(hp), hp.shape[-1], _p(dadd), B, n, npad, dpad, _p(inv), _p(yp), _p(z),
              _p(ld), _p(zz), _p(info), _p(Lg), _p(a_g), _p(nrm_g), _p(hp_g), _p(dadd_g), ng, ngp, _p(inv_g), _p(yp_g), _p(z_g),
              _p(ld_g), _p(zz_g), _p(info_g), _p(work), _stream())
    # the leading block itself, for the kernels that walk a whole factor (backward solve, gradient, full covariance)
    _lib.call("gpr_copy_leading_block", _p(L), B, npad, _p(Lg), ngp, ngs, 1, _stream())
    inv[:, :gt] = inv_g[0, :gt]
    _count(2)
    alpha_g = torch.empty(1, ngp, dtype=F64, device=dev)
    alpha = torch.empty(B, npad, dtype=F64, device=dev)
    for (Lk, ik, Bk, npk, zk, ak, tag) in ((Lg, inv_g, 1, ngp, z_g, alpha_g, "_g"), (L, inv, B, npad, z, alpha, "")):
        twork = _work("trsv" + tag, int(lib.gpr_trsv_work_bytes(Bk)), dev)
        _lib.call("gpr_trsv_backward", _p(Lk), _p(ik), Bk, npk, _p(zk), _p(ak), _p(twork), _stream())
        _count(1)
    st_g = FitState(1, ng, ngp, d, dpad, has_noise, hp_g, a_g, nrm_g, Lg, inv_g, z_g, alpha_g, ld_g, zz_g, info_g, prior_g)
    st_l = FitState(B, n, npad, d, dpad, has_noise, hp, a, nrm, L, inv, z, alpha, ld, zz, info, prior, st_g)
    return st_g, st_l


def potrf_inplace(K: torch.Tensor, y: Optional[torch.Tensor] = None):
    """In-place batched Cholesky of a padded [B, npad, npad] buffer holding the lower triangle of K
    (identity on the padding).  Returns (invdiag, z, logdet_half, zz, info).  tc.cholesky, gpr.py:69."""
    B, npad, _ = K.shape
    nt = npad // TILE
    dev = K.device
    invdiag = torch.empty(B, nt, TILE, TILE, dtype=F64, device=dev)
    logdet_half = torch.empty(B, dtype=F64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    z = zz = None
    if y is not None:
        z = torch.empty(B, npad, dtype=F64, device=dev)
        zz = torch.empty(B, dtype=F64, device=dev)
    work = _work("potrf", int(_lib.load().gpr_potrf_work_bytes(B, npad)), dev)
    _lib.call("gpr_potrf_batched", _p(K), 0, None, None, None, 0, None, B, npad, npad, 0,
              _p(invdiag), _p(y), _p(z), _p(logdet_half), _p(zz), _p(info), _p(work), _stream())
    _count(2 * nt - 1)
    return invdiag, z, logdet_half, zz, info


_scratch_cache = {}


def release_scratch() -> None:
    """Drop the cached workspaces (they are grow-only; a long-lived process that once ran a huge case can hand them back)."""
    _scratch_cache.clear()


def _work(kind: str, nbytes: int, dev: torch.device) -> torch.Tensor:
    """Grow-only per-device workspace (the library never allocates; calls on one stream are ordered)."""
    key = (dev.index, kind)
    buf = _scratch_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _scratch_cache[key] = buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    return buf


_sm_count = {}


def _use_gemm_form(st: FitState, m: int) -> bool:
    """Prediction form: blocked substitution (one serial task per expert and 128 test points — the bench path) or the
    triangular GEMM with L^-1 (independent tiles; needs the inverse once per fit).  PYGPR_PREDICT_FORM=subst|gemm forces."""
    from .config import options
    form = options.predict_form
    if form in ("subst", "gemm"):
        return form == "gemm"
    dev = st.L.device.index
    if dev not in _sm_count:
        _sm_count[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    tasks = st.B * ((m + TILE - 1) // TILE)
    vt_bytes = tasks * TILE * st.npad * 8
    return st.npad >= 2 * TILE and 2 * tasks < _sm_count[dev] and vt_bytes + st.B * st.npad * st.npad * 8 < (8 << 30)


def predict(st: FitState, xs: torch.Tensor, want_var: bool = True, scratch_kind: str = "predict", allow_gemm: bool = True
            ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """mean [B, m] and variance [B, m] of every expert at xs [m, d]  (gpr.py:76-106)."""
    assert st.alpha is not None
    m = xs.shape[0]
    dev = xs.device
    from .config import options
    if want_var and options.emulate_slices > 0 and can_emulate_plain(st) and st.npad >= options.emulate_min_order:
        return predict_emulated(st, xs, options.emulate_slices, options.emulate_chunk, tag="_" + scratch_kind)
    mean = torch.empty(st.B, m, dtype=F64, device=dev)
    var = torch.empty(st.B, m, dtype=F64, device=dev) if want_var else None
    if want_var and allow_gemm and _use_gemm_form(st, m):
        # few (expert, 128-test-point) tasks — a single large GP, a handful of experts, few test points: the serial
        # substitutions would leave most SMs idle, the triangular GEMM with the cached L^-1 has (m/128)(n/128) tiles per expert
        lib = _lib.load()
        mt = (m + TILE - 1) // TILE
        Linv = tril_inverse(st)
        vt = _work(scratch_kind + "_vt", st.B * mt * TILE * st.npad * 8, dev)
        scratch = _work(scratch_kind + "_gemm", int(lib.gpr_predict_gemm_scratch_bytes(st.B, st.npad, st.dpad, m)), dev)
        _lib.call("gpr_predict_gemm", _p(Linv), _p(st.a), _p(st.nrm), _p(st.z), _p(st.hp), st.hp.shape[-1], _p(st.prior), _p(xs),
                  st.B, st.n, st.npad, st.d, st.dpad, m, _p(mean), _p(var), _p(vt), _p(scratch), _stream(), 0, None, None)
        _count(4)
        return mean, var
    scratch = _work(scratch_kind, int(_lib.load().gpr_predict_scratch_bytes(st.npad)), dev)
    _lib.call("gpr_predict", _p(st.L), _p(st.invdiag), _p(st.a), _p(st.nrm), _p(st.alpha), _p(st.hp), st.hp.shape[-1],
              _p(st.prior), _p(xs), st.B, st.n, st.npad, st.d, st.dpad, m, _p(mean), _p(var), int(want_var),
              _p(scratch), _stream())
    _count(1)
    return mean, var


_side_streams = {}


def side_stream(dev: torch.device, high_priority: bool = True):
    """Side streams for the communication expert's work, one per device and priority.  Its factorisation (dozens of small
    dependent launches) goes on the HIGH-priority one so each launch gets an SM as soon as one frees up instead of queueing
    behind the local experts' launch; its prediction goes on the normal-priority one, see predict_pair."""
    key = (dev.index, high_priority)
    side = _side_streams.get(key)
    if side is None:
        side = _side_streams[key] = torch.cuda.Stream(dev, priority=-1 if high_priority else 0)
    return side


def predict_pair(st_g: FitState, st_l: FitState, xs: torch.Tensor):
    """GRBCM.predict's two predictions (gr_bcm.py:152-153) issued so that they overlap.  The local experts' persistent
    kernel (one CTA per SM, tasks of ~n^2*128 flops off an atomic counter) goes FIRST on the current stream; the
    communication expert's kernel (ceil(m/128) tasks, each 1/4 the size) follows on a side stream and has no SM to run on
    until persistent CTAs retire — so its short tasks fill the ragged last wave of the big launch instead of adding
    their own time (measured at 32 experts, the 8-GPU share of config 4: 281 -> 273 ms per 4096 test points).
    Returns ((mean_g, var_g), (mean_l, var_l)), all ordered on the current stream."""
    dev = xs.device
    if st_l.shared_from is st_g:
        from .config import options
        if options.emulate_slices > 0 and can_emulate(st_g, st_l) and st_l.npad >= options.emulate_min_order:
            return predict_shared_emulated(st_g, st_l, xs, options.emulate_slices, options.emulate_chunk)
        return predict_shared(st_g, st_l, xs)
    main = torch.cuda.current_stream(dev)
    side = side_stream(dev, high_priority=False)
    side.wait_stream(main)                    # inputs are ready; taken BEFORE the big launch so the side stream does not wait for it
    ml, vl = predict(st_l, xs, True)
    with torch.cuda.stream(side):
        mg, vg = predict(st_g, xs, True, scratch_kind="predict_side", allow_gemm=False)   # substitution tasks: tail fillers
    main.wait_stream(side)
    mg.record_stream(main)
    vg.record_stream(main)
    return (mg, vg), (ml, vl)


def tril_inverse(st: FitState) -> torch.Tensor:
    """L^-1 of a fitted state (cached).  Overwrites the strict block-upper part of st.L with L^-T, like lml_grad."""
    if st.Linv is None:
        X = torch.empty(st.B, st.npad, st.npad, dtype=F64, device=st.L.device)
        scratch = torch.empty(int(_lib.load().gpr_potri_scratch_bytes(st.B, st.npad)), dtype=torch.uint8, device=st.L.device)
        _lib.call("gpr_tril_inverse", _p(st.L), _p(st.invdiag), st.B, st.npad, _p(X), _p(scratch), _stream())
        _count(3)
        st.Linv = X
        st.upper_is_inverse = True
    return st.Linv


def predict_shared(st_g: FitState, st_l: FitState, xs: torch.Tensor):
    """The two predictions of GRBCM.predict when the local experts were fitted on top of the communication expert's factor
    (fit(shared=st_g)): the communication expert's V_g^T = K*_g L_g^-T is kept and every local expert continues the
    substitution from block row floor(ng/128) (gpr_predict_gemm + gpr_predict_shared).  Same return value as predict_pair."""
    assert st_l.shared_from is st_g and st_g.B == 1
    m = xs.shape[0]
    dev = xs.device
    lib = _lib.load()
    mt = (m + TILE - 1) // TILE
    ngs = shared_rows(st_g)
    mg = torch.empty(1, m, dtype=F64, device=dev)
    vg = torch.empty(1, m, dtype=F64, device=dev)
    # the same sums over the shared rows only (== mg, vg when the communication set is a whole number of blocks)
    mg_s = mg if ngs == st_g.n else torch.empty(1, m, dtype=F64, device=dev)
    vg_s = vg if ngs == st_g.n else torch.empty(1, m, dtype=F64, device=dev)
    vt = _work("predict_vt", mt * TILE * st_g.npad * 8, dev)
    # GEMM form: (m/128) x (ng/128) independent tiles instead of m/128 serial substitutions
    Linv = tril_inverse(st_g)
    scratch_g = _work("predict_gemm", int(lib.gpr_predict_gemm_scratch_bytes(1, st_g.npad, st_g.dpad, m)), dev)
    _lib.call("gpr_predict_gemm", _p(Linv), _p(st_g.a), _p(st_g.nrm), _p(st_g.z), _p(st_g.hp), st_g.hp.shape[-1],
              _p(st_g.prior), _p(xs), 1, st_g.n, st_g.npad, st_g.d, st_g.dpad, m, _p(mg), _p(vg), _p(vt), _p(scratch_g), _stream(),
              ngs if ngs != st_g.n else 0, _p(mg_s) if ngs != st_g.n else None, _p(vg_s) if ngs != st_g.n else None)
    _count(3)
    ml = torch.empty(st_l.B, m, dtype=F64, device=dev)
    vl = torch.empty(st_l.B, m, dtype=F64, device=dev)
    scratch = _work("predict", int(lib.gpr_predict_scratch_bytes(st_l.npad)), dev)
    _lib.call("gpr_predict_shared", _p(st_l.L), _p(st_l.invdiag), _p(st_l.a), _p(st_l.nrm), _p(st_l.z), _p(st_l.hp),
              st_l.hp.shape[-1], _p(xs), st_l.B, st_l.n, st_l.npad, st_l.d, st_l.dpad, m, _p(ml), _p(vl), _p(scratch), _stream(),
              ngs, _p(vt), st_g.npad, _p(mg_s), _p(vg_s))
    _count(2)
    return (mg, vg), (ml, vl)


def kinv(st: FitState) -> torch.Tensor:
    """K^-1 [B, n, n] of a fitted state (cached): tc.cholesky_inverse.  Leaves L^-T in the strict block-upper part of st.L."""
    if st.Kinv is None:
        out = torch.empty(st.B, st.n, st.n, dtype=F64, device=st.L.device)
        scratch = torch.empty(int(_lib.load().gpr_potri_scratch_bytes(st.B, st.npad)), dtype=torch.uint8, device=st.L.device)
        _lib.call("gpr_potri_batched", _p(st.L), _p(st.invdiag), st.B, st.n, st.npad, _p(out), st.n, _p(scratch), _stream())
        _count(3)
        st.Kinv = out
    return st.Kinv


def lml_grad_emulated(st: FitState, slices: int = 7, chunk: int = 8) -> torch.Tensor:
    """lml_grad with K^-1 = U U^T (U = L^-T) on the INT8 tensor cores (opt-in): L^-T stays the FP64 blocked
    trtri (in place), then per chunk of experts the rows of U go through gpr_triu_rows -> gpr_ozaki_split, ONE batched
    gpr_ozaki_gemm_nt (a_upper: the zero k-blocks before the diagonal are skipped; c_lower: only the lower tiles) writes K^-1,
    and gpr_grad_from_kinv runs gpr_lml_grad's fused reduction on it (dK never formed).  loss.py:59-128."""
    assert st.alpha is not None
    ensure_inverse_factors(st.shared_from, st)               # L^-T of every expert, in place (the shared block's is copied)
    dev, B, npad = st.L.device, st.B, st.npad
    P = st.d + 1 + int(st.has_noise)
    grad = torch.empty(B, P, dtype=F64, device=dev)
    # experts per batched product: what the free HBM holds of the three staging buffers (rows of U, its digit planes, K^-1)
    per = npad * npad * (16 + slices) + (1 << 20)
    held = sum(b.numel() for k, b in _scratch_cache.items() if k[0] == dev.index and str(k[1]).startswith("emg_"))
    free, _ = torch.cuda.mem_get_info(dev)
    E = int(max(1, min(chunk, B, (free + held - (2 << 30)) // per)))
    cache = getattr(st, "plane_cache", None)
    if E < min(chunk, B) and cache and len(cache) > 2:
        # the prediction's cached digit planes (kept while HBM is idle) are in the way of a full chunk: give them back
        cache.clear()
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info(dev)
        E = int(max(1, min(chunk, B, (free + held - (2 << 30)) // per)))
    Ud = _f64_work("emg_u", E * npad * npad, dev).view(E, npad, npad)
    Kinv = _f64_work("emg_kinv", E * npad * npad, dev).view(E, npad, npad)
    buf = DigitPlanes(E * npad, npad, slices, _work("emg_p", slices * E * npad * npad, dev), _f64_work("emg_s", E * npad, dev))
    scratch = _work("emg_scratch", int(_lib.load().gpr_lml_grad_scratch_bytes(E, npad, st.d)), dev)
    for e0 in range(0, B, E):
        e1 = min(e0 + E, B)
        ne = e1 - e0
        _lib.call("gpr_triu_rows", _p(st.L[e0:e1]), _p(st.invdiag[e0:e1]), ne, npad, _p(Ud), _stream())
        pu = ozaki_split(Ud[:ne].reshape(ne * npad, npad), slices, out=buf)
        ozaki_gemm_nt(pu, pu, out=Kinv[:ne], batch=ne, a_upper=0, c_lower=True)
        _lib.call("gpr_grad_from_kinv", _p(Kinv), npad, _p(st.a[e0:e1]), _p(st.nrm[e0:e1]), _p(st.alpha[e0:e1]), _p(st.hp[e0:e1]),
                  st.hp.shape[-1], int(st.has_noise), ne, st.n, npad, st.d, st.dpad, _p(grad[e0:e1]), _p(scratch), _stream())
        _count(4)
    return grad


def lml_grad(st: FitState) -> torch.Tensor:
    """d NLML / d hp [B, P] from a finished fit (destroys the strict block-upper part of st.L)   loss.py:59-128."""
    assert st.alpha is not None
    from .config import options
    if (options.emulate_grad_slices > 0 and st.npad > TILE and st.npad >= options.emulate_min_order
            and (st.shared_from is None or st.shared_from.n % TILE == 0)):
        return lml_grad_emulated(st, options.emulate_grad_slices, options.emulate_chunk)
    P = st.d + 1 + int(st.has_noise)
    grad = torch.empty(st.B, P, dtype=F64, device=st.L.device)
    nbytes = int(_lib.load().gpr_lml_grad_scratch_bytes(st.B, st.npad, st.d))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=st.L.device)
    sg = st.shared_from
    if sg is not None and sg.n % TILE == 0 and not (st.shared_upper and sg.Kinv is not None):
        # first gradient on a shared-block fit: form K_g^-1 / L_g^-T once and fetch L_g^-T into every expert's leading block
        kinv(sg)
        _lib.call("gpr_copy_leading_block", _p(st.L), st.B, st.npad, _p(sg.L), sg.npad, sg.n, 2, _stream())
        _count(1)
        st.shared_upper = True
    if sg is not None and st.shared_upper and sg.Kinv is not None and sg.n % TILE == 0:
        # shared leading block: neither L_g^-T nor K_g^-1 is recomputed per expert (gpr_lml_grad_shared)
        _lib.call("gpr_lml_grad_shared", _p(st.L), _p(st.invdiag), _p(st.a), _p(st.nrm), _p(st.alpha), _p(st.hp), st.hp.shape[-1],
                  int(st.has_noise), st.B, st.n, st.npad, st.d, st.dpad, _p(grad), _p(scratch), _stream(), sg.n, _p(sg.Kinv), sg.n)
        _count(4)
        return grad
    _lib.call("gpr_lml_grad", _p(st.L), _p(st.invdiag), _p(st.a), _p(st.nrm), _p(st.alpha), _p(st.hp), st.hp.shape[-1],
              int(st.has_noise), st.B, st.n, st.npad, st.d, st.dpad, _p(grad), _p(scratch), _stream())
    _count(4 if st.npad > TILE else 3)
    return grad


def grbcm_partial(mean_l: torch.Tensor, var_l: torch.Tensor, var_g: torch.Tensor, first_global_index: int,
                  want_rows: bool = False):
    """Per-rank partial sums S [3, m] over the local experts (gr_bcm.py:128-144, split for sharding)."""
    nc_local, m = mean_l.shape
    S = torch.empty(3, m, dtype=F64, device=var_g.device)
    beta_rows = torch.empty(nc_local, m, dtype=F64, device=var_g.device) if want_rows else None
    prec_rows = torch.empty(nc_local, m, dtype=F64, device=var_g.device) if want_rows else None
    _lib.call("grbcm_partial", _p(mean_l), _p(var_l), _p(var_g), nc_local, first_global_index, m, _p(S),
              _p(beta_rows), _p(prec_rows), _stream())
    _count(1)
    return S, beta_rows, prec_rows


def grbcm_finalize(S: torch.Tensor, mean_g: torch.Tensor, var_g: torch.Tensor, want_rows: bool = False):
    m = mean_g.shape[0]
    mu = torch.empty(m, dtype=F64, device=S.device)
    var = torch.empty(m, dtype=F64, device=S.device)
    beta0 = torch.empty(m, dtype=F64, device=S.device) if want_rows else None
    prec0 = torch.empty(m, dtype=F64, device=S.device) if want_rows else None
    _lib.call("grbcm_finalize", _p(S), _p(mean_g), _p(var_g), m, _p(mu), _p(var), _p(beta0), _p(prec0), _stream())
    _count(1)
    return mu, var, beta0, prec0


# ---- "next" rows: full predictive covariance and its GRBCM aggregation -------------------------------
def predict_full(st: FitState, xs: torch.Tensor, cov=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """mean [B, m] and full predictive covariance [B, m, m] = K** - K* K^-1 K*^T   (gpr.py:108-120).

    K** = cov.kernel(params, xp) includes the white noise on its diagonal (no jitter), gpr.py:110.
    """
    assert st.alpha is not None
    m = xs.shape[0]
    dev = xs.device
    mpad = round_up(m, TILE)
    as_, nrms = scale_inputs(xs, st.hp, st.B, mpad)
    noise_var = (st.hp[:, st.d + 1] * st.hp[:, st.d + 1]).contiguous() if st.has_noise else None
    mean = torch.empty(st.B, m, dtype=F64, device=dev)
    var = torch.empty(st.B, m, dtype=F64, device=dev)
    covm = torch.empty(st.B, m, m, dtype=F64, device=dev)
    scratch = torch.empty(int(_lib.load().gpr_predict_full_scratch_bytes(st.B, st.npad, m)), dtype=torch.uint8, device=dev)
    _lib.call("gpr_predict_full", _p(st.L), _p(st.invdiag), _p(st.a), _p(st.nrm), _p(st.alpha), _p(st.hp), st.hp.shape[-1],
              _p(st.prior), _p(noise_var), _p(xs), _p(as_), _p(nrms), st.B, st.n, st.npad, st.d, st.dpad, m, mpad,
              _p(mean), _p(var), _p(covm), _p(scratch), _stream())
    _count(2)
    return mean, covm


def spd_inverse(mats: torch.Tensor) -> torch.Tensor:
    """Batched inverse of SPD matrices [B, m, m] through the library's Cholesky + potri
    (tc.cholesky + tc.cholesky_solve(identity) / tc.cholesky_inverse, gr_bcm.py:101-103,112)."""
    B, m, _ = mats.shape
    mpad = round_up(m, TILE)
    buf = torch.zeros(B, mpad, mpad, dtype=F64, device=mats.device)
    buf[:, :m, :m] = mats
    if mpad > m:
        idx = torch.arange(m, mpad, device=mats.device)
        buf[:, idx, idx] = 1.0
    invdiag, _, _, _, info = potrf_inplace(buf)
    bad = torch.nonzero(info)
    if bad.numel():
        raise torch.linalg.LinAlgError("cholesky: matrix %d is not positive-definite (leading minor %d)" %
                                       (int(bad[0, 0]), int(info[int(bad[0, 0])])))
    out = torch.empty(B, m, m, dtype=F64, device=mats.device)
    scratch = torch.empty(int(_lib.load().gpr_potri_scratch_bytes(B, mpad)), dtype=torch.uint8, device=mats.device)
    _lib.call("gpr_potri_batched", _p(buf), _p(invdiag), B, m, mpad, _p(out), m, _p(scratch), _stream())
    _count(3)
    return out


def grbcm_aggregate_full(model, ys_g, covars_g, ys_l, covars_l):
    """GRBCM.aggregate(var="full") + aggregate_full_covar (gr_bcm.py:99-114,116-149), experts possibly sharded."""
    from . import dist as gdist
    from .covar import _back
    out_like = ys_g
    mean_g = _dev(ys_g).reshape(-1)
    m = mean_g.shape[0]
    cov_g = _dev(covars_g).reshape(m, m)
    mean_l = _dev(ys_l).reshape(-1, m)
    cov_l = _dev(covars_l).reshape(-1, m, m)
    var_g = torch.diagonal(cov_g).contiguous()                                   # gr_bcm.py:122
    var_l = torch.diagonal(cov_l, dim1=-2, dim2=-1).contiguous()                 # gr_bcm.py:123
    S, brow, prow = grbcm_partial(mean_l, var_l, var_g, model.expert_offset, want_rows=True)
    if model.sharded:
        gdist.all_reduce_sum_(S, model.group)
    _, _, b0, p0 = grbcm_finalize(S, mean_g, var_g, want_rows=True)
    model.beta = _back(torch.cat((b0[None], brow)), out_like)
    model.prec = _back(torch.cat((p0[None], prow)), out_like)
    # prec = sum_c Kinv_c o (beta_c[i] + beta_c[j]) / 2 ; the communication expert's term is added by the rank owning expert 0
    own_g = model.expert_offset == 0
    mats = torch.cat((cov_g[None], cov_l)) if own_g else cov_l
    betas = torch.cat((b0[None], brow)) if own_g else brow
    kinv = spd_inverse(mats.contiguous())                                        # gr_bcm.py:100-103
    prec = torch.empty(m, m, dtype=F64, device=mean_g.device)
    _lib.call("grbcm_full_combine", _p(kinv), _p(betas.contiguous()), betas.shape[0], m, _p(prec), _stream())
    _count(1)
    if model.sharded:
        gdist.all_reduce_sum_(prec, model.group)
    covars = spd_inverse(prec[None])[0]                                          # gr_bcm.py:112
    total = S[2] + mean_g * p0 * b0                                              # (ys * prec * beta).sum(0), gr_bcm.py:147
    mu = total * torch.diagonal(covars)
    return _back(mu, out_like), _back(covars, out_like)


# ---- dense path: model-level support for any Covar composition (kernel matrices materialised by the kernel-level API) ----------
@dataclass
class DenseState:
    """Fit of materialised kernel matrices K [B, n, n] (noise and jitter included): what FitState is to the fused path."""
    B: int
    n: int
    npad: int
    L: torch.Tensor
    invdiag: torch.Tensor
    z: torch.Tensor
    alpha: Optional[torch.Tensor]
    logdet_half: torch.Tensor
    zz: torch.Tensor
    info: torch.Tensor
    Linv: Optional[torch.Tensor] = None
    Kinv: Optional[torch.Tensor] = None
    shared_from = None

    nlml = FitState.nlml
    check = FitState.check


def fit_dense(K: torch.Tensor, y: torch.Tensor, want_alpha: bool = True) -> DenseState:
    """tc.cholesky + tc.cholesky_solve of gpr.py:67-72 / loss.py:37-51 for a materialised K [B, n, n] (device, float64)."""
    B, n, _ = K.shape
    npad = round_up(n, TILE)
    dev = K.device
    buf = torch.zeros(B, npad, npad, dtype=F64, device=dev)
    buf[:, :n, :n] = K
    if npad > n:
        idx = torch.arange(n, npad, device=dev)
        buf[:, idx, idx] = 1.0
    ypad = torch.zeros(B, npad, dtype=F64, device=dev)
    ypad[:, :n] = y
    invdiag, z, logdet_half, zz, info = potrf_inplace(buf, ypad)
    alpha = None
    if want_alpha:
        alpha = torch.empty(B, npad, dtype=F64, device=dev)
        twork = _work("trsv", int(_lib.load().gpr_trsv_work_bytes(B)), dev)
        _lib.call("gpr_trsv_backward", _p(buf), _p(invdiag), B, npad, _p(z), _p(alpha), _p(twork), _stream())
        _count(1)
    return DenseState(B, n, npad, buf, invdiag, z, alpha, logdet_half, zz, info)


def predict_dense(st: DenseState, Ks: torch.Tensor, prior: torch.Tensor, Kss: Optional[torch.Tensor] = None):
    """mean [B, m], variance [B, m] (and the full covariance when Kss [B, m, m] is given) from materialised cross-covariances
    Ks [B, m, n] and the prior variance PER TEST POINT prior [B, m] (= diag cov.kernel(params, xp), gpr.py:98,102):
    gpr.py:76-120 with V^T = K* L^-T as a triangular GEMM against the cached L^-1."""
    B, m, n = Ks.shape
    assert prior.shape == (B, m)
    dev = Ks.device
    lib = _lib.load()
    mpad = round_up(m, TILE)
    ksp = torch.zeros(B, mpad, st.npad, dtype=F64, device=dev)
    ksp[:, :m, :n] = Ks
    X = tril_inverse(st)
    mean = torch.empty(B, m, dtype=F64, device=dev)
    var = torch.empty(B, m, dtype=F64, device=dev)
    vt = torch.empty(B, mpad, st.npad, dtype=F64, device=dev)
    scratch = _work("predict_dense", int(lib.gpr_predict_dense_scratch_bytes(B, st.npad, m)), dev)
    _lib.call("gpr_predict_dense", _p(X), _p(ksp), _p(st.z), _p(prior.contiguous()), B, st.npad, m, _p(mean), _p(var), _p(vt),
              _p(scratch), _stream())
    _count(2)
    covm = None
    if Kss is not None:
        covm = torch.empty(B, m, m, dtype=F64, device=dev)
        _lib.call("gpr_covar_dense", _p(vt), _p(Kss.contiguous()), B, st.npad, m, _p(covm), _stream())
        _count(1)
    return mean, var, covm


def lml_grad_dense(st: DenseState, dK: torch.Tensor) -> torch.Tensor:
    """d NLML / d hp [B, P] = 1/2 sum (K^-1 - alpha alpha^T) o dK_p for materialised dK [B, P, n, n]   (loss.py:77-84,116-121)."""
    B, P, n, _ = dK.shape
    Kinv = kinv(st)
    grad = torch.empty(B, P, dtype=F64, device=dK.device)
    scratch = torch.empty(int(_lib.load().gpr_trace_grad_scratch_bytes(B, P, n)), dtype=torch.uint8, device=dK.device)
    _lib.call("gpr_trace_grad", _p(Kinv), _p(st.alpha), st.npad, _p(dK.contiguous()), B, P, n, _p(grad), _p(scratch), _stream())
    _count(2)
    return grad


# ---- "next" row N4: FP64-accurate products on the INT8 tcgen05 tensor cores (csrc/ozaki_gemm.cu) ------------------------------
@dataclass
class DigitPlanes:
    """Balanced radix-256 digit planes of a matrix X [rows, k] (gpr_ozaki_split): planes [slices, rpad, kpad] int8, scale [rpad]."""
    rows: int
    k: int
    slices: int
    planes: torch.Tensor
    scale: torch.Tensor


def ozaki_split(X: torch.Tensor, slices: int, bound: float = 0.0, out: Optional[DigitPlanes] = None, batch: int = 1,
                batch_stride: int = 0, rows: Optional[int] = None, k: Optional[int] = None) -> DigitPlanes:
    """Digit planes of a row-major FP64 matrix (last dimension contiguous).  bound > 0: |X| <= bound is known, every row is
    scaled by it (one pass over X); otherwise each row by its own maximum.  out: planes / scale buffers to write into.
    batch > 1: `batch` windows of `rows` x `k` elements starting at X's first element, batch_stride doubles apart, row stride
    X.stride(0) (e.g. the same block of every expert's factor); the planes are stacked along the rows (rows % 128 == 0)."""
    assert X.dim() == 2 and X.stride(1) == 1 and X.dtype == F64
    rows = X.shape[0] if rows is None else rows
    k = X.shape[1] if k is None else k
    assert batch == 1 or rows % 128 == 0
    lib = _lib.load()
    rpad, kpad = int(lib.gpr_ozaki_rows_padded(rows)), int(lib.gpr_ozaki_k_padded(k))
    if out is None:
        planes = torch.empty(slices, batch * rpad, kpad, dtype=torch.int8, device=X.device)
        scale = torch.empty(batch * rpad, dtype=F64, device=X.device)
    else:
        planes = out.planes.view(-1)[: slices * batch * rpad * kpad].view(slices, batch * rpad, kpad)
        scale = out.scale[: batch * rpad]
    _lib.call("gpr_ozaki_split", _p(X), X.stride(0), rows, k, slices, float(bound), _p(planes), _p(scale), batch, batch_stride, _stream())
    _count(1)
    return DigitPlanes(batch * rows, k, slices, planes, scale)


def ozaki_gemm_nt(A: DigitPlanes, B: DigitPlanes, out: Optional[torch.Tensor] = None, beta: float = 0.0,
                  b_lower: Optional[int] = None, batch: int = 1, a_upper: Optional[int] = None, c_lower: bool = False,
                  alpha: float = 1.0, b_shared: bool = False, a_win: Optional[Tuple[int, int]] = None,
                  b_win: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """C = beta*C + A B^T from digit planes, on tcgen05.mma kind::i8 (gpr_ozaki_gemm_nt).
    a_win / b_win = (first row, rows): the product uses that window of rows of (every batch entry of) the operand
    (gpr_ozaki_gemm_nt_window; both multiples of 128 here).

    b_lower = o: B[j, k] = 0 for k > j + o (True means o = 0), the zero k-blocks are skipped; a_upper = o: A[i, k] = 0 for
    k < i + o (rows of an upper-triangular factor); c_lower: only the tiles touching the lower triangle of C are computed and
    written.  batch > 1: A and B hold `batch` operands stacked along the rows (rows % 128 == 0 each); the result is [batch, M, N]."""
    assert A.k == B.k and A.slices == B.slices and A.rows % batch == 0 and (b_shared or B.rows % batch == 0)
    Mt, Nt = A.rows // batch, (B.rows if b_shared else B.rows // batch)
    assert batch == 1 or (Mt % 128 == 0 and Nt % 128 == 0)
    (a_off, M), (b_off, N) = (a_win or (0, Mt)), (b_win or (0, Nt))
    windowed = a_win is not None or b_win is not None
    assert a_off % 128 == 0 and b_off % 128 == 0 and a_off + M <= Mt and b_off + N <= Nt and (not windowed or (M % 128 == 0 and N % 128 == 0))
    shape = (M, N) if batch == 1 else (batch, M, N)
    if out is None:
        out = torch.empty(shape, dtype=F64, device=A.planes.device)
        beta = 0.0
    if batch == 1 and out.dim() == 3 and out.shape[0] == 1:
        out = out[0]
    assert tuple(out.shape) == shape and out.stride(-1) == 1 and (batch == 1 or out.stride(0) % 2 == 0)
    off = -1 if b_lower is None or b_lower is False else (0 if b_lower is True else int(b_lower))
    if windowed:
        rnd = lambda v: round_up(v, 128)
        _lib.call("gpr_ozaki_gemm_nt_window", _p(A.planes), _p(A.scale), M, rnd(Mt), a_off, _p(B.planes), _p(B.scale), N, rnd(Nt), b_off,
                  A.k, A.slices, off, float(beta), _p(out), out.stride(-2), batch, out.stride(0) if batch > 1 else 0,
                  -1 if a_upper is None else int(a_upper), int(c_lower), float(alpha), int(b_shared and batch > 1), _stream())
    else:
        _lib.call("gpr_ozaki_gemm_nt", _p(A.planes), _p(A.scale), M, _p(B.planes), _p(B.scale), N, A.k, A.slices,
                  off, float(beta), _p(out), out.stride(-2), batch, out.stride(0) if batch > 1 else 0,
                  -1 if a_upper is None else int(a_upper), int(c_lower), float(alpha), int(b_shared and batch > 1), _stream())
    _count(1)
    return out


def _inverse_plan(nt: int, gt: int, B: int, sms: int):
    """Segment size (in 128-row blocks) for the assembled inverse, or 0 for the plain FP64 trtri: the part [gt, nt) is cut into
    4 or 2 equal diagonal segments of at least 4 blocks (the FP64 chain then works inside segments only and the rest — 3/4 to
    15/16 of the flops — are GEMMs), provided the persistent task schedule of trtri_kernel applies."""
    from .config import options
    if options.emulate_inverse_slices <= 0 or B * (nt - 1) < 2 * sms:
        return 0
    span = nt - gt
    for parts in (4, 2):
        if span % parts == 0 and span // parts >= 4 and (gt == 0 or gt % (span // parts) == 0):
            return span // parts
    return 0


def _inverse_factors(st: FitState, gt: int) -> None:
    """U = L^-T of every expert, in place in the strict block-upper part of st.L (block columns < gt already hold the shared L_g^-T).

    Plain form: the FP64 blocked trtri (gpr_trtri_batched), a triangular-solve chain at the DMMA rate.  Assembled form (opt-in,
    config.options.emulate_inverse_slices): trtri only inside diagonal segments; every block between two adjacent groups of
    segments A (above) and B (below) is   U_AB = -U_AA L_BA^T U_BB   — two products on the INT8 tensor cores per merge
    (gpr_ozaki_gemm_nt; T^T = U_AA L_BA^T, then U_AB = -T^T X_BB^T with X_BB = U_BB^T), merged pairwise up to the whole matrix."""
    from .config import options
    lib = _lib.load()
    dev, B, npad = st.L.device, st.B, st.npad
    nt = npad // TILE
    if dev.index not in _sm_count:
        _sm_count[dev.index] = torch.cuda.get_device_properties(dev.index).multi_processor_count
    seg = _inverse_plan(nt, gt, B, _sm_count[dev.index])
    scratch = torch.empty(int(lib.gpr_potri_scratch_bytes(B, npad)), dtype=torch.uint8, device=dev)
    _lib.call("gpr_trtri_batched", _p(st.L), _p(st.invdiag), B, npad, _p(scratch), _stream(), gt, seg)
    _count(2)
    if seg == 0:
        return
    slices = options.emulate_inverse_slices
    # merges, finest first: pairs of segments inside [gt, nt), then pairs of pairs, ..., finally the shared block against the rest
    merges, width = [], seg
    while width < nt - gt:
        for a0 in range(gt, nt, 2 * width):
            merges.append((a0, a0 + width, a0 + 2 * width))
        width *= 2
    if gt > 0:
        merges.append((0, gt, nt))
    hmax = max(max(a1 - a0, a2 - a1) for a0, a1, a2 in merges) * TILE
    per = hmax * hmax * (24 + 4 * slices) + (1 << 20)          # three FP64 windows + four plane sets per expert of a chunk
    free, _ = torch.cuda.mem_get_info(dev)
    held = sum(b.numel() for k, b in _scratch_cache.items() if k[0] == dev.index and str(k[1]).startswith("inv_"))
    E = int(max(1, min(options.emulate_chunk, B, (free + held - (2 << 30)) // per)))
    for e0 in range(0, B, E):
        ne = min(e0 + E, B) - e0
        Lc, Ic = st.L[e0:e0 + ne], st.invdiag[e0:e0 + ne]
        for a0, a1, a2 in merges:
            hA, hB = (a1 - a0) * TILE, (a2 - a1) * TILE
            UA = _f64_work("inv_ua", ne * hA * hA, dev).view(ne * hA, hA)
            XB = _f64_work("inv_xb", ne * hB * hB, dev).view(ne * hB, hB)
            Tt = _f64_work("inv_tt", ne * hA * hB, dev).view(ne, hA, hB)
            _lib.call("gpr_factor_window", _p(Lc), _p(Ic), ne, npad, 0, a0 * TILE, hA, a0 * TILE, hA, _p(UA), _stream())   # U_AA
            _lib.call("gpr_factor_window", _p(Lc), _p(Ic), ne, npad, 1, a1 * TILE, hB, a1 * TILE, hB, _p(XB), _stream())   # X_BB
            pUA = ozaki_split(UA, slices)
            pXB = ozaki_split(XB, slices)
            # L_BA straight out of the factor: the same [hB x hA] window of every expert, npad^2 doubles apart
            LBA = Lc[0, a1 * TILE:, a0 * TILE:]
            pLBA = ozaki_split(LBA, slices, batch=ne, batch_stride=npad * npad, rows=hB, k=hA)
            ozaki_gemm_nt(pUA, pLBA, out=Tt, batch=ne, a_upper=0)                      # T^T = U_AA L_BA^T   [hA, hB]
            pTt = ozaki_split(Tt.view(ne * hA, hB), slices)
            UAB = Lc[:, a0 * TILE:a1 * TILE, a1 * TILE:a2 * TILE]                       # written in place, row stride npad
            ozaki_gemm_nt(pTt, pXB, out=UAB, batch=ne, b_lower=0, alpha=-1.0)          # U_AB = -T^T X_BB^T
            _count(2)


def ensure_inverse_factors(st_g: Optional[FitState], st_l: FitState) -> None:
    """L^-T of every expert of st_l, in place in the strict block-upper part of its factor (once per fit): the emulated-FP64
    prediction multiplies with rows of the explicit inverse.  With a shared communication block (st_g) its L_g^-T is copied into
    the leading block, not recomputed."""
    if getattr(st_l, "upper_is_inverse", False):
        if st_l.plane_cache is None:                          # L^-T was put in place by tril_inverse: the planes of its rows may be kept too
            st_l.plane_cache = {}
        return
    gt = 0
    if st_g is not None:
        assert st_l.shared_from is st_g and st_g.n % TILE == 0
        if not getattr(st_g, "upper_is_inverse", False):
            tril_inverse(st_g)                                # L_g^-T into the block-upper part of the communication factor
        if not st_l.shared_upper:
            _lib.call("gpr_copy_leading_block", _p(st_l.L), st_l.B, st_l.npad, _p(st_g.L), st_g.npad, st_g.n, 2, _stream())
            _count(1)
            st_l.shared_upper = True
        gt = st_g.n // TILE
    if st_l.npad > TILE:
        _inverse_factors(st_l, gt)
    st_l.upper_is_inverse = True
    # the digit planes of the inverse rows are kept across predictions for as many expert chunks as the free HBM holds beside the
    # transient chunk buffers (config 4 on ONE GPU: 137 GB of factors leave room for about a third of the 51 GB of planes)
    st_l.plane_cache = {}


PLANE_CACHE_RESERVE = 6 << 30    # HBM left alone by the plane cache: transient buffers of later calls (lml_grad_emulated takes the cache back if it needs more)


def can_emulate(st_g: FitState, st_l: FitState) -> bool:
    """Shapes the emulated-FP64 prediction of a shared-block GRBCM handles: the shared block a whole number of 128-row blocks (the
    experts' own rows may be ragged: they are padded like the factor)."""
    return (st_l.shared_from is st_g and st_g.B == 1 and st_g.n % TILE == 0 and st_l.n > st_g.n)


def can_emulate_plain(st) -> bool:
    """A fused fit (not the dense path) whose experts stand alone; any order (padded to whole 128-row blocks like the factor)."""
    return isinstance(st, FitState) and st.alpha is not None


def _f64_work(kind: str, numel: int, dev: torch.device) -> torch.Tensor:
    return _work(kind, numel * 8, dev)[: numel * 8].view(F64)


def _emulated_experts(st: FitState, xs: torch.Tensor, row0: int, mean0: torch.Tensor, var0: torch.Tensor, slices: int, chunk: int,
                      tag: str = ""):
    """mean [B, m], var [B, m] of the experts of `st` from ONE emulated-FP64 GEMM per expert:
        V^T = K* X^T,  X = rows row0.. of L^-1 (lower trapezoidal),  mean = mean0 + V^T z[row0:],  var = var0 - |V|^2
    (gpr.py:79-85, 100-104; mean0 / var0 [m] carry the part of the sums that belongs to the rows before row0 — the communication
    expert's prediction for a shared-block GRBCM, 0 and the prior variance otherwise).  All experts of `st` share one
    hyper-parameter vector when row0 > 0; otherwise each expert's own amplitude bounds its K*.  xs: [m, d], m % 128 == 0."""
    m, dev = xs.shape[0], xs.device
    n, npad, B = st.n, st.npad, st.B
    nl = npad - row0                                          # rows of L^-1 used, identity padding included (its K* columns are zero)
    mean = torch.empty(B, m, dtype=F64, device=dev)
    var = torch.empty(B, m, dtype=F64, device=dev)
    cache = getattr(st, "plane_cache", None)
    E = max(1, min(chunk, B))
    if cache is not None and cache.get("layout") != (E, slices):
        cache.clear()                                         # planes cached for another chunking / plane count: start over
        cache["layout"] = (E, slices)
        cache["full"] = False                                 # set once a chunk did not fit: later chunks stay transient
    per_expert_var0 = var0.dim() == 2
    sig2 = getattr(st, "sig2_max", None)                      # 0 < K* <= sig^2 (covar.py:149): the bound of the digit split
    if sig2 is None:                                          # read back once per fit: a device->host sync per prediction would
        sig2 = st.sig2_max = float((st.hp[:, 0] * st.hp[:, 0]).max())   # keep the host from queueing ahead of the device
    # transient staging, reused by every chunk (stream-ordered): K* and X in FP64, their digit planes, V^T.  (Preparing chunk c+1 on
    # a side stream while the tensor cores work on chunk c was measured and dropped: the GEMM's 222 KB of shared memory per SM
    # leaves no room for the kernel-matrix builder beside it, so the two only time-slice.)
    # K* goes straight into its digit planes when the fused builder's staging fits (gpr_se_cross_planes); else FP64 K*, then split
    fused = bool(_lib.load().gpr_se_cross_planes_ok(st.dpad, slices))
    Ks = None
    if not fused:
        Ks = _f64_work("emu_ks" + tag, E * m * npad, dev).view(E, m, npad)
        if npad > n:
            Ks[:, :, n:].zero_()                              # columns of the padding: gpr_se_cross_ld writes the first n of every row
    X = _f64_work("emu_x" + tag, E * nl * npad, dev).view(E, nl, npad)
    V = _f64_work("emu_v" + tag, E * m * nl, dev).view(E, m, nl)
    bufa = DigitPlanes(E * m, npad, slices, _work("emu_pa" + tag, slices * E * m * npad, dev), _f64_work("emu_sa" + tag, E * m, dev))
    bufb = None                                               # transient planes of X, allocated when a chunk is not (or cannot be) cached
    for e0 in range(0, B, E):
        e1 = min(e0 + E, B)
        ne = e1 - e0
        # scaled test points of the chunk's experts (per-expert length scales), K* [ne, m, n] (covar.py:151-160), its digit planes
        as_e, nrms_e = scale_inputs(xs, st.hp[e0:e1], ne, m)
        if fused:
            pa = DigitPlanes(ne * m, npad, slices, bufa.planes.view(-1)[: slices * ne * m * npad].view(slices, ne * m, npad),
                             bufa.scale[: ne * m])
            _lib.call("gpr_se_cross_planes", _p(as_e), _p(nrms_e), m, _p(st.a[e0:e1]), _p(st.nrm[e0:e1]), npad, _p(st.hp[e0:e1]),
                      st.hp.shape[-1], ne, n, st.dpad, slices, sig2, _p(pa.planes), _p(pa.scale), _stream())
            _count(1)
        else:
            _lib.call("gpr_se_cross_ld", _p(as_e), _p(nrms_e), m, _p(st.a[e0:e1]), _p(st.nrm[e0:e1]), npad, _p(st.hp[e0:e1]),
                      st.hp.shape[-1], ne, m, n, st.dpad, _p(Ks), npad, m * npad, _stream())
            _count(1)
            pa = ozaki_split(Ks[:ne].reshape(ne * m, npad), slices, bound=sig2, out=bufa)
        pb = cache.get((e0, e1)) if cache is not None else None
        if pb is None:
            _lib.call("gpr_tril_inverse_rows", _p(st.L[e0:e1]), _p(st.invdiag[e0:e1]), ne, npad, row0, _p(X), _stream())
            _count(1)
            keep = False
            if cache is not None and not cache["full"]:
                free, _ = torch.cuda.mem_get_info(dev)
                keep = free > 2 * slices * E * nl * npad + (PLANE_CACHE_RESERVE if bufb is None else PLANE_CACHE_RESERVE // 2)
                cache["full"] = not keep
            if keep:
                pb = cache[(e0, e1)] = ozaki_split(X[:ne].reshape(ne * nl, npad), slices)
            else:
                if bufb is None:
                    bufb = DigitPlanes(E * nl, npad, slices, _work("emu_pb" + tag, slices * E * nl * npad, dev), _f64_work("emu_sb" + tag, E * nl, dev))
                pb = ozaki_split(X[:ne].reshape(ne * nl, npad), slices, out=bufb)
        ozaki_gemm_nt(pa, pb, out=V[:ne], b_lower=row0, batch=ne)
        if per_expert_var0:                                   # independent experts: each has its own prior variance
            for i in range(ne):
                _lib.call("gpr_vt_reduce", _p(V[i]), nl, _p(st.z[e0 + i, row0:]), st.z.stride(0), _p(mean0), _p(var0[e0 + i]), 1, m, nl,
                          _p(mean[e0 + i]), _p(var[e0 + i]), _stream())
            _count(ne)
        else:
            _lib.call("gpr_vt_reduce", _p(V), nl, _p(st.z[e0:e1, row0:]), st.z.stride(0), _p(mean0), _p(var0), ne, m, nl, _p(mean[e0:e1]),
                      _p(var[e0:e1]), _stream())
            _count(1)
    return mean, var


def _pad_points(xs: torch.Tensor):
    """Test points in whole 128-row tiles: the last one is repeated (and its results dropped again by the caller)."""
    m = xs.shape[0]
    pad = round_up(m, TILE) - m
    return (torch.cat((xs, xs[-1:].expand(pad, -1))).contiguous() if pad else xs), m


def predict_emulated(st: FitState, xs: torch.Tensor, slices: int = 6, chunk: int = 8, tag: str = ""):
    """engine.predict (gpr.py:76-106) with the K* L^-T product of every expert on the INT8 tensor cores (opt-in):
    `slices` digit planes per operand (6 = FP64 grade, 4 = FP32 grade); the factorisation stays FP64."""
    assert can_emulate_plain(st)
    xp, m = _pad_points(xs)
    ensure_inverse_factors(None, st)
    zero = torch.zeros(xp.shape[0], dtype=F64, device=xs.device)
    prior = st.prior[:, None].expand(st.B, xp.shape[0]).contiguous()
    mean, var = _emulated_experts(st, xp, 0, zero, prior, slices, chunk, tag)
    return mean[:, :m].contiguous(), var[:, :m].contiguous()


def predict_shared_emulated(st_g: FitState, st_l: FitState, xs: torch.Tensor, slices: int = 6, chunk: int = 8):
    """predict_shared with the prediction products on the INT8 tensor cores (opt-in).

    For every local expert  V_l^T = K* X_l^T  with X_l = rows ng.. of L^-1 (lower trapezoidal) is ONE emulated-FP64 GEMM
    (gpr_ozaki_gemm_nt, `slices` digit planes per operand, exact int32 accumulation); mean and variance follow as in
    predict_shared (mean_g + V_l^T z_l, var_g - |V_l|^2).  The communication expert is predicted the same way (one triangular
    product with L_g^-1, at no fewer than 6 planes in the FP64-grade modes); the factorisations stay FP64.  Experts go
    through in chunks of `chunk` (transient K*, X_l, digit planes and V_l^T; the digit planes of X_l are kept in
    ``st_l.plane_cache`` when HBM allows).  Same return value as predict_pair."""
    assert can_emulate(st_g, st_l)
    xp, m = _pad_points(xs)
    slices_g = max(slices, 6) if slices >= 5 else slices     # the communication expert enters every expert's result
    mg, vg = predict_emulated(st_g, xp, slices_g, 1, tag="_g")
    ensure_inverse_factors(st_g, st_l)
    ml, vl = _emulated_experts(st_l, xp, st_g.n, mg[0], vg[0], slices, chunk)
    return (mg[:, :m].contiguous(), vg[:, :m].contiguous()), (ml[:, :m].contiguous(), vl[:, :m].contiguous())
