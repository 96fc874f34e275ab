"""Generalized Robust Bayesian Committee Machine with the reference's API (PyGPR/gr_bcm.py)."""
from __future__ import annotations

import torch

from . import dist as gdist
from . import engine
from .config import options
from .covar import _back, fused_spec
from .gpr import GPR, Exact_GP


class GRBCM(GPR):
    """GRBCM (arXiv:1806.00720): ``nc`` local experts, each trained on the communication set (xg, yg) plus its own
    block (gr_bcm.py:19-29), and one communication expert on (xg, yg) alone; predictions are combined with
    the precision weights of gr_bcm.py:116-149.

    Same constructor and attributes as the reference (``gpg, gpl, nc, nsc, ng, dim, beta, prec``).  Extra,
    optional keyword arguments drive expert sharding over GPUs (one process per GPU):

    * ``shard=True``      — ``xl, yl`` hold ALL experts on every rank; this rank keeps its contiguous block;
    * ``expert_offset=k, nc_total=N`` — ``xl, yl`` already hold only this rank's block, which starts at global
      expert ``k`` of ``N``;
    * ``group``           — the torch.distributed process group (default: world).

    ``predict`` then all-reduces the ``[3, m]`` partial sums (NCCL) and returns identical results on all ranks.
    """

    def __init__(self, xl, yl, xg, yg, cov, hp=None, **kargs):
        self.group = kargs.get("group", None)
        rank, ws = gdist.world(self.group)
        nc_given = xl.shape[0]
        if kargs.get("shard", False) and ws > 1:
            lo, hi = gdist.shard_bounds(nc_given, ws, rank)
            xl, yl = xl[lo:hi], yl[lo:hi]
            self.expert_offset, self.nc_total = lo, nc_given
            self.sharded = True
        else:
            self.expert_offset = int(kargs.get("expert_offset", 0))
            self.nc_total = int(kargs.get("nc_total", nc_given))
            # a model that holds every expert is NOT sharded, whatever torch.distributed's state: it enters no collective
            # (replicated models on several ranks would otherwise sum their identical partial sums world_size times)
            self.sharded = bool(kargs.get("sharded", self.nc_total != xl.shape[0] or self.expert_offset != 0)) and ws > 1

        nc = xl.shape[0]
        nls = xl.shape[1]
        ng = xg.shape[0]
        dim = xg.shape[1]

        # communication set FIRST, then the local block (gr_bcm.py:19-26)
        x = torch.cat((xg.to(xl.device)[None].expand(nc, ng, dim), xl), dim=1)
        y = torch.cat((yg.to(yl.device)[None].expand(nc, ng), yl), dim=1)

        self.gpg = Exact_GP(xg, yg, cov)                      # gr_bcm.py:28
        self.gpl = Exact_GP(x, y, cov)                        # gr_bcm.py:29

        self.nc = nc
        self.nsc = nls
        self.ng = ng
        self.dim = dim
        self._weights = None
        self._share_key, self._share_val = None, False
        self._beta = NotImplemented
        self._prec = NotImplemented

    def _weight(self, k):
        if self._weights is not None:
            first, rows = self._weights[k]
            return _back(torch.cat((first[None], rows)), self._weights[2])
        return NotImplemented

    @property
    def beta(self):
        if self._beta is None:
            self._beta = self._weight(0)
        return self._beta

    @beta.setter
    def beta(self, v):
        self._beta = v

    @property
    def prec(self):
        if self._prec is None:
            self._prec = self._weight(1)
        return self._prec

    @prec.setter
    def prec(self, v):
        self._prec = v

    # ---------------------------------------------------------------------------------------------
    def aggregate(self, ys_g, covars_g, ys_l, covars_l, var="diag"):
        """Precision-weighted combination (gr_bcm.py:116-149).  With sharding the three sums over the expert
        axis are formed per rank and all-reduced; ``beta``/``prec`` then hold row 0 + this rank's rows."""
        if var != "diag":
            return self._aggregate_full(ys_g, covars_g, ys_l, covars_l)
        mean_g, var_g = engine._dev(ys_g).reshape(-1), engine._dev(covars_g).reshape(-1)
        m = mean_g.shape[0]
        return self._aggregate_device(mean_g, var_g, engine._dev(ys_l).reshape(-1, m), engine._dev(covars_l).reshape(-1, m), ys_g)

    def _aggregate_device(self, mean_g, var_g, mean_l, var_l, out_like):
        """aggregate() on device tensors (mean_g, var_g [m]; mean_l, var_l [nc, m]): nothing but the two result vectors
        leaves the device (and those only when the caller's tensors live on the host)."""
        if mean_l.shape[0] == 0:                              # a rank without local experts (more ranks than experts)
            S = torch.zeros(3, mean_g.shape[0], dtype=mean_g.dtype, device=mean_g.device)
            brow = prow = torch.empty(0, mean_g.shape[0], dtype=mean_g.dtype, device=mean_g.device)
        else:
            S, brow, prow = engine.grbcm_partial(mean_l, var_l, var_g, self.expert_offset, want_rows=True)
        if self.sharded:
            gdist.all_reduce_sum_(S, self.group)              # the only data-path collective: [3, m] doubles
        mu, v, b0, p0 = engine.grbcm_finalize(S, mean_g, var_g, want_rows=True)
        # self.beta / self.prec (gr_bcm.py:135-136): kept on the device and moved to the callers' device on first read — a
        # [nc+1, m] pair that most callers never look at should not cost a device->host copy per prediction
        self._weights = ((b0, brow), (p0, prow), out_like)
        self._beta = self._prec = None
        return _back(mu, out_like), _back(v, out_like)

    def _aggregate_full(self, ys_g, covars_g, ys_l, covars_l):
        return engine.grbcm_aggregate_full(self, ys_g, covars_g, ys_l, covars_l)

    def update(self):
        """Fit both models (the reference does it lazily inside the two predict calls, gr_bcm.py:152-153 -> gpr.py:78).
        When both are stale the communication expert (one matrix: dozens of small dependent launches) is factorised on a
        high-priority side stream while the local experts' launch runs on the current one."""
        g, l = self.gpg, self.gpl
        if self._shares_hyperparameters():
            # one hyper-parameter vector for the whole committee: the local experts' factors start with the communication
            # expert's (gr_bcm.py:19-26 puts its training set first in every expert), so it is fitted first and reused
            emulated = options.emulate_fit_slices > 0 and engine.fit_emulated_pays(l.x.shape[0] if l.x.dim() == 3 else 1, l.x.shape[-2])
            if g.need_upd and options.fused_fit and not emulated:
                self._fused_update()                         # both factorisations in one launch
            else:
                if g.need_upd:
                    g.update()
                    l.need_upd = True
                if l.need_upd:
                    l._launch_update(shared=g._state)
                    l._finish_update()
            if l._state.shared_from is g._state:
                engine.tril_inverse(g._state)                # once per fit: the communication expert predicts in GEMM form
            return None
        if l._state is not None and l._state.shared_from is not None and not l.need_upd and g.need_upd:
            l.need_upd = True                                 # the local factors were built on a communication factor that is stale
        if not (g.need_upd and l.need_upd) or fused_spec(g.cov, self.dim) is None:
            g.update()
            l.update()
            return None
        dev = engine.device()
        main = torch.cuda.current_stream(dev)
        side = engine.side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            g._launch_update(ws_tag="_side")
        l._launch_update()
        main.wait_stream(side)
        g._state.record_stream(main)
        g._finish_update()
        l._finish_update()
        return None

    def _fused_update(self) -> None:
        """gpg.update() + gpl.update() (gr_bcm.py:152-153 -> gpr.py:65-74) as one launch (engine.fit_grbcm)."""
        g, l = self.gpg, self.gpl
        spec = fused_spec(g.cov, self.dim)
        xg, yg, hp_g = g._device_inputs()
        x, y, hp = l._device_inputs()
        old, l._state, g._state = l._state, None, None
        npad = engine.round_up(x.shape[1], engine.TILE)
        L_buf = old.L if old is not None and tuple(old.L.shape) == (x.shape[0], npad, npad) and old.L.device == x.device else None
        del old
        g._state, l._state = engine.fit_grbcm(xg, yg, hp_g, x, y, hp, spec[0], L_out=L_buf)
        g._finish_update()
        l._finish_update()

    def _shares_hyperparameters(self) -> bool:
        """True when every local expert's hyper-parameter row equals the communication expert's vector (the GRBCM paper's
        setting, and what GRBCM_MLE trains) and the communication set holds at least one 128-row block — the conditions of
        the shared leading block (engine.fit(shared=...)); a ragged rest of the communication set is each expert's own."""
        g, l = self.gpg, self.gpl
        if not options.share:                                        # A/B knob: treat the experts as unrelated matrices
            return False
        if fused_spec(g.cov, self.dim) is None or self.ng < engine.TILE or self.nsc < 1 or self.nc < 1:
            return False
        # the comparison reads every hyper-parameter row back: do it once per (params, params) pair, not once per predict
        key = (id(g.params), g.params._version, id(l.params), l.params._version)
        if self._share_key != key:
            pg = g.params.detach().reshape(-1)
            pl = l.params.detach().reshape(-1, l.params.shape[-1])
            self._share_val = bool(pl.shape[-1] == pg.shape[0] and (pl.to(pg.device) == pg).all())
            self._share_key = key
        return self._share_val

    def predict(self, xs, var="diag"):
        if var == "diag" and fused_spec(self.gpg.cov, self.dim) is not None:
            # gr_bcm.py:152-155 with the two predictions overlapped on the device (engine.predict_pair) and the per-expert
            # means / variances handed to the aggregation kernels where they are: [nc, m] intermediates never visit the host
            self.update()
            xd = engine._dev(xs).reshape(-1, xs.shape[-1])
            # test points go through in chunks: the per-chunk scratch (the communication expert's V_g^T is
            # [chunk, ng]) stays bounded however many points are asked for; the kernels see chunks as independent batches
            step = max(int(options.predict_chunk), 1)
            if self.nc == 0:
                mg, vg = engine.predict(self.gpg._state, xd, True)
                ml = vl = mg[:0]
            elif xd.shape[0] <= step:
                (mg, vg), (ml, vl) = engine.predict_pair(self.gpg._state, self.gpl._state, xd)
            else:
                parts = [engine.predict_pair(self.gpg._state, self.gpl._state, xd[s:s + step].contiguous())
                         for s in range(0, xd.shape[0], step)]
                mg, vg = (torch.cat([p[0][k] for p in parts], dim=1) for k in (0, 1))
                ml, vl = (torch.cat([p[1][k] for p in parts], dim=1) for k in (0, 1))
            return self._aggregate_device(mg[0], vg[0], ml, vl, self.gpg.x)   # gr_bcm.py:155
        ys_g, covars_g = self.gpg.predict(xs, var=var)        # gr_bcm.py:152
        ys_l, covars_l = self.gpl.predict(xs, var=var)        # gr_bcm.py:153
        if self.nc == 1:
            ys_l = ys_l.reshape(1, -1)
        return self.aggregate(ys_g, covars_g, ys_l, covars_l, var=var)  # gr_bcm.py:155

    # ---- hyper-parameter training --------------------------------------------------------------
    def train(self, method="CG", jac=True):
        """Per-expert maximum-likelihood training: one scipy minimisation per local expert, then the communication expert.

        The reference's ``train`` / ``cost_fun_*`` (gr_bcm.py:36-97) call names that do not exist and cannot run; this is
        what they were written to do, on the working ``MLE``.  Results land in ``gpl.params[cen]`` / ``gpg.params``, the final
        NLML and gradient in ``gpl.llhd, gpl.jac_llhd, gpg.llhd, gpg.jac_llhd``; returns the communication expert's result.
        """
        import scipy.optimize as opt
        from .loss import MLE
        gpl, gpg = self.gpl, self.gpg
        hp = torch.clone(gpl.params).reshape(self.nc, -1)
        gpl.llhd = torch.empty(self.nc)
        gpl.jac_llhd = torch.empty_like(hp)

        def fit_one(model, start):
            loss = MLE(model)
            fun = loss.loss_and_grad if jac else loss.loss
            res = opt.minimize(fun, start.detach().cpu().numpy(), jac=True if jac else None, method=method)
            fin_l, fin_j = loss.loss_and_grad(res.x)
            return res, float(fin_l), torch.from_numpy(fin_j.copy())

        for cen in range(self.nc):
            res, l, j = fit_one(Exact_GP(gpl.x[cen], gpl.y[cen], gpl.cov), hp[cen])
            hp[cen] = torch.from_numpy(res.x)
            gpl.llhd[cen] = l
            gpl.jac_llhd[cen] = j
        gpl.set_params(hp)
        res, l, j = fit_one(Exact_GP(gpg.x, gpg.y, gpg.cov), gpg.params)
        gpg.set_params(torch.from_numpy(res.x))
        gpg.llhd, gpg.jac_llhd = torch.tensor(l), j
        return res

    def train_batched(self, steps: int = 10, eps: float = 1e-4):
        """B200-native training loop: ALL local experts (and the communication expert) take a quadratic-fit gradient
        step per iteration, ``theta <- theta - gamma J`` with ``gamma`` from the parabola of ``get_learn_rate`` (hp_update.py:14-22),
        i.e. one batched NLML + gradient and two batched NLML evaluations per step for the whole committee.

        The line search runs on the device: the iterates, the three NLML vectors, slope / curvature / gamma and the masked update
        are device tensors, the factorisations report failure through their ``info`` vector (no ``check()`` read-back), and nothing
        crosses PCIe inside the loop — the NLML history comes back once at the end.  Experts whose parabola is not convex
        (gamma <= 0 or not finite) or whose kernel matrix is not positive definite at a trial point stay where they are.
        Per-expert hyper-parameters need no collective; returns the NLML history [steps, nc] (numpy)."""
        import numpy as np
        from .loss import MLE
        hist = []
        for model in (self.gpl, self.gpg):
            loss = MLE(model)
            shape = tuple(model.params.shape)
            hp = engine._dev(model.params.detach()).to(torch.float64).reshape(-1, shape[-1]).clone()
            for _ in range(steps):
                st = loss._fit(hp, True, check=False)                  # the three evaluations of hp_update.py:14-17
                f0, good = st.nlml(), st.info == 0
                J = loss._grad(st)
                stm = loss._fit(hp - eps * J, False, check=False)
                f_minus, good = stm.nlml(), good & (stm.info == 0)
                stp = loss._fit(hp + eps * J, False, check=False)
                f_plus, good = stp.nlml(), good & (stp.info == 0)
                slope = (f_minus - f_plus) / (2.0 * eps)               # C1, hp_update.py:19
                curv = (f_minus + f_plus - 2.0 * f0) / (2.0 * eps ** 2)  # C2, hp_update.py:20
                gamma = -0.5 * (slope / curv)                          # hp_update.py:22
                step = gamma[:, None] * J
                ok = good & torch.isfinite(step).all(-1) & (gamma > 0)
                hp = hp - torch.where(ok[:, None], step, torch.zeros_like(step))
                if model is self.gpl:
                    hist.append(f0)
            model.set_params(_back(hp.reshape(shape), model.params))
        return torch.stack(hist).cpu().numpy() if hist else np.zeros((0, self.nc))

    def gather_weights(self):
        """(beta, prec) for ALL experts [nc_total+1, m] when sharded (all-gather of the local rows)."""
        if not self.sharded:
            return self.beta, self.prec
        _, ws = gdist.world(self.group)
        counts = [gdist.shard_bounds(self.nc_total, ws, r) for r in range(ws)]
        counts = [hi - lo for lo, hi in counts]
        out = []
        for t in (self.beta, self.prec):
            td = engine._dev(t)
            out.append(_back(torch.cat((td[:1], gdist.all_gather_rows(td[1:].contiguous(), counts, self.group))), t))
        return tuple(out)


class GRBCM_MLE:
    """Summed negative log marginal likelihood of a GRBCM at ONE shared hyper-parameter vector (BASELINE
    config 5: "GRBCM hyperparameter training loop (summed LML gradient allreduce)").

        loss(theta) = NLML_g(theta) + sum_i NLML_i(theta)         theta: ndarray [P]

    Same duck type as ``Loss`` (``loss / grad / loss_and_grad`` on numpy, ``loss_value / grad_value``, ``model`` with
    ``params / set_params``), so ``CG``, ``CG_Quad``, ``BFGS_Quad``, ``Nelder_Mead`` and ``get_learn_rate`` drive it unchanged.
    Local experts are evaluated in batches of ``chunk`` to bound HBM (0 = as many as fit in the free device memory, each
    expert needing one [npad, npad] factor); with experts sharded over ranks the ``[1+P]`` partial sums are all-reduced (NCCL),
    the replicated communication expert is added once.
    """

    def __init__(self, model: "GRBCM", chunk: int = 0) -> None:
        self.grbcm = model
        self.chunk = chunk
        self.loss_value = NotImplemented
        self.grad_value = NotImplemented
        self.params = torch.clone(model.gpg.params)

    @property
    def model(self):
        """Optimisers call ``loss.model.params / set_params``: the loss is its own model.  (A property, not ``self.model = self``:
        the reference cycle kept a dropped instance — and its HBM-sized factor buffer — alive until the next cyclic GC pass.)"""
        return self

    def set_params(self, params) -> None:
        """Write the shared vector into every expert and mark the factorisations stale."""
        self.params = torch.clone(params)
        g = self.grbcm
        g.gpg.set_params(params)
        g.gpl.set_params(params.reshape(1, -1).repeat(g.nc, 1))

    def _eval(self, theta, want_grad: bool):
        import numpy as np
        from .loss import MLE
        g = self.grbcm
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        P = theta.shape[-1]
        acc = np.zeros(1 + P)
        step = self.chunk
        if step <= 0:
            # sized once per instance and mode: a later evaluation sees less free memory (the staging buffers of the first one are
            # cached) and would pick a different, smaller batch
            key = (options.emulate_grad_slices > 0, options.emulate_fit_slices > 0)
            if getattr(self, "_auto", (None, 0))[0] != key:
                self._auto = (key, self._auto_chunk())
            step = self._auto[1]
        # the communication expert at theta: its NLML/gradient is added once below, and (one shared vector, tile-aligned
        # communication set) its factor is the leading block of every local expert's factor (engine.fit(shared=...))
        glob = getattr(self, "_glob", None)
        if glob is None:
            glob = self._glob = MLE(g.gpg)
        st_g = glob._fit(theta, want_alpha=True)
        share = st_g if (isinstance(st_g, engine.FitState) and g.ng >= engine.TILE and options.share) else None
        if share is not None and want_grad and g.ng % engine.TILE == 0:
            engine.kinv(st_g)      # K_g^-1 and L_g^-T once: the experts' gradients reuse both (gpr_lml_grad_shared)
        for s in range(0, g.nc, step):
            e = min(s + step, g.nc)
            part = getattr(self, "_part", None)
            if part is None:
                part = self._part = MLE(Exact_GP(g.gpl.x[s:e], g.gpl.y[s:e], g.gpl.cov))
            else:                                            # same instance, next batch of experts: its factor buffer is reused
                part.model.x, part.model.y = g.gpl.x[s:e], g.gpl.y[s:e]
            part.shared_state = share
            th = np.tile(theta, (e - s, 1))
            if want_grad:
                l, j = part.loss_and_grad(th)
                acc[0] += np.sum(l)
                acc[1:] += np.sum(np.reshape(j, (e - s, P)), axis=0)
            else:
                acc[0] += np.sum(part.loss(th))
        t = torch.from_numpy(acc)
        if g.sharded:
            t = engine._dev(t)
            gdist.all_reduce_sum_(t, g.group)         # [1+P] doubles
            acc = t.cpu().numpy()
        l = float(st_g.nlml()[0])
        if want_grad:
            j = glob._grad(st_g)[0].cpu().numpy()           # last: it overwrites the block-upper part of the shared factor
            return float(acc[0] + l), acc[1:] + j
        return float(acc[0] + l), None

    def _auto_chunk(self) -> int:
        """Experts per batch when the caller gave none: what fits in the free device memory (one factor per expert plus the
        inputs and workspaces, a few GB held back), all of them if possible.  The previous batch's factor buffer is reused."""
        g = self.grbcm
        if g.nc <= 1:
            return 1
        npad = engine.round_up(g.ng + g.nsc, engine.TILE)
        per = npad * npad * 8 + 3 * npad * engine.TILE * 8
        free, _ = torch.cuda.mem_get_info(engine.device())
        held = self._part._L_buf.numel() * 8 if getattr(self, "_part", None) is not None and self._part._L_buf is not None else 0
        reserve = 6 << 30
        # the emulated forms stage FP64 windows and digit planes for a chunk of experts at a time (engine.lml_grad_emulated: rows of
        # U, their planes, K^-1; engine.fit_emulated: two sets of planes of the widest [below x panel] window): leave room for full
        # chunks — measured at config 5, 64 experts in one batch with 5-expert staging 3.87 s, 10 per batch 3.48 s
        if options.emulate_grad_slices > 0:
            reserve += options.emulate_chunk * npad * npad * (16 + options.emulate_grad_slices)
        if options.emulate_fit_slices > 0:
            ngt = (g.ng // engine.TILE) * engine.TILE if options.share else 0
            reserve += min(options.emulate_fit_chunk, g.nc) * 2 * options.emulate_fit_slices * (npad - ngt) * max(ngt, 1024)
        fit = int(max(1, min(g.nc, (free + held - reserve) // per)))
        if options.emulate_grad_slices > 0 or options.emulate_fit_slices > 0:
            # tools/train_chunk_sweep.py (config 5's share, 64 experts of order 16384): 8 / 12 / 16 / 24 / 32 / 48 experts per batch take
            # 3.71 / 3.55 / 3.47 / 3.54 / 3.91 / 3.88 s — beyond ~24 the staging of the emulated products no longer gets whole chunks
            fit = min(fit, 2 * max(options.emulate_chunk, 4))
        return fit

    def loss(self, params):
        self.loss_value, _ = self._eval(params, False)
        return self.loss_value

    def grad(self, params):
        _, self.grad_value = self._eval(params, True)
        return self.grad_value

    def loss_and_grad(self, params):
        self.loss_value, self.grad_value = self._eval(params, True)
        return (self.loss_value, self.grad_value)


def log_likelihood_batched(x, y, hp, cov, **kwargs):
    """Batched NLML of experts (x:[nc,n,d], y:[nc,n], hp:[nc,P]) -> numpy [nc].

    The reference's function of this name (gr_bcm.py:158-176) cannot run (it calls ``cov`` as a function and
    drops two of the three terms); this is the quantity it was meant to return, computed like MLE.loss.
    """
    from .loss import MLE
    model = Exact_GP(x, y, cov)
    return MLE(model).loss(hp.detach().cpu().numpy() if isinstance(hp, torch.Tensor) else hp)
