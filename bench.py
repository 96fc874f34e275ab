#!/usr/bin/env python
"""Benchmark of the PyGPR hot path on B200: GRBCM predictions/sec at n=1M, d=16 (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--test-points M] [--small]

Workload (reading A): 256 subsets x 4096 points of a 16-d synthetic regression problem; one subset
is the communication set, so there are 255 local experts of order 8192 plus one communication expert of order
4096.  A *step* is one GRBCM.predict of a batch of M test points with the factorisations resident (exactly what
the reference's need_upd flag gives on every predict after the first).  The fit (batched Cholesky) and one
LML-gradient evaluation are timed once before the steps and reported in the same JSON line as FP64 TFLOP/s.

  value : whole-job predictions/sec, inputs resident in HBM, CUDA events, max over ranks
  e2e   : the same through the public API (pygpr_b200.GRBCM.predict) with HOST tensors in and out
          (pinned test points H2D + mean/variance D2H inside the timed region)
  roofline : the dominant kernel (predict_kernel) against the FP64 GEMM peak measured in this run with cuBLAS
  cpu_baseline : the CPU oracle (torch-CPU restatement of the reference) on a bounded sample, host cores stated

With N>1 (torchrun, one rank per GPU) the 255 local experts are split into contiguous blocks and the only
collective is the all-reduce of the [3, M] partial sums (strong scaling: the total work is fixed).
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GRBCM predictions/sec at n=1M,d=16"
UNIT = "predictions/s"


def workload(small: bool):
    if small:  # smoke-sized variant for quick checks (not the headline)
        return dict(name="grbcm_small_64x1024_d8(A)", M_sub=16, m0=1024, d=8)
    return dict(name="grbcm_n1M_d16_256x4096(A): xg[4096,16], xl[255,4096,16], 255 local experts of order 8192 + comm expert",
                M_sub=256, m0=4096, d=16)


def make_data(w, seed=0):
    """x ~ U[0,1]^d, y = sin(sum x) + 0.1 N(0,1); subset 0 is the communication set."""
    g = torch.Generator().manual_seed(seed)
    M_sub, m0, d = w["M_sub"], w["m0"], w["d"]
    x = torch.rand(M_sub, m0, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(M_sub, m0, generator=g, dtype=torch.float64)
    return x[1:], y[1:], x[0], y[0]


def make_hp(B, d):
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0
    hp[:, 1:d + 1] = 2.0 / math.sqrt(d)
    hp[:, d + 1] = 0.3
    return hp


def test_points(m, d, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(m, d, generator=g, dtype=torch.float64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], text=True, timeout=5)
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.02)   # back-to-back queries (~50 ms each): the timed region is only ~0.6 s at 8 GPUs

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measure_dgemm_peak(dev):
    """FP64 GEMM peak of THIS device, cuBLAS 8192^3 NT, best of 8 (MEASURED_PEAKS.json has no FP64 entry)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b.t(), out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b.t(), out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best * 1e-9


L2_NOTE = ("no flush: every step streams this rank's expert factors (137 GB at config 4 on one GPU, 17 GB per GPU on eight: far "
           "beyond the 126 MB L2), so no input of a step survives in L2 from the previous one")


def config_of(w, m):
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": w["name"], "m_per_step": m, "l2": L2_NOTE}


def cpu_sample(w, m_s, nc_s):
    """Bounded CPU sample of the same workload: the first nc_s local experts + the communication expert, m_s test points,
    through the UNMODIFIED reference when a copy is importable (baseline/_ref; kind "reference"), else the oracle port."""
    import warnings
    from oracle import cpu_arm
    warnings.simplefilter("ignore")
    xl, yl, xg, yg = make_data(w)
    d = w["d"]
    cs = cpu_arm.CpuSample(xl[:nc_s].clone(), yl[:nc_s].clone(), xg, yg, make_hp(nc_s, d), make_hp(1, d)[0])
    cs.fit()
    return cs


def scale_to_job(w, nc_s):
    """Local experts are equal-sized and independent: the whole job costs the sample times the N^2 work ratio."""
    nc = w["M_sub"] - 1
    n_l, n_g = 2 * w["m0"], w["m0"]
    return (nc * n_l ** 2 + n_g ** 2) / float(nc_s * n_l ** 2 + n_g ** 2)


def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (the unmodified package from
    baseline/_ref driven through its public API; the oracle port only if no copy is present), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = workload(args.small)
    nc = w["M_sub"] - 1
    nc_s, m_s = (8, args.m) if not args.small else (4, min(args.m, 256))
    torch.set_num_threads(os.cpu_count() or 1)
    cs = cpu_sample(w, m_s, nc_s)
    xs = test_points(m_s, w["d"])
    for _ in range(args.warmup):
        cs.predict(xs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cs.predict(xs)
    dt = (time.perf_counter() - t0) / args.steps
    ratio = scale_to_job(w, nc_s)
    value = m_s / (dt * ratio)
    sample = ("%s: %d of %d local experts (chunks of 2) + communication expert, %d test points per step, factors resident; "
              "predictions/s scaled by the N^2 work ratio %.2f to the whole job" %
              ("unmodified reference (baseline/_ref) through GRBCM / Exact_GP.predict / aggregate" if cs.kind == "reference"
               else "oracle port (no reference copy present)", nc_s, nc, m_s, ratio))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(w, m_s),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cs.kind, "sample": sample,
                             "fit_s_sample": cs.fit_s},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    out.write(json.dumps(line) + "\n")
    out.flush()
    return 0


def train_workload(small: bool):
    """BASELINE config 5 (reading A): 512 subsets x 8192 points, d = 32 -> local experts of order 16384 + a communication expert of
    order 8192, P = 34 hyper-parameters, ONE shared vector; 64 local experts per rank (the 8-GPU share of the 511)."""
    if small:
        return dict(name="grbcm_train_small_d8: 6 experts/rank of order 1024 + comm 512", per_rank=6, m0=512, d=8)
    return dict(name="grbcm_train_n4M_d32_512x8192(A): 64 local experts per rank of order 16384 + comm expert of order 8192, P=34",
                per_rank=64, m0=8192, d=32)


def run_train(args, P, eng, world, rank, dev, barrier, max_over_ranks, peak_tf):
    """One evaluation of the summed NLML and its gradient at one shared hyper-parameter vector through the PUBLIC training API
    (GRBCM_MLE.loss_and_grad -> MLE -> gpr_potrf_batched_shared + gpr_lml_grad_shared, experts in HBM-sized chunks, [1+P]
    all-reduce over the ranks): the inner step of config 5's training loop (loss.py:92-128, hp_update.py:6-28)."""
    tw = train_workload(args.small)
    per, m0, d = tw["per_rank"], tw["m0"], tw["d"]
    g = torch.Generator().manual_seed(100 + rank)
    xl = torch.rand(per, m0, d, generator=g, dtype=torch.float64)
    yl = torch.sin(xl.sum(-1)) + 0.1 * torch.randn(per, m0, generator=g, dtype=torch.float64)
    g0 = torch.Generator().manual_seed(99)                   # the communication set is the same on every rank
    xg = torch.rand(m0, d, generator=g0, dtype=torch.float64)
    yg = torch.sin(xg.sum(-1)) + 0.1 * torch.randn(m0, generator=g0, dtype=torch.float64)
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    model = P.GRBCM(xl.to(dev), yl.to(dev), xg.to(dev), yg.to(dev), cov, expert_offset=rank * per, nc_total=per * world)
    theta = make_hp(1, d)[0]
    mle = P.GRBCM_MLE(model)                                  # chunk = what fits in the free HBM
    mle.set_params(theta)
    chunk = mle._auto_chunk()
    mle.loss(theta.numpy())                                   # warm-up: allocations, kernel loading
    barrier(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); f_loss = mle.loss(theta.numpy()); e1.record(); barrier()
    loss_ms = max_over_ranks(e0.elapsed_time(e1))
    barrier(); e0.record(); f, grad = mle.loss_and_grad(theta.numpy()); e1.record(); barrier()
    lg_ms = max_over_ranks(e0.elapsed_time(e1))
    assert math.isfinite(f) and bool(torch.isfinite(torch.from_numpy(grad)).all()) and abs(f - f_loss) <= 1e-9 * abs(f)
    from pygpr_b200 import config
    del mle
    gc.collect()
    torch.cuda.empty_cache()                                  # the FP64 evaluation's factor buffer goes back before the next one allocates
    config.options.emulate_grad_slices = 7                    # the same evaluation with L^-T's off-segment blocks and K^-1 on the
    config.options.emulate_inverse_slices = 7                 # INT8 tensor cores (opt-in), and so are the Cholesky's panel solves and
    config.options.emulate_fit_slices = 7                     # trailing updates (engine.fit_emulated)
    try:
        mle_e = P.GRBCM_MLE(model)
        mle_e.set_params(theta)
        chunk_e = mle_e._auto_chunk()
        mle_e.loss_and_grad(theta.numpy())                    # warm-up: staging buffers, kernel loading (like the FP64 evaluation above)
        barrier(); e0.record(); f_e, grad_e = mle_e.loss_and_grad(theta.numpy()); e1.record(); barrier()
        lg_e_ms = max_over_ranks(e0.elapsed_time(e1))
    finally:
        config.options.emulate_grad_slices = 0
        config.options.emulate_inverse_slices = 0
        config.options.emulate_fit_slices = 0
    emulated = {"slices": 7, "loss_and_grad_ms": lg_e_ms, "speedup_vs_fp64_path": lg_ms / lg_e_ms, "chunk_experts": chunk_e,
                "nlml_vs_fp64_path": abs(f_e - f) / abs(f), "grad_vs_fp64_path": float(np.abs(grad_e - grad).max() / np.abs(grad).max())}
    del mle_e
    n_l, n_g = 2.0 * m0, float(m0)
    nc_total = per * world
    # executed flops: Cholesky (n^3 - ng^3)/3 + L^-T and K^-1/gradient 2 (n^3 - ng^3)/3 per local expert (shared leading block),
    # + the replicated communication expert's n_g^3 on every rank; SURVEY 8(d)'s count N^3 per expert beside it
    flops = nc_total * (n_l ** 3 - n_g ** 3) + world * n_g ** 3
    flops_unshared = nc_total * n_l ** 3 + world * n_g ** 3
    out = {"workload": tw["name"], "experts_per_rank": per, "experts_total": nc_total, "chunk_experts": chunk, "P": d + 2,
           "loss_and_grad_ms": lg_ms, "loss_ms": loss_ms, "tflops": flops / lg_ms * 1e-9, "frac_of_peak": flops / lg_ms * 1e-9 / (peak_tf * world),
           "unshared_count_tflops": flops_unshared / lg_ms * 1e-9, "nlml": f, "grad_norm": float((grad * grad).sum() ** 0.5),
           "collective": "all_reduce(sum, f64) of [1+P] = %d doubles per evaluation over %d rank(s)" % (1 + d + 2, world),
           "emulated_fp64_grad": emulated,
           "what": "GRBCM_MLE.loss_and_grad at one shared hyper-parameter vector, wall time of the public call incl. host->device "
                   "staging of each expert chunk and the [1+P] read-back"}
    del model
    torch.cuda.empty_cache()
    return out


def train_parity(P, small):
    """Config-5 shape at reduced N against the CPU arm (the unmodified reference when present): GRBCM_MLE's summed NLML and
    gradient for 2 local experts + communication expert, d = 32 (d = 8 for --small)."""
    from oracle import cpu_arm
    d, m0 = (8, 256) if small else (32, 1024)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(3, m0, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(3, m0, generator=g, dtype=torch.float64)
    theta = make_hp(1, d)[0]
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    model = P.GRBCM(x[1:], y[1:], x[0], y[0], cov)
    f, grad = P.GRBCM_MLE(model, chunk=1).loss_and_grad(theta.numpy())
    cs = cpu_arm.CpuSample(x[1:], y[1:], x[0], y[0], theta.repeat(2, 1), theta)
    f_ref, g_ref = cs.loss_and_grad_sum()
    return {"vs": cs.kind, "experts": 2, "order": 2 * m0, "d": d, "nlml": abs(f - f_ref) / abs(f_ref), "grad": cpu_arm.rel_err(grad, g_ref)}


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on communicator creation,
    whatever NCCL_DEBUG_FILE says on some hosts), so file descriptor 1 is pointed at stderr for the whole run and the line goes
    to a private duplicate of the original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--test-points", dest="m", type=int, default=4096, help="test points per step")
    ap.add_argument("--small", action="store_true", help="smoke-sized workload (not the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-grad", action="store_true", help="skip the LML-gradient timing")
    ap.add_argument("--no-train", action="store_true", help="skip the config-5 training evaluation (GRBCM_MLE.loss_and_grad)")
    ap.add_argument("--emulate", type=int, default=6, help="also time the opt-in emulated-FP64 prediction (INT8 tcgen05 tensor cores) "
                                                            "with this many digit planes; 0 = skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, out)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL writes its version banner (and anything NCCL_DEBUG asks for) to stdout
        # unless it is given a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/pygpr_b200_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)

    import pygpr_b200 as P
    from pygpr_b200 import dist as gdist
    from pygpr_b200 import engine as eng

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    w = workload(args.small)
    d, m = w["d"], args.m
    xl, yl, xg, yg = make_data(w)
    nc_total = xl.shape[0]
    lo, hi = gdist.shard_bounds(nc_total, world, rank)
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    model = P.GRBCM(xl[lo:hi].contiguous(), yl[lo:hi].contiguous(), xg, yg, cov, expert_offset=lo, nc_total=nc_total)
    model.gpl.set_params(make_hp(hi - lo, d))
    model.gpg.set_params(make_hp(1, d)[0])
    del xl, yl
    peak_tf = measure_dgemm_peak(dev)

    # ---- fit: batched Cholesky (+ fused forward solve, backward solve), timed once after a warm-up pass -------------
    n_l, n_g, nc_loc = 2 * w["m0"], w["m0"], hi - lo
    model.update(); barrier()
    model.gpl.need_upd = True; model.gpg.need_upd = True   # the refit reuses the factor buffers (137 GB at N=1): no cudaMalloc in the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record(); model.update(); e1.record(); barrier()
    fit_ms = max_over_ranks(e0.elapsed_time(e1))
    st = model.gpl._state
    # One hyper-parameter vector for the committee (the GRBCM paper's setting and this workload's): every local expert's
    # factor starts with the communication expert's 4096 x 4096 factor, which is computed once and reused
    # (gpr_potrf_batched_shared / gpr_predict_shared).  The flop counts below are those of that algorithm; the counts of
    # SURVEY 8(d), which treat the experts as unrelated matrices, are reported next to them as "*_unshared_count".
    shared = st.shared_from is not None
    fit_flops_unshared = nc_total * n_l ** 3 / 3.0 + world * n_g ** 3 / 3.0  # the comm expert is replicated on every rank
    fit_flops = nc_total * (n_l ** 3 - n_g ** 3) / 3.0 + world * n_g ** 3 / 3.0 if shared else fit_flops_unshared

    # ---- steps: GRBCM.predict on m test points, factors resident -----------------------------------------------------
    xs_host = test_points(m, d).pin_memory()
    xs_dev = xs_host.to(dev)

    pk_ms = []
    last = {}

    def step_device():
        # communication expert on a side stream, local experts on this one (engine.predict_pair); k0..k1 brackets the
        # local experts' persistent kernel, which shares the SMs with the communication expert's ceil(m/128) tasks
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        (mg, vg), (ml, vl) = eng.predict_pair(model.gpg._state, st, xs_dev)
        k1.record()
        pk_ms.append((k0, k1))
        last.update(mg=mg, vg=vg, ml=ml, vl=vl)
        S, _, _ = eng.grbcm_partial(ml, vl, vg[0], lo)
        gdist.all_reduce_sum_(S)
        return eng.grbcm_finalize(S, mg[0], vg[0])[:2]

    for _ in range(args.warmup):
        step_device()
    pk_ms.clear()
    launches0 = eng.LAUNCHES
    gpu_index = int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[local_rank]) if os.environ.get("CUDA_VISIBLE_DEVICES") else local_rank
    with ClockSampler(gpu_index) as cs:
        barrier(); e0.record()
        for _ in range(args.steps):
            mu, var = step_device()
        e1.record(); barrier()
    step_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = eng.LAUNCHES - launches0
    kern_ms = sum(a.elapsed_time(b) for a, b in pk_ms) / len(pk_ms)
    kern_ms = max_over_ranks(kern_ms)
    value = m / (step_ms * 1e-3)
    clocks = cs.summary()
    ml_fp64, vl_fp64, mg_fp64, vg_fp64 = last["ml"], last["vl"], last["mg"], last["vg"]

    # ---- e2e: public API, host tensors in and out -------------------------------------------------------------------
    for _ in range(2):
        model.predict(xs_host, var="diag")
    barrier(); t0 = time.perf_counter()
    for _ in range(args.steps):
        mu_h, var_h = model.predict(xs_host, var="diag")
    barrier(); e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = max_over_ranks(e2e_ms)
    assert not mu_h.is_cuda and bool(torch.isfinite(mu_h).all()) and bool((var_h > 0).all())

    # ---- LML gradient (trtri + fused lauum/gradient) on the resident factors ---------------------------------------------
    grad_ms = None
    if not args.no_grad:
        eng.lml_grad(st); barrier()
        barrier(); e0.record(); g = eng.lml_grad(st); e1.record(); barrier()
        grad_ms = max_over_ranks(e0.elapsed_time(e1))
        assert bool(torch.isfinite(g).all())

    # ---- opt-in emulated-FP64 prediction: the same step with the local experts' products on the INT8 tcgen05 tensor cores -------
    emu = None
    if args.emulate > 0 and shared and eng.can_emulate(model.gpg._state, st):
        from pygpr_b200 import config
        config.options.emulate_slices = args.emulate
        config.options.emulate_inverse_slices = 7            # L^-T assembled from segment inverses by emulated products as well
        barrier(); e0.record(); eng.ensure_inverse_factors(model.gpg._state, st); e1.record(); barrier()
        prep_ms = max_over_ranks(e0.elapsed_time(e1))        # once per fit: L^-T of every expert (in place)
        for _ in range(2):
            step_device()
        barrier(); e0.record()
        for _ in range(args.steps):
            mu_e, var_e = step_device()
        e1.record(); barrier()
        emu_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        # the same through the public API with host tensors (pinned H2D of the test points, D2H of mean / variance in the timed region)
        model.predict(xs_host, var="diag")
        barrier(); t0 = time.perf_counter()
        for _ in range(args.steps):
            model.predict(xs_host, var="diag")
        barrier(); emu_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
        config.options.emulate_slices = 0
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        emu = {"slices": args.emulate, "plane_products": args.emulate * (args.emulate + 1) // 2, "value": m / (emu_ms * 1e-3), "unit": UNIT,
               "ms_per_step": emu_ms, "speedup_vs_fp64_path": step_ms / emu_ms, "prepare_ms_once_per_fit": prep_ms,
               "inverse_row_planes_cached_chunks": [len([k for k in (getattr(st, "plane_cache", None) or {}) if isinstance(k, tuple)]),
                                                    -(-nc_loc // config.options.emulate_chunk)],
               "hbm_free_gb": torch.cuda.mem_get_info(dev)[0] / 1e9,
               "e2e": {"value": m / (emu_e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": emu_e2e_ms, "h2d_bytes_per_step": m * d * 8,
                       "d2h_bytes_per_step": 2 * m * 8},
               "tflops_equiv": (nc_loc * (float(n_l) ** 2 - float(n_g) ** 2) + float(n_g) ** 2) * m / emu_ms * 1e-9,
               # INT8 tensor work of a step (plane products x 2 MNK of every expert's product) over the WHOLE step — digit-plane
               # preparation, reductions and aggregation included — against the INT8 UMMA rate measured on this pool
               # (tools/microbench/umma_i8_rate.cu, profiles/r02_umma_i8_rate.jsonl: 4518 TOP/s at N >= 128)
               "roofline": {"bound": "tensor (int8 tcgen05)", "unit": "TOP/s", "peak": 4518.0,
                            "achieved": args.emulate * (args.emulate + 1) // 2 * (nc_loc * (float(n_l) ** 2 - float(n_g) ** 2) + float(n_g) ** 2) * m / emu_ms * 1e-9,
                            "frac": args.emulate * (args.emulate + 1) // 2 * (nc_loc * (float(n_l) ** 2 - float(n_g) ** 2) + float(n_g) ** 2) * m / emu_ms * 1e-9 / 4518.0,
                            "peak_source": "measured: 128 x N x 32 kind::i8 UMMA issue rate on all 148 SMs, N = 128 / 256 (profiles/r02_umma_i8_rate.jsonl)"},
               "aggregate_mean_vs_fp64_path": rel(mu_e, mu), "aggregate_var_vs_fp64_path": rel(var_e, var),
               "expert_mean_vs_fp64_path": rel(last["ml"], ml_fp64), "expert_var_vs_fp64_path": rel(last["vl"], vl_fp64),
               "what": "GRBCM.predict with options.emulate_slices: per expert V_l^T = K* X_l^T (X_l = rows of L^-1) as one Ozaki-split "
                       "GEMM on tcgen05.mma kind::i8 (TMA operand loads, TMEM int32 accumulators, FP64 recombination); opt-in, "
                       "the headline `value` is the FP64 DMMA path"}
        emu_last = dict(last)
        last.update(ml=ml_fp64, vl=vl_fp64, mg=mg_fp64, vg=vg_fp64)
        if grad_ms is not None:
            # the NLML gradient with K^-1 = L^-T L^-1 on the INT8 tensor cores (7 digit planes; L^-T stays the FP64 blocked trtri)
            config.options.emulate_grad_slices = 7
            eng.lml_grad(st); barrier()
            st.upper_is_inverse = False                         # like lml_grad of the FP64 path, the timed call recomputes L^-T
            barrier(); e0.record(); g_e = eng.lml_grad(st); e1.record(); barrier()
            eg_ms = max_over_ranks(e0.elapsed_time(e1))
            barrier(); e0.record(); eng.lml_grad(st); e1.record(); barrier()
            eg2_ms = max_over_ranks(e0.elapsed_time(e1))       # L^-T already in place (e.g. after an emulated prediction of the same fit)
            config.options.emulate_grad_slices = 0
            config.options.emulate_inverse_slices = 0
            emu["lml_grad"] = {"slices": 7, "ms": eg_ms, "speedup_vs_fp64_path": grad_ms / eg_ms, "ms_with_inverse_in_place": eg2_ms,
                               "grad_vs_fp64_path": float((g_e - g).abs().max() / g.abs().max()),
                               "what": "engine.lml_grad with options.emulate_grad_slices / emulate_inverse_slices = 7: L^-T from FP64 inverses of "
                                       "diagonal segments + emulated-FP64 products between them, K^-1 = U U^T as one batched emulated-FP64 "
                                       "GEMM per chunk of experts, then the fused gradient reduction on K^-1"}
        config.options.emulate_inverse_slices = 0
        # the factorisation itself: right-looking blocked Cholesky with emulated-FP64 panel solves / trailing updates (7 planes)
        config.options.emulate_fit_slices = 7
        for _ in range(2):
            model.gpl.need_upd = True; model.gpg.need_upd = True
            barrier(); e0.record(); model.update(); e1.record(); barrier()
        config.options.emulate_fit_slices = 0
        ef_ms = max_over_ranks(e0.elapsed_time(e1))
        st_e = model.gpl._state
        emu["fit"] = {"slices": 7, "ms": ef_ms, "speedup_vs_fp64_path": fit_ms / ef_ms, "tflops_equiv": fit_flops / ef_ms * 1e-9,
                      "nlml_vs_fp64_path": float((st_e.nlml() - st.nlml()).abs().max() / st.nlml().abs().max()),
                      "alpha_vs_fp64_path": float((st_e.alpha - st.alpha).abs().max() / st.alpha.abs().max()),
                      "what": "GRBCM.update with options.emulate_fit_slices: K built in place, diagonal blocks of 1024 rows by the FP64 "
                              "Cholesky, panel solves and trailing updates as batched emulated-FP64 GEMMs (the shared communication block is "
                              "one panel against the cached L_g^-1), forward solve fused panel by panel"}
        del st_e

    # checksum of the aggregate, identical on every rank after the all-reduce: equal across N = 1/2/4/8 to ~1e-12 means the
    # sharded aggregation reproduces the single-GPU one (the driver-side multi-GPU parity check)
    agg_checksum = [float(mu.sum()), float(var.sum())]

    # ---- everything the parity leg needs from the GPU, taken now: the prediction model's HBM is handed to the training step next ----
    nc_s = min(8 if not args.small else 4, nc_loc - nc_loc % 2)

    def sample_side(src):
        """Per-expert values of the first nc_s local experts + their aggregate (through the aggregation KERNELS), on the host."""
        ml_s, vl_s = src["ml"][:nc_s].contiguous(), src["vl"][:nc_s].contiguous()
        S_s, _, _ = eng.grbcm_partial(ml_s, vl_s, src["vg"][0], 0)
        mu_s, var_s = eng.grbcm_finalize(S_s, src["mg"][0], src["vg"][0])[:2]
        return {"ml": ml_s.cpu(), "vl": vl_s.cpu(), "mg": src["mg"][0].cpu(), "vg": src["vg"][0].cpu(), "mu": mu_s.cpu(), "var": var_s.cpu()}

    nl_fit = st.nlml().cpu().numpy()
    gpu_side = None
    if rank == 0 and not args.no_cpu_baseline:
        gpu_side = sample_side(last)
        gpu_side["nlml_l"] = st.nlml()[:nc_s].cpu()
        gpu_side["nlml_g"] = float(model.gpg._state.nlml()[0])
        gpu_side["grad_g"] = P.MLE(model.gpg).loss_and_grad(make_hp(1, d)[0].numpy())[1]
        if emu is not None:
            gpu_side["emu"] = sample_side(emu_last)

    # ---- NLML + gradient of the local experts through the PUBLIC API: MLE(model.gpl).loss_and_grad (loss.py:92-128).  One factor
    # buffer per evaluation; beside the model's 137 GB there is no room for a second one on one GPU, so MLE borrows the model's
    # (the model is marked stale, as after set_params) — nothing below needs the prediction factors any more
    mle_api = None
    if not args.no_grad:
        mle = P.MLE(model.gpl)
        hp_np = make_hp(hi - lo, d).numpy()
        mle.loss(hp_np)                                       # warm-up (takes over the factor buffer)
        barrier(); e0.record(); nl_api, gr_api = mle.loss_and_grad(hp_np); e1.record(); barrier()
        api_ms = max_over_ranks(e0.elapsed_time(e1))
        api_flops = nc_total * float(n_l) ** 3                # N^3 per expert: Cholesky N^3/3 + L^-T N^3/3 + K^-1 N^3/3 (general path)
        mle_api = {"ms": api_ms, "tflops": api_flops / api_ms * 1e-9, "frac_of_peak": api_flops / api_ms * 1e-9 / (peak_tf * world),
                   "what": "MLE(model.gpl).loss_and_grad(hp [nc, P]) — per-expert NLML and gradient through the reference-shaped API, "
                           "incl. the [nc, 1+P] read-back; experts treated as unrelated matrices (N^3 flops each)",
                   "nlml_vs_fit": float(np.abs(np.asarray(nl_api) - nl_fit).max() / np.abs(nl_fit).max())}
        del mle

    # ---- config 5: one training evaluation through GRBCM_MLE (all ranks; takes over the HBM the prediction model held) ----------
    train = None
    if not args.no_train:
        del model, st, last, ml_fp64, vl_fp64, mg_fp64, vg_fp64
        if emu is not None:
            del emu_last
        eng.release_scratch()
        torch.cuda.empty_cache()
        train = run_train(args, P, eng, world, rank, dev, barrier, max_over_ranks, peak_tf)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # algorithmic flops per launch pair: N^2 per test point per expert; with the shared leading block the first
    # n_g rows of every expert's substitution are the communication expert's and are done once
    pred_flops_unshared = (nc_loc * float(n_l) ** 2 + float(n_g) ** 2) * m
    pred_flops_kernel = (nc_loc * (float(n_l) ** 2 - float(n_g) ** 2) + float(n_g) ** 2) * m if shared else pred_flops_unshared
    roof = {"bound": "tensor", "kernel": ("predict_kernel<shared> (local experts continue the communication expert's substitution; its V_g^T comes "
                                         "from vt_gemm_kernel inside the same bracket)") if shared else
            "predict_kernel (local experts; the communication expert's launch overlaps it on a side stream)", "achieved": pred_flops_kernel / kern_ms * 1e-9, "peak": peak_tf,
            "unit": "TFLOP/s", "frac": pred_flops_kernel / kern_ms * 1e-9 / peak_tf,
            # dram__bytes_read + dram__bytes_write of this launch from the ncu --set full captures of the default 1-GPU config:
            # shared block 1.811 TB + 73.1 GB (profiles/r01_s3_ncu_predict_shared_bench_scale.json), general path 2.533 TB +
            # 68.7 GB (profiles/r01_s2_ncu_predict_kernel_bench_scale.json); other configurations were not captured
            "traffic": (1.884e12 if shared else 2.602e12) if (world == 1 and not args.small and m == 4096) else None,
            "shared_leading_block": shared, "unshared_count_tflops": pred_flops_unshared / kern_ms * 1e-9,
            "peak_source": "cuBLAS dgemm 8192^3 (A B^T) measured in this run on this GPU (MEASURED_PEAKS.json has no FP64 entry)",
            "kernel_ms": kern_ms, "flops_per_launch": pred_flops_kernel}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(w, m),
            "run": {"experts_per_rank": nc_loc, "parallelism": "experts sharded over %d GPU(s); all-reduce of [3,m] sums" % world,
                    "shared_leading_block": shared, "factors_gb_this_rank": nc_loc * n_l * n_l * 8 / 1e9},
            "roofline": roof,
            "fit": {"ms": fit_ms, "tflops": fit_flops / fit_ms * 1e-9, "frac_of_peak": fit_flops / fit_ms * 1e-9 / (peak_tf * world),
                    "flops": fit_flops, "unshared_count_tflops": fit_flops_unshared / fit_ms * 1e-9,
                    "what": "batched Cholesky incl. kernel build, forward+backward solve; N^3/3 per expert, minus the shared "
                            "communication block (n_g^3/3 per expert) when one hyper-parameter vector is shared"},
            "clocks": clocks, "gpu_launches": launches, "agg_checksum": agg_checksum,
            # bytes that really cross PCIe per e2e step: the pinned test points in, the two result vectors out (the per-expert
            # [nc, m] means / variances stay on the device, gr_bcm.GRBCM._aggregate_device)
            "e2e": {"value": m / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": m * d * 8, "d2h_bytes_per_step": 2 * m * 8,
                    "ms_per_step": e2e_ms}}
    if emu is not None:
        line["emulated_fp64"] = emu
    if train is not None:
        line["train"] = train
    if mle_api is not None:
        line["mle_public_api"] = mle_api
    if grad_ms is not None:
        gflops_unshared = nc_total * float(n_l) ** 3 * 2.0 / 3.0
        gflops = nc_total * (float(n_l) ** 3 - float(n_g) ** 3) * 2.0 / 3.0 if shared else gflops_unshared
        line["lml_grad"] = {"ms": grad_ms, "tflops": gflops / grad_ms * 1e-9, "frac_of_peak": gflops / grad_ms * 1e-9 / (peak_tf * world),
                            "unshared_count_tflops": gflops_unshared / grad_ms * 1e-9,
                            "what": "L^-T + fused K^-1/gradient reduction, 2N^3/3 per expert (Cholesky reported under fit), minus the "
                                    "communication block's share (2 n_g^3/3 per expert) when one hyper-parameter vector is shared",
                            "cholesky_plus_grad_tflops": (gflops + fit_flops) / (grad_ms + fit_ms) * 1e-9}
    if not args.no_cpu_baseline:
        # ---- CPU arm on a bounded sample + PARITY of this run's GPU results against it ---------------------------------------
        # rank 0 owns local experts 0..nc_s-1 at every N (contiguous shards), so the same experts are compared in every SCALE line
        from oracle import cpu_arm
        torch.set_num_threads(os.cpu_count() or 1)
        cs = cpu_sample(w, m, nc_s)
        xs_cpu = xs_host.clone()
        cs.predict(xs_cpu)                                     # warm-up
        t0 = time.perf_counter(); ref = cs.predict(xs_cpu); pred_s = time.perf_counter() - t0
        ratio = scale_to_job(w, nc_s)
        line["cpu_baseline"] = {"value": m / (pred_s * ratio), "unit": UNIT, "cores": torch.get_num_threads(), "kind": cs.kind,
                                "sample": "%s: %d of %d local experts + comm expert, %d test points, predict %.2f s (fit %.2f s); "
                                          "scaled by the N^2 work ratio %.2f" %
                                          ("unmodified reference (baseline/_ref)" if cs.kind == "reference" else "oracle port", nc_s,
                                           nc_total, m, pred_s, cs.fit_s, ratio)}
        nl_g_ref, nl_l_ref = cs.nlml()
        gr_ref = cs.grad_comm()
        gs = gpu_side
        par = {"vs": cs.kind, "experts": nc_s, "test_points": m, "tolerance": 1e-8, "metric": "max|a-b| / max|b|",
               "mean_expert": cpu_arm.rel_err(gs["ml"], ref["mean_l"]), "var_expert": cpu_arm.rel_err(gs["vl"], ref["var_l"]),
               "mean_comm": cpu_arm.rel_err(gs["mg"], ref["mean_g"]), "var_comm": cpu_arm.rel_err(gs["vg"], ref["var_g"]),
               "mean_aggregate_sample": cpu_arm.rel_err(gs["mu"], ref["mu"]), "var_aggregate_sample": cpu_arm.rel_err(gs["var"], ref["var"]),
               "nlml_expert": cpu_arm.rel_err(gs["nlml_l"], nl_l_ref), "nlml_comm": abs(gs["nlml_g"] - nl_g_ref) / abs(nl_g_ref),
               "grad_comm": cpu_arm.rel_err(gs["grad_g"], gr_ref)}
        if train is not None:
            tp = train_parity(P, args.small)
            train["parity_at_reduced_n"] = tp
            par["train_nlml_reduced_n"], par["train_grad_reduced_n"] = tp["nlml"], tp["grad"]
        par["worst"] = max(v for k, v in par.items() if k.startswith(("mean", "var", "nlml", "grad", "train")))
        par["ok"] = par["worst"] <= par["tolerance"]
        line["parity"] = par
        if emu is not None:
            ge = gs["emu"]
            emu["parity"] = {"vs": cs.kind, "experts": nc_s, "mean_expert": cpu_arm.rel_err(ge["ml"], ref["mean_l"]),
                             "var_expert": cpu_arm.rel_err(ge["vl"], ref["var_l"]),
                             "mean_aggregate_sample": cpu_arm.rel_err(ge["mu"], ref["mu"]),
                             "var_aggregate_sample": cpu_arm.rel_err(ge["var"], ref["var"])}
            emu["parity"]["ok"] = max(v for k, v in emu["parity"].items() if k.startswith(("mean", "var"))) <= 1e-8
    out.write(json.dumps(line) + "\n")
    out.flush()
    if world > 1:
        dist.destroy_process_group()
    if "parity" in line and not line["parity"]["ok"]:
        sys.stderr.write("bench.py: PARITY FAILED against the CPU arm: %r\n" % (line["parity"],))
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
