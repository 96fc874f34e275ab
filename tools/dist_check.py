"""Multi-GPU check under torchrun (one rank per GPU, NCCL): a sharded GRBCM (shard=True) must give, on EVERY rank, the
prediction, the all-gathered weights, and GRBCM_MLE's summed NLML + gradient of the unsharded model (computed by rank 0 on its
own GPU with all experts).  Exercises both data-path collectives: all-reduce of the [3, m] sums and of the [1+P] scalars.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import math, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pygpr_b200 as P


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solo = [dist.new_group([r]) for r in range(world)][rank]   # a one-rank group: collectives on it are no-ops
    worst = 0.0
    for (nc, nls, ng, d, m, per) in [(7, 200, 256, 5, 300, False), (5, 150, 100, 3, 77, True), (9, 130, 128, 8, 129, False)]:
        g = torch.Generator().manual_seed(100 + nc)
        xl = torch.rand(nc, nls, d, generator=g, dtype=torch.float64)
        yl = torch.sin(xl.sum(-1)) + 0.1 * torch.randn(nc, nls, generator=g, dtype=torch.float64)
        xg = torch.rand(ng, d, generator=g, dtype=torch.float64)
        yg = torch.sin(xg.sum(-1)) + 0.1 * torch.randn(ng, generator=g, dtype=torch.float64)
        xs = torch.rand(m, d, generator=g, dtype=torch.float64)
        hp = torch.empty(nc, d + 2, dtype=torch.float64)
        hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
        if per:
            hp = hp * (0.8 + 0.4 * torch.rand(nc, d + 2, generator=g, dtype=torch.float64))
        cov = P.Compose([P.Squared_exponential(), P.White_noise()])
        sharded = P.GRBCM(xl, yl, xg, yg, cov, shard=True)
        lo, hi = sharded.expert_offset, sharded.expert_offset + sharded.nc
        sharded.gpl.set_params(hp[lo:hi]); sharded.gpg.set_params(hp[0])
        mu, var = sharded.predict(xs, var="diag")
        beta, prec = sharded.gather_weights()
        # the unsharded model, on this rank's GPU, outside the process group's collectives
        full = P.GRBCM(xl, yl, xg, yg, cov, group=solo)
        full.gpl.set_params(hp); full.gpg.set_params(hp[0])
        mu0, var0 = full.predict(xs, var="diag")

        def rel(a, b):
            return float((torch.as_tensor(a) - torch.as_tensor(b)).abs().max() / torch.as_tensor(b).abs().max())

        errs = {"mu": rel(mu, mu0), "var": rel(var, var0), "beta": rel(beta, full.beta), "prec": rel(prec, full.prec)}
        if not per:
            f, gr = P.GRBCM_MLE(sharded, chunk=2).loss_and_grad(hp[0].numpy())
            f0, gr0 = P.GRBCM_MLE(full, chunk=0).loss_and_grad(hp[0].numpy())
            errs["mle"] = abs(f - f0) / abs(f0)
            errs["mle_grad"] = float(np.max(np.abs(gr - gr0)) / np.max(np.abs(gr0)))
        worst = max(worst, max(errs.values()))
        print("rank %d/%d nc=%d experts [%d,%d) shared_block=%s: %s" % (rank, world, nc, lo, hi,
              sharded.gpl._state.shared_from is not None, {k: "%.1e" % v for k, v in errs.items()}), flush=True)
    # the C ABI's own communicator (gpr_comm_*): same sums as torch.distributed, and a sharded prediction routed through it
    from pygpr_b200 import config, dist as gdist
    a = torch.rand(3, 1000, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(rank))
    b = a.clone()
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    config.options.comm = "cabi"
    gdist.all_reduce_sum_(a)
    torch.cuda.synchronize()
    errs = {"cabi_allreduce": float((a - b).abs().max())}
    mu_c, var_c = sharded.predict(xs, var="diag")                  # last case above, [3, m] sums through gpr_comm_allreduce_f64
    errs["cabi_predict"] = max(float((mu_c - mu).abs().max()), float((var_c - var).abs().max()))
    config.options.comm = "torch"
    gdist.cabi_comm_destroy()
    worst = max(worst, max(errs.values()))
    print("rank %d/%d C-ABI communicator: %s" % (rank, world, {k: "%.1e" % v for k, v in errs.items()}), flush=True)
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t[0]) < 1e-10
    if rank == 0:
        print("dist_check:", "ok" if ok else "FAILED", "world", world, "worst rel err %.2e" % float(t[0]), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
