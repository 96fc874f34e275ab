"""Device-timed pass over the BASELINE.json configs that are parity cases rather than the bench line (configs 1-3):
public API, host tensors in/out. Prints one JSON line per config."""
import json, math, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import pygpr_b200 as P

def hp_well(B, d):
    hp = torch.empty(B, d + 2); hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    return hp

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out

def main():
    args = [a for a in sys.argv[1:] if a != "--emulate"]
    which = args or ["1", "2", "3"]
    mode = "fp64_dmma"
    if "--emulate" in sys.argv:   # every opt-in emulated-FP64 form at once (INT8 tcgen05 tensor cores)
        from pygpr_b200 import config
        config.options.emulate_slices = 6
        config.options.emulate_grad_slices = config.options.emulate_inverse_slices = config.options.emulate_fit_slices = 7
        mode = "emulated (predict 6 planes; fit / L^-T / K^-1 7 planes)"
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    torch.manual_seed(0)
    if "1" in which:  # exact GPR, 1-D sine, n=1000
        x = 10 * torch.rand(1000, 1); y = torch.sin(x[:, 0]); xs = 10 * torch.rand(1000, 1)
        gp = P.Exact_GP(x, y, cov); gp.set_params(hp_well(1, 1)[0])
        def fit(): gp.need_upd = True; gp.update()
        tf, _ = timed(fit)
        tp, _ = timed(lambda: gp.predict(xs, var="diag"))
        mle = P.MLE(gp); hp = gp.params.numpy()
        tl, _ = timed(lambda: mle.loss(hp)); tg, _ = timed(lambda: mle.loss_and_grad(hp))
        print(json.dumps({"mode": mode, "config": "1: exact GPR n=1000 d=1 m=1000", "fit_ms": tf * 1e3, "predict_ms": tp * 1e3,
                          "pred_per_s": 1000 / tp, "loss_ms": tl * 1e3, "loss_and_grad_ms": tg * 1e3}), flush=True)
    if "2" in which:  # exact GPR ARD-SE n=16384, d=8, LML + gradient
        n, d = 16384, 8
        x = torch.rand(n, d); y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(n)
        gp = P.Exact_GP(x, y, cov); gp.set_params(hp_well(1, d)[0])
        mle = P.MLE(gp); hp = gp.params.numpy()
        tl, nl = timed(lambda: mle.loss(hp), 2); tg, lg = timed(lambda: mle.loss_and_grad(hp), 2)
        tr, gamma = timed(lambda: P.get_learn_rate(gp.params, mle, 1e-4), 1)
        xs = torch.rand(1024, d)
        gp.update()
        tp, _ = timed(lambda: gp.predict(xs, var="diag"), 2)
        mu2, var2 = gp.predict(xs, var="diag")
        print(json.dumps({"mode": mode, "config": "2: exact GPR n=16384 d=8 LML+grad", "predict_1024_ms": tp * 1e3,
                          "pred_checksum": [float(torch.as_tensor(mu2).sum()), float(torch.as_tensor(var2).sum())], "loss_ms": tl * 1e3, "loss_tflops": n ** 3 / 3 / tl * 1e-12,
                          "loss_and_grad_ms": tg * 1e3, "loss_and_grad_tflops": float(n) ** 3 / tg * 1e-12, "nlml": float(nl),
                          "grad_norm": float(np.linalg.norm(lg[1])), "get_learn_rate_ms": tr * 1e3, "gamma": float(gamma)}), flush=True)
    if "3" in which:  # GRBCM 64 x 2048 + comm, d=8
        M, m0, d, m = 64, 2048, 8, 4096
        x = torch.rand(M, m0, d); y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(M, m0); xs = torch.rand(m, d)
        model = P.GRBCM(x[1:], y[1:], x[0], y[0], cov)
        model.gpl.set_params(hp_well(M - 1, d)); model.gpg.set_params(hp_well(1, d)[0])
        def fit():
            model.gpl.need_upd = True; model.gpg.need_upd = True; model.update()
        tf, _ = timed(fit, 2)
        tp, _ = timed(lambda: model.predict(xs, var="diag"), 2)
        flops = (M - 1) * (2 * m0) ** 3 / 3 + m0 ** 3 / 3
        print(json.dumps({"mode": mode, "config": "3: GRBCM 64x2048 d=8 (63 experts of order 4096 + comm), m=4096", "fit_ms": tf * 1e3,
                          "fit_tflops_unshared_count": flops / tf * 1e-12, "predict_ms": tp * 1e3, "pred_per_s": m / tp,
                          "shared_leading_block": model.gpl._state.shared_from is not None,
                          "agg_checksum": [float(torch.as_tensor(model.predict(xs, var="diag")[0]).sum()),
                                           float(torch.as_tensor(model.predict(xs, var="diag")[1]).sum())]}), flush=True)
    if "5" in which:  # one GPU's slice of config 5: GRBCM training, experts of order 16384 (8192 comm + 8192 local), d=32
        nc, m0, d = 16, 8192, 32
        x = torch.rand(nc + 1, m0, d); y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(nc + 1, m0)
        model = P.GRBCM(x[1:], y[1:], x[0], y[0], cov)
        theta = hp_well(1, d)[0]
        loss = P.GRBCM_MLE(model, chunk=0)
        th = theta.numpy().astype(np.float64)
        tl, l = timed(lambda: loss.loss(th), 1)
        tg, lg = timed(lambda: loss.loss_and_grad(th), 1)
        flops = nc * float(2 * m0) ** 3 + float(m0) ** 3
        print(json.dumps({"config": "5 (slice): GRBCM_MLE summed NLML + gradient, %d experts of order 16384 + comm expert, d=32 "
                                    "(one of 8 GPUs holds 64 of the 511)" % nc, "loss_ms": tl * 1e3, "loss_and_grad_ms": tg * 1e3,
                          "loss_and_grad_tflops_unshared_count": flops / tg * 1e-12, "nlml": float(l),
                          "grad_norm": float(np.linalg.norm(lg[1]))}), flush=True)

if __name__ == "__main__":
    main()
