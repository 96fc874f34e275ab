"""One rank's share of config 4 on one GPU: `nc` local experts of order 8192 + the communication expert (order 4096).
nc=32 is what each of 8 GPUs holds.  Times GRBCM.update (fit) and one predict step, to see what the 8-GPU run will do."""
import math, os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pygpr_b200 as P
from pygpr_b200 import engine as eng

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
m0, d = 4096, 16
g = torch.Generator().manual_seed(0)
x = torch.rand(nc + 1, m0, d, generator=g, dtype=torch.float64)
y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(nc + 1, m0, generator=g, dtype=torch.float64)
hp = torch.empty(nc, d + 2, dtype=torch.float64); hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
cov = P.Compose([P.Squared_exponential(), P.White_noise()])
model = P.GRBCM(x[1:], y[1:], x[0], y[0], cov, expert_offset=0, nc_total=255)
model.gpl.set_params(hp); model.gpg.set_params(hp[0])
xs = torch.rand(m, d, generator=g, dtype=torch.float64).cuda()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def refit_overlapped():
    model.gpl.need_upd = True; model.gpg.need_upd = True
    model.update()


def refit_serial():
    model.gpl.need_upd = True; model.gpg.need_upd = True
    model.gpg.update(); model.gpl.update()


model.update()
flops = nc * 8192 ** 3 / 3 + 4096 ** 3 / 3
for name, fn in (("serial", refit_serial), ("overlapped", refit_overlapped)):
    t = timed(fn)
    print(f"fit {name:10s} nc={nc}: {t:8.2f} ms  {flops / t * 1e-9:6.2f} TFLOP/s", flush=True)
st_g, st_l = model.gpg._state, model.gpl._state
pf = (nc * 8192.0 ** 2 + 4096.0 ** 2) * m
t = timed(lambda: (eng.predict(st_g, xs, True), eng.predict(st_l, xs, True)))
print(f"predict serial     nc={nc} m={m}: {t:8.2f} ms  {pf / t * 1e-9:6.2f} TFLOP/s", flush=True)
t = timed(lambda: eng.predict_pair(st_g, st_l, xs))
print(f"predict overlapped nc={nc} m={m}: {t:8.2f} ms  {pf / t * 1e-9:6.2f} TFLOP/s", flush=True)
t = timed(lambda: eng.lml_grad(st_l))
print(f"lml_grad           nc={nc}: {t:8.2f} ms  {nc * 8192.0 ** 3 * 2 / 3 / t * 1e-9:6.2f} TFLOP/s", flush=True)
