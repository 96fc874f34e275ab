"""Config 5's per-rank share (64 experts of order 16384 + comm 8192, d = 32) through GRBCM_MLE.loss_and_grad with every emulated form on:
wall time against the number of experts per batch.

    python tools/train_chunk_sweep.py [chunks=10,16,32,46]
"""
import gc
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygpr_b200 as P  # noqa: E402
from pygpr_b200 import config  # noqa: E402

chunks = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10, 16, 32, 46]
per, m0, d = 64, 8192, 32
g = torch.Generator().manual_seed(100)
xl = torch.rand(per, m0, d, generator=g, dtype=torch.float64)
yl = torch.sin(xl.sum(-1)) + 0.1 * torch.randn(per, m0, generator=g, dtype=torch.float64)
xg = torch.rand(m0, d, generator=g, dtype=torch.float64)
yg = torch.sin(xg.sum(-1)) + 0.1 * torch.randn(m0, generator=g, dtype=torch.float64)
cov = P.Compose([P.Squared_exponential(), P.White_noise()])
model = P.GRBCM(xl.cuda(), yl.cuda(), xg.cuda(), yg.cuda(), cov)
theta = torch.empty(d + 2, dtype=torch.float64)
theta[0] = 1.0; theta[1:d + 1] = 2.0 / math.sqrt(d); theta[d + 1] = 0.3
config.options.emulate_grad_slices = config.options.emulate_inverse_slices = config.options.emulate_fit_slices = 7
for c in chunks:
    mle = P.GRBCM_MLE(model, chunk=c)
    mle.set_params(theta)
    mle.loss_and_grad(theta.numpy())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f, gr = mle.loss_and_grad(theta.numpy())
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"chunk_experts": c, "auto_chunk_would_be": P.GRBCM_MLE(model)._auto_chunk(), "loss_and_grad_ms": ms, "nlml": f,
                      "hbm_free_gb": torch.cuda.mem_get_info()[0] / 1e9}), flush=True)
    del mle
    gc.collect()
    from pygpr_b200 import engine
    engine.release_scratch()
    torch.cuda.empty_cache()
