"""torchrun --nproc-per-node N tools/dist_multi_check.py : distributed fit + LML vs the single-GPU path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle import stpy_oracle as O
from stpy_b200.kernels import KernelFunction
from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
from stpy_b200.distributed import DistributedGP

lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank = dist.get_rank()
for n, nbw in ((1000, 128), (5001, 256), (9000, 512), (9100, 1024)):
    x, y = O.make_data(n, 8, seed=0)
    k = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.fit_gp(x.cuda(), y.cuda())
    ref = float(gp.log_marginal(k, {}, 1.0))
    for la, depth in ((True, None), (True, 1), (True, 3), (False, None)):
        dg = DistributedGP(k, s=0.1, nbw=nbw, lookahead=la, depth=depth)
        dg.p2p = la  # peer-memory sweep with look-ahead, NCCL-broadcast sweep without
        dg.fit_gp(x.cuda(), y.cuda())
        dg.fit_gp(x.cuda(), y.cuda())  # second fit: flags carry a new epoch
        val = float(dg.log_marginal(1.0))
        ea = float((dg.A - gp.A).abs().max() / gp.A.abs().max())
        if rank == 0:
            print("world=%d n=%d nbw=%d lookahead=%s depth=%s: lml diff %.3e alpha relerr %.3e" % (dist.get_world_size(), n, nbw, la, dg.depth, abs(val - ref), ea), flush=True)
        assert abs(val - ref) < 1e-8 and ea < 1e-9
        dg.check()
        if la and depth is None:
            xt, _ = O.make_data(100, 8, seed=9)
            mu, sd = dg.mean_std(xt.cuda())
            mu1, sd1 = gp.mean_std(xt.cuda())
            ep = max(float((mu - mu1).abs().max() / mu1.abs().max()), float((sd ** 2 - sd1 ** 2).abs().max() / (sd1 ** 2).abs().max()))
            if rank == 0:
                print("   distributed mean_std vs single GPU: %.2e" % ep, flush=True)
            assert ep < 1e-10
        dg.close()
# RFF regression with row-sharded normal equations (one all-reduce) vs the single-rank fit
import numpy as np
from stpy_b200.embeddings.embedding import RFFEmbedding
from stpy_b200.continuous_processes.kernelized_features import KernelizedFeatures
from stpy_b200.sweep import lml_sweep, lml_sweep_distributed
x, y = O.make_data(20001, 16, seed=3)
xt, _ = O.make_data(64, 16, seed=4)
np.random.seed(0)
emb = RFFEmbedding(gamma=1.0, m=512, d=16)
kf1 = KernelizedFeatures(embedding=emb, m=512, s=0.1, lam=1.0, d=16)
kf1.fit_gp(x.cuda(), y.cuda())
kfd = KernelizedFeatures(embedding=emb, m=512, s=0.1, lam=1.0, d=16)
kfd.distributed = True
kfd.fit_gp(x.cuda(), y.cuda())
m1, s1 = kf1.mean_std(xt.cuda())
md, sd = kfd.mean_std(xt.cuda())
e1 = float((m1 - md).abs().max() / m1.abs().max()); e2 = float((s1 - sd).abs().max() / s1.abs().max())
if rank == 0:
    print("rff sharded: mean relerr %.2e std relerr %.2e" % (e1, e2), flush=True)
assert e1 < 1e-10 and e2 < 1e-9
ks = [KernelFunction(kernel_name="squared_exponential", gamma=float(g), d=8) for g in np.logspace(-0.5, 0.5, 6)] + \
     [KernelFunction(kernel_name="matern", gamma=float(g), nu=2.5, d=8) for g in np.logspace(-0.5, 0.5, 5)]
xs, ys = O.make_data(1500, 8, seed=5)
v1 = lml_sweep(ks, xs.cuda(), ys.cuda(), s=0.1)
vd = lml_sweep_distributed(ks, xs.cuda(), ys.cuda(), s=0.1)
if rank == 0:
    print("sweep sharded: max diff %.2e" % float((v1 - vd).abs().max()), flush=True)
assert float((v1 - vd).abs().max()) < 1e-9
if rank == 0:
    print("dist multi ok")
dist.barrier()
dist.destroy_process_group()
