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
for n, nbw in ((1000, 128), (5001, 256), (9000, 512)):
    x, y = O.make_data(n, 8, seed=0)
    k = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8)
    gp = GaussianProcess(kernel=k, s=0.1)
    gp.fit_gp(x.cuda(), y.cuda())
    ref = float(gp.log_marginal(k, {}, 1.0))
    for la in (True, False):
        dg = DistributedGP(k, s=0.1, nbw=nbw, lookahead=la)
        dg.fit_gp(x.cuda(), y.cuda())
        val = float(dg.log_marginal(1.0))
        ea = float((dg.A - gp.A).abs().max() / gp.A.abs().max())
        if rank == 0:
            print("world=%d n=%d nbw=%d lookahead=%s: lml diff %.3e alpha relerr %.3e" % (dist.get_world_size(), n, nbw, la, abs(val - ref), ea), flush=True)
        assert abs(val - ref) < 1e-8 and ea < 1e-9
if rank == 0:
    print("dist multi ok")
dist.barrier()
dist.destroy_process_group()
