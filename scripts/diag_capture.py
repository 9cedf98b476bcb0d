import os, sys, time, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
from gbp_b200.dist import PartitionedBAGraph
from gbp_b200.synthetic import make_synthetic
CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
prob = make_synthetic(40, 6000, 8, seed=5)
pg = PartitionedBAGraph(prob, CFG, rank=rank, world=world, device=rank, stream=s.cuda_stream, dist=dist, torch_stream=s)
pg.generate_priors_var(50.0); pg.update_all_beliefs()
for i in range(3): pg.synchronous_iteration(robustify=True, local_relin=True)
torch.cuda.synchronize(); print(rank, "eager ok", pg.metrics(), flush=True)
ok = pg.capture(local_relin=True, robustify=True)
print(rank, "capture returned", ok, flush=True)
for i in range(3): pg.synchronous_iteration(robustify=True, local_relin=True)
torch.cuda.synchronize(); print(rank, "replay ok", pg.metrics(), flush=True)
pg.close()
dist.destroy_process_group()
print(rank, 'destroyed', flush=True)
