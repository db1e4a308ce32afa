"""Tuning of the pipelined peer-store exchange under torchrun: ms/step of the sharded 1024^3
Cahn-Hilliard step for (nchunks, comm_ctas) pairs, one solver, CUDA events, max over ranks.
  torchrun --nproc-per-node P scripts/tune_dist.py [grid] [chunks:ctas ...]"""
import os
import sys

import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import dist as gdist  # noqa: E402
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
args = sys.argv[1:]
n = int(args[0]) if args else 1024
pairs = [tuple(int(x) for x in a.split(":")) for a in args[1:]] or [(1, 0), (4, 32), (4, 48), (4, 64), (8, 48)]
cells = n ** 3 // world
model = gpf.NewModel()
conc = gpf.NewField("conc", cells, None, pinned=True)
conc.Data[:] = 0.0
conc.Data[::7] = 0.5 + 0.01 * rank
model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
model.AddField(conc)
model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
solver = gdist.ShardedSolver(model, n, synthetic.CAHN_HILLIARD_DT, device=local, exchange="peer")
solver.Upload()
for nch, ctas in pairs:
    solver.nchunks, solver.comm_ctas = nch, ctas
    solver.StepDevice(3)
    torch.cuda.synchronize(); tdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(solver.stream)
    solver.StepDevice(10)
    e1.record(solver.stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda", dtype=torch.float64)
    tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
    if rank == 0:
        print(f"P={world} {n}^3 chunks={nch} comm_ctas={ctas}: {ms.item():.3f} ms/step = {n ** 3 / ms.item() / 1e6:.1f} G cell-updates/s", flush=True)
tdist.barrier()
tdist.destroy_process_group()
