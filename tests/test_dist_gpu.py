"""Slab-sharded CUDA phases (gopf_dist_*) against the unsharded oracle.

world = 1 runs on any GPU box (the exchange degenerates to a copy but every phase kernel,
including the split row maps, is exercised); world = 2 needs two GPUs and NCCL.
"""
import os
import socket

import numpy as np
import pytest
import torch

from gopf_b200 import synthetic
from oracle import pf as opf

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle(n, nsteps):
    total = n ** 3
    m = opf.NewModel()
    f = opf.NewField("conc", total, synthetic.cahn_hilliard_initial(total, 0))
    m.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    opf.NewSolver(m, [n, n, n], synthetic.CAHN_HILLIARD_DT).Propagate(nsteps)
    return f.Data


def _worker(rank, world, port, n, split, out_dir, exchange):
    import torch.distributed as tdist
    from gopf_b200 import dist as gdist
    from gopf_b200 import pf as gpf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    tdist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world)
    try:
        cells = n ** 3 // world
        model = gpf.NewModel()
        f = gpf.NewField("conc", cells, synthetic.cahn_hilliard_initial(cells, 0, offset=rank * cells))
        model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        model.AddField(f)
        model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
        s = gdist.ShardedSolver(model, n, synthetic.CAHN_HILLIARD_DT, device=rank, exchange=exchange)
        s.Upload()
        for i, k in enumerate(split):
            s.StepDevice(k)
            if i == 0 and len(split) > 1:
                s.Download()  # a host read-back between epochs must not disturb the device state
        s.Download()
        torch.cuda.synchronize()
        assert s.phases.kernel_launches() > 0
        np.save(os.path.join(out_dir, f"slab{rank}.npy"), f.Data)
    finally:
        tdist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "dma", "nccl"])
@pytest.mark.parametrize("world,n,split", [(1, 32, (4, 3)), (1, 64, (10,)), (2, 32, (4, 3)), (2, 64, (10,)), (4, 64, (3, 2)),
                                           (8, 64, (5,))])
def test_sharded_cuda_matches_oracle(tmp_path, world, n, split, exchange):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), n, split, str(tmp_path), exchange), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
    ref = _oracle(n, sum(split))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-10


def _single_gpu(n, nsteps):
    """The same run on one GPU through the fused single-field kernels (the P = 1 member of SURVEY.md 8d's
    "P-GPU vs 1-GPU" comparison)."""
    from gopf_b200 import pf as gpf
    total = n ** 3
    m = gpf.NewModel()
    f = gpf.NewField("conc", total, synthetic.cahn_hilliard_initial(total, 0))
    m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = gpf.NewSolver(m, [n, n, n], synthetic.CAHN_HILLIARD_DT)
    assert s.IsFused
    s.Upload()
    s.StepDevice(nsteps)
    s.Download()
    out = f.Data.copy()
    s.close()
    return out


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("world,n,split", [(1, 256, (6, 4)), (2, 256, (10,)), (4, 256, (10,)), (8, 256, (10,)),
                                           (2, 512, (4,)), (4, 512, (4,)), (8, 512, (4,))])
def test_sharded_matches_single_gpu_at_benchmark_scale(tmp_path, world, n, split, exchange):
    """P-GPU vs 1-GPU at 256^3 and 512^3, <= 1e-13 (same kernels, different tiling and
    exchange; world = 1 runs every PEER / split-row-map instantiation of the 256-cell lines on any box)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), n, split, str(tmp_path), exchange), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
    ref = _single_gpu(n, sum(split))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-13


def _pipelined_world1(rank, world, port, n, steps, out_dir):
    """One rank, but the real-space side pipelined in 8 plane chunks with the peer-storing pass on a second stream
    (what P > 1 runs): copy-engine kernels on plane chunks under a grid cap, incl. the paired real-space kernel."""
    import torch.distributed as tdist
    from gopf_b200 import dist as gdist
    from gopf_b200 import pf as gpf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GOPF_TMA_MIN_N"] = str(n)
    torch.cuda.set_device(0)
    tdist.init_process_group("gloo", rank=0, world_size=1)
    try:
        cells = n ** 3
        model = gpf.NewModel()
        f = gpf.NewField("conc", cells, synthetic.cahn_hilliard_initial(cells, 0))
        model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        model.AddField(f)
        model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
        s = gdist.ShardedSolver(model, n, synthetic.CAHN_HILLIARD_DT, device=0, exchange="peer")
        s.Upload()
        with torch.cuda.stream(s.stream):
            s.a_valid = gdist.run_steps_peer(s.phases, s.barrier, s.S, s.A, steps, s.a_valid, s.slab, 8, 48)
        s.Download()
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, "slab0.npy"), f.Data)
    finally:
        tdist.destroy_process_group()


def test_pipelined_plane_chunks_with_copy_engine_kernels_match_single_gpu(tmp_path):
    """The chunked, two-stream real-space side of the sharded step at 512-cell lines with the copy-engine kernels
    switched on for them (GOPF_TMA_MIN_N=512): against the single-GPU fused step, <= 1e-13."""
    import torch.multiprocessing as mp
    n, steps = 512, 3
    mp.spawn(_pipelined_world1, args=(1, _free_port(), n, steps, str(tmp_path)), nprocs=1, join=True)
    got = np.load(tmp_path / "slab0.npy")
    ref = _single_gpu(n, steps)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-13
