"""World-size-2 (and 4) gloo run of the slab-sharded orchestration (gopf_b200/dist.py) on CPU.

The CUDA phases are replaced by a numpy stand-in written to the same layout contract
(include/gopf_cuda.h "slab-sharded step"); the orchestration, the all-to-all block order
and the split pack/unpack maps are the code under test.  The result must equal the
unsharded oracle.  The real CUDA phases are checked against the same oracle on GPUs in
tests/test_dist_gpu.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from gopf_b200 import dist as gdist
from gopf_b200 import synthetic
from oracle import pf as opf
from oracle import pfutil as opfutil


class NumpyPhases:
    """Phase semantics of csrc/dist_solver.h in numpy (Cahn-Hilliard via the oracle's Model)."""

    def __init__(self, n, world, rank, dt):
        self.n, self.world, self.rank, self.m, self.dt = n, world, rank, n // world, dt
        n_loc = self.m * n * n
        # oracle model on the LOCAL k-points: bricks alias the arrays set in kspace_step
        self.model = opf.NewModel()
        self.field = opf.NewField("conc", n_loc)
        self.model.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        self.model.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        self.model.AddField(self.field)
        self.model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
        self.model.Init()
        # Freq of every local spectrum cell [k0][k1l][k2]: reference node = (k0*n + k1)*n + k2
        k0, k1l, k2 = np.meshgrid(np.arange(n), np.arange(self.m), np.arange(n), indexing="ij")
        node = (k0 * n + (rank * self.m + k1l)) * n + k2
        full = opfutil.NewFFTW([n, n, n]).freq_table()
        tab = full[node.reshape(-1)]
        self.freq = opf.Frequency(lambda i: list(tab[i]), lambda cnt: tab)

    @staticmethod
    def v(t, *shape):
        return t.numpy().reshape(*shape)

    def _pack(self, y, send):  # y[i0l][k1][k2] -> send[q][i0l][k1l][k2]
        m, n, w = self.m, self.n, self.world
        self.v(send, w, m, m, n)[...] = y.reshape(m, w, m, n).transpose(1, 0, 2, 3)

    def forward_local(self, W, send):
        w = self.v(W, self.m, self.n, self.n)
        self._pack(np.fft.fft(np.fft.fft(w, axis=2), axis=1), send)

    def forward_mid(self, W, send):
        self._pack(np.fft.fft(self.v(W, self.m, self.n, self.n), axis=1), send)

    def forward_finish(self, T):
        t = self.v(T, self.n, self.m, self.n)
        t[...] = np.fft.fft(t, axis=0)

    def inverse_start(self, S, T):
        self.v(T, self.n, self.m, self.n)[...] = np.fft.ifft(self.v(S, self.n, self.m, self.n), axis=0) * self.n

    def inverse_mid(self, recv, W):  # recv[p][i0l][k1l][k2] -> y[i0l][k1 = p*m + k1l][k2]
        m, n, w = self.m, self.n, self.world
        y = self.v(recv, w, m, m, n).transpose(1, 0, 2, 3).reshape(m, n, n)
        self.v(W, m, n, n)[...] = np.fft.ifft(y, axis=1) * n

    def real_step(self, W):
        w = self.v(W, self.m, self.n, self.n)
        c = np.fft.ifft(w, axis=2) * self.n / float(self.n) ** 3
        w[...] = np.fft.fft(opfutil.go_cpow(c, 3.0), axis=2)

    def kspace_step(self, T, S):
        t = self.v(T, self.n, self.m, self.n)
        g = np.fft.fft(t, axis=0).reshape(-1)
        s = S.numpy()
        self.field.Data[:] = s
        self.model.DerivedFields[0].Data[:] = g
        rhs = self.model.GetRHS(0, self.freq, 0.0)
        den = self.model.GetDenum(0, self.freq, 0.0)
        s[:] = (s + self.dt * rhs) / (1.0 - self.dt * den)
        self.inverse_start(S, T)

    def inverse_finish(self, W, out):
        self.v(out, self.m, self.n, self.n)[...] = np.fft.ifft(self.v(W, self.m, self.n, self.n), axis=2) * self.n / float(self.n) ** 3

    def advance(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, nsteps, split_steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = n // world
        cells = m * n * n
        init = synthetic.cahn_hilliard_initial(cells, 0, offset=rank * cells)  # this rank's slab of the global field
        mk = lambda: torch.zeros(cells, dtype=torch.complex128)
        S, A, B = mk(), mk(), mk()
        A.numpy()[:] = init
        phases = NumpyPhases(n, world, rank, synthetic.CAHN_HILLIARD_DT)

        def a2a(dst, src):
            tdist.all_to_all_single(torch.view_as_real(dst), torch.view_as_real(src))

        gdist.upload(phases, a2a, A, S, B)
        valid = False
        for k in split_steps:  # several run_steps calls: a_valid must carry across them
            valid = gdist.run_steps(phases, a2a, S, A, B, k, valid)
        assert sum(split_steps) == nsteps
        out = mk()
        gdist.download(phases, a2a, S, A, B, out)
        np.save(os.path.join(out_dir, f"slab{rank}.npy"), out.numpy())
    finally:
        tdist.destroy_process_group()


@pytest.mark.parametrize("world,n,split", [(2, 16, (3, 2)), (4, 16, (5,)), (2, 8, (1, 1, 1, 1, 1))])
def test_sharded_orchestration_matches_unsharded_oracle(tmp_path, world, n, split):
    nsteps = sum(split)
    mp.spawn(_worker, args=(world, _free_port(), n, nsteps, split, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
    # unsharded oracle on the same global field (slab seeds derive from the global index)
    total = n ** 3
    m = opf.NewModel()
    f = opf.NewField("conc", total, synthetic.cahn_hilliard_initial(total, 0))
    m.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    opf.NewSolver(m, [n, n, n], synthetic.CAHN_HILLIARD_DT).Propagate(nsteps)
    err = np.linalg.norm(got - f.Data) / np.linalg.norm(f.Data)
    assert err < 1e-12, err


def test_slab_seeds_are_consistent_for_every_world_size():
    n = 8
    whole = synthetic.cahn_hilliard_initial(n ** 3, 0)
    for world in (1, 2, 4, 8):
        cells = n ** 3 // world
        parts = [synthetic.cahn_hilliard_initial(cells, 0, offset=r * cells) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), whole)
