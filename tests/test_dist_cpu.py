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


class NumpyPeerPhases(NumpyPhases):
    """The peer-store phases of csrc/dist_solver.h.  On GPUs a *_peer phase stores each row
    straight into its owner's X / Y over NVLink; gloo has no peer memory, so the stand-in packs
    and runs the all-to-all inside the phase.  Orchestration (gopf_b200.dist.run_steps_peer:
    phase order, X-validity bookkeeping, barrier placement) is the code under test."""

    def __init__(self, n, world, rank, dt):
        super().__init__(n, world, rank, dt)
        cells = self.m * n * n
        self.X = torch.zeros(cells, dtype=torch.complex128)
        self.Y = torch.zeros(cells, dtype=torch.complex128)
        self._send = torch.zeros(cells, dtype=torch.complex128)
        self.barriers = 0

    def _a2a(self, dst, src):
        tdist.all_to_all_single(torch.view_as_real(dst), torch.view_as_real(src))

    def barrier(self):
        self.barriers += 1
        tdist.barrier()

    def set_grid_cap(self, ctas, reserve=False):  # launch shaping only; nothing to emulate
        self.grid_caps = getattr(self, "grid_caps", []) + [(ctas, reserve)]

    def inverse_start_peer(self, S):
        self.inverse_start(S, self._send)
        self._a2a(self.X, self._send)

    def inverse_mid_x(self, W):
        self.inverse_mid(self.X, W)

    def forward_mid_peer(self, W):
        self.forward_mid(W, self._send)
        self._a2a(self.Y, self._send)

    def forward_local_peer(self, W):
        self.forward_local(W, self._send)
        self._a2a(self.Y, self._send)

    def forward_finish_peer(self, S):
        S.copy_(self.Y)
        self.forward_finish(S)

    def kspace_step_peer(self, S):
        self.kspace_step(self.Y, S)   # leaves the first inverse pass of the new S in Y's storage
        self._send.copy_(self.Y)
        self._a2a(self.X, self._send)


class NumpyDmaPhases(NumpyPeerPhases):
    """Chunked phases + copy-engine exchange of csrc/dist_solver.h; a DMA copy into a peer's
    buffer is emulated by an all-to-all of the chunk."""

    def inverse_mid_planes_x(self, W, b, c):
        m, n, w = self.m, self.n, self.world
        y = self.v(self.X, w, m, m, n)[:, b:b + c].transpose(1, 0, 2, 3).reshape(c, n, n)
        self.v(W, m, n, n)[b:b + c] = np.fft.ifft(y, axis=1) * n

    def real_step_planes(self, W, b, c):
        w = self.v(W, self.m, self.n, self.n)[b:b + c]
        cc = np.fft.ifft(w, axis=2) * self.n / float(self.n) ** 3
        w[...] = np.fft.fft(opfutil.go_cpow(cc, 3.0), axis=2)

    def forward_mid_planes(self, W, send, b, c):
        m, n, w = self.m, self.n, self.world
        y = np.fft.fft(self.v(W, m, n, n)[b:b + c], axis=1)
        self.v(send, w, m, m, n)[:, b:b + c] = y.reshape(c, w, m, n).transpose(1, 0, 2, 3)

    def kspace_step_cols_y(self, S, Tout, kb, kc):
        n, m = self.n, self.m
        t = self.v(self.Y, n, m, n)[:, kb:kb + kc]
        g = np.fft.fft(t, axis=0)
        s = self.v(S, n, m, n)
        sel = np.zeros((n, m, n), dtype=bool)
        sel[:, kb:kb + kc] = True
        idx = np.nonzero(sel.reshape(-1))[0]
        self.field.Data[:] = S.numpy()
        self.model.DerivedFields[0].Data[:] = 0.0
        self.model.DerivedFields[0].Data[idx] = g.reshape(-1)
        rhs = self.model.GetRHS(0, self.freq, 0.0)
        den = self.model.GetDenum(0, self.freq, 0.0)
        new = ((S.numpy() + self.dt * rhs) / (1.0 - self.dt * den)).reshape(n, m, n)
        s[:, kb:kb + kc] = new[:, kb:kb + kc]
        self.v(Tout, n, m, n)[:, kb:kb + kc] = np.fft.ifft(s[:, kb:kb + kc], axis=0) * n

    def exchange_forward(self, send, b, c):
        m, n, w = self.m, self.n, self.world
        tmp = torch.from_numpy(np.ascontiguousarray(self.v(send, w, m, m, n)[:, b:b + c]).reshape(-1))
        got = torch.zeros_like(tmp)
        self._a2a(got, tmp)
        self.v(self.Y, w, m, m, n)[:, b:b + c] = got.numpy().reshape(w, c, m, n)

    def exchange_inverse(self, T, kb, kc):
        m, n, w = self.m, self.n, self.world
        tmp = torch.from_numpy(np.ascontiguousarray(self.v(T, w, m, m, n)[:, :, kb:kb + kc]).reshape(-1))
        got = torch.zeros_like(tmp)
        self._a2a(got, tmp)
        self.v(self.X, w, m, m, n)[:, :, kb:kb + kc] = got.numpy().reshape(w, m, kc, n)

    def exchange_join(self):
        pass

    def forward_mid_peer_planes(self, W, b, c, max_ctas):
        self.forward_mid_planes(W, self._send, b, c)
        self.exchange_forward(self._send, b, c)


def _dma_worker(rank, world, port, n, split_steps, nchunks, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slab = n // world
        cells = slab * n * n
        mk = lambda: torch.zeros(cells, dtype=torch.complex128)
        S, A, SEND = mk(), mk(), mk()
        A.numpy()[:] = synthetic.cahn_hilliard_initial(cells, 0, offset=rank * cells)
        phases = NumpyDmaPhases(n, world, rank, synthetic.CAHN_HILLIARD_DT)
        gdist.upload_dma(phases, phases.barrier, A, S, SEND, slab)
        valid = False
        for k in split_steps:
            valid = gdist.run_steps_dma(phases, phases.barrier, S, A, SEND, k, valid, slab, nchunks)
        out = mk()
        gdist.download_dma(phases, phases.barrier, S, A, SEND, out, valid, slab)
        np.save(os.path.join(out_dir, f"slab{rank}.npy"), out.numpy())
    finally:
        tdist.destroy_process_group()


def _peer_worker(rank, world, port, n, split_steps, out_dir, nchunks=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cells = (n // world) * n * n
        mk = lambda: torch.zeros(cells, dtype=torch.complex128)
        S, A = mk(), mk()
        A.numpy()[:] = synthetic.cahn_hilliard_initial(cells, 0, offset=rank * cells)
        phases = NumpyDmaPhases(n, world, rank, synthetic.CAHN_HILLIARD_DT)
        gdist.upload_peer(phases, phases.barrier, A, S)
        valid = False
        mid = mk()
        for i, k in enumerate(split_steps):
            valid = gdist.run_steps_peer(phases, phases.barrier, S, A, k, valid, n // world, nchunks, 8)
            if i == 0:  # a download between epochs must not disturb the run (X stays valid)
                valid = gdist.download_peer(phases, phases.barrier, S, A, mid, valid)
                assert valid
        out = mk()
        valid = gdist.download_peer(phases, phases.barrier, S, A, out, valid)
        # 1 (upload) + first step's 2 around inverse_start_peer + 2 per step
        assert phases.barriers == 1 + 2 + 2 * sum(split_steps)
        np.save(os.path.join(out_dir, f"slab{rank}.npy"), out.numpy())
    finally:
        tdist.destroy_process_group()


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


@pytest.mark.parametrize("world,n,split,nchunks", [(2, 16, (3, 2), 1), (4, 16, (5,), 1), (2, 16, (2, 2), 4), (4, 16, (3,), 3)])
def test_peer_store_orchestration_matches_unsharded_oracle(tmp_path, world, n, split, nchunks):
    nsteps = sum(split)
    mp.spawn(_peer_worker, args=(world, _free_port(), n, split, str(tmp_path), nchunks), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
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


@pytest.mark.parametrize("world,n,split,nchunks", [(2, 16, (3, 2), 4), (4, 16, (4,), 3), (2, 8, (2,), 1)])
def test_dma_pipelined_orchestration_matches_unsharded_oracle(tmp_path, world, n, split, nchunks):
    nsteps = sum(split)
    mp.spawn(_dma_worker, args=(world, _free_port(), n, split, nchunks, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
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


def test_chunks_cover_the_range():
    for extent in (1, 2, 7, 16, 128):
        for k in (1, 3, 4, 200):
            parts = gdist.chunks(extent, k)
            assert parts[0][0] == 0 and sum(c for _, c in parts) == extent
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(len(parts) - 1))
            assert all(c >= 1 for _, c in parts)


def test_slab_seeds_are_consistent_for_every_world_size():
    n = 8
    whole = synthetic.cahn_hilliard_initial(n ** 3, 0)
    for world in (1, 2, 4, 8):
        cells = n ** 3 // world
        parts = [synthetic.cahn_hilliard_initial(cells, 0, offset=r * cells) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), whole)
