"""Slab-sharded stepping: one process per GPU, torch.distributed for the exchange.

Layout and phase contract: include/gopf_cuda.h ("slab-sharded step") and
gopf_b200/csrc/dist_solver.h.  ``run_steps`` / ``upload`` / ``download`` below are
the whole orchestration; they only need an object with the phase methods and an
``all_to_all(dst, src)`` callable, so the same code runs over NCCL with the CUDA
phases (``CudaPhases``) and, in the CPU tests, over gloo with a numpy stand-in.
"""
from __future__ import annotations

import ctypes

from ._lib import check, lib


def run_steps(phases, all_to_all, S, A, B, nsteps: int, a_valid: bool) -> bool:
    """nsteps sharded Euler steps.  S: spectrum [k0][k1_local][k2]; A, B: work buffers.
    ``a_valid`` says A already holds the first inverse pass of S.  Returns the new a_valid."""
    for _ in range(nsteps):
        if not a_valid:
            phases.inverse_start(S, A)      # S -> A, inverse along axis 0
        all_to_all(B, A)                     # [q][i0l][k1l][k2] blocks -> [p][i0l][k1l(p)][k2]
        phases.inverse_mid(B, A)             # inverse axis 1, unpack folded into the read
        phases.real_step(A)                  # inverse axis 2, /N, g(c), forward axis 2
        phases.forward_mid(A, B)             # forward axis 1, pack folded into the write
        all_to_all(A, B)                     # -> [k0][k1l][k2]
        phases.kspace_step(A, S)             # forward axis 0, Euler update, inverse axis 0 -> A
        phases.advance()
        a_valid = True
    return a_valid


def upload(phases, all_to_all, real_slab, S, B):
    """real_slab [i0l][i1][i2] (destroyed) -> S = its share of the spectrum."""
    phases.forward_local(real_slab, B)
    all_to_all(S, B)
    phases.forward_finish(S)


def download(phases, all_to_all, S, A, B, real_out):
    """S -> real_out = this rank's slab of the real-space field."""
    phases.inverse_start(S, A)
    all_to_all(B, A)
    phases.inverse_mid(B, A)
    phases.inverse_finish(A, real_out)


class CudaPhases:
    """The phase methods over ``gopf_dist_*`` on torch CUDA tensors (complex128)."""

    def __init__(self, model, n: int, world: int, rank: int, dt: float, device: int):
        self._h = ctypes.c_void_p()
        check(lib().gopf_dist_solver_create(model._h, int(n), int(world), int(rank), ctypes.c_double(dt), int(device),
                                            ctypes.byref(self._h)))
        self._model = model
        cells = ctypes.c_int64(0)
        check(lib().gopf_dist_solver_local_cells(self._h, ctypes.byref(cells)))
        self.local_cells = cells.value

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def set_stream(self, stream: int):
        check(lib().gopf_dist_solver_set_stream(self._h, ctypes.c_void_p(stream)))

    def forward_local(self, W, send):
        check(lib().gopf_dist_forward_local(self._h, self._p(W), self._p(send)))

    def forward_finish(self, T):
        check(lib().gopf_dist_forward_finish(self._h, self._p(T)))

    def inverse_start(self, S, T):
        check(lib().gopf_dist_inverse_start(self._h, self._p(S), self._p(T)))

    def inverse_mid(self, recv, W):
        check(lib().gopf_dist_inverse_mid(self._h, self._p(recv), self._p(W)))

    def real_step(self, W):
        check(lib().gopf_dist_real_step(self._h, self._p(W)))

    def forward_mid(self, W, send):
        check(lib().gopf_dist_forward_mid(self._h, self._p(W), self._p(send)))

    def kspace_step(self, T, S):
        check(lib().gopf_dist_kspace_step(self._h, self._p(T), self._p(S)))

    def inverse_finish(self, W, real_out):
        check(lib().gopf_dist_inverse_finish(self._h, self._p(W), self._p(real_out)))

    def advance(self):
        check(lib().gopf_dist_advance(self._h))

    def get_time(self) -> float:
        t = ctypes.c_double(0.0)
        check(lib().gopf_dist_solver_get_time(self._h, ctypes.byref(t)))
        return t.value

    def kernel_launches(self, reset: bool = False) -> int:
        n = ctypes.c_int64(0)
        check(lib().gopf_dist_solver_kernel_launches(self._h, ctypes.byref(n), 1 if reset else 0))
        return n.value

    def close(self):
        if self._h:
            lib().gopf_dist_solver_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedSolver:
    """pf.Solver for a slab-sharded cubic grid.  ``model`` is a gopf_b200.pf.Model whose single
    field's host ``Data`` holds THIS rank's slab (planes [rank*n/world, (rank+1)*n/world))."""

    def __init__(self, model, n: int, dt: float, device: int, group=None):
        import torch
        import torch.distributed as tdist
        self.torch, self.tdist, self.group = torch, tdist, group
        self.world = tdist.get_world_size(group)
        self.rank = tdist.get_rank(group)
        self.n, self.dt, self.model = n, dt, model
        self.device = torch.device("cuda", device)
        self.phases = CudaPhases(model, n, self.world, self.rank, dt, device)
        cells = self.phases.local_cells
        mk = lambda: torch.empty(cells, dtype=torch.complex128, device=self.device)
        self.S, self.A, self.B = mk(), mk(), mk()
        self.a_valid = False
        self.on_device = False
        # One real stream for kernels, copies and the collective's stream dependencies.  (The
        # legacy default stream has handle 0, which the C ABI reads as "use the plan's stream".)
        self.stream = torch.cuda.Stream(device=self.device)
        self.phases.set_stream(self.stream.cuda_stream)

    def all_to_all(self, dst, src):
        if self.world == 1:
            dst.copy_(src)
            return
        self.tdist.all_to_all_single(self.torch.view_as_real(dst), self.torch.view_as_real(src), group=self.group)

    def Upload(self):
        with self.torch.cuda.stream(self.stream):
            host = self.torch.from_numpy(self.model.Fields[0].Data)
            self.A.copy_(host, non_blocking=True)
            upload(self.phases, self.all_to_all, self.A, self.S, self.B)
        self.a_valid = False
        self.on_device = True

    def StepDevice(self, nsteps: int):
        with self.torch.cuda.stream(self.stream):
            self.a_valid = run_steps(self.phases, self.all_to_all, self.S, self.A, self.B, nsteps, self.a_valid)

    def Download(self):
        with self.torch.cuda.stream(self.stream):
            download(self.phases, self.all_to_all, self.S, self.A, self.B, self.A)  # last pass in place on A
            host = self.torch.from_numpy(self.model.Fields[0].Data)
            host.copy_(self.A)
        self.stream.synchronize()
        self.a_valid = False

    def Synchronize(self):
        self.stream.synchronize()

    def Propagate(self, nsteps: int):
        self.Upload()
        self.StepDevice(nsteps)
        self.Download()
