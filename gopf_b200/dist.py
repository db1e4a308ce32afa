"""Slab-sharded stepping: one process per GPU, torch.distributed for the exchange.

Layout and phase contract: include/gopf_cuda.h ("slab-sharded step") and
gopf_b200/csrc/dist_solver.h.  ``run_steps`` / ``upload`` / ``download`` below are
the whole orchestration; they only need an object with the phase methods and an
``all_to_all(dst, src)`` callable, so the same code runs over NCCL with the CUDA
phases (``CudaPhases``) and, in the CPU tests, over gloo with a numpy stand-in.
"""
from __future__ import annotations

import ctypes

from ._lib import GopfError, check, lib


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


def chunks(extent: int, nchunks: int):
    """[(begin, count)] covering range(extent) in at most nchunks near-equal pieces."""
    nchunks = max(1, min(nchunks, extent))
    base, rem = divmod(extent, nchunks)
    out, b = [], 0
    for i in range(nchunks):
        c = base + (1 if i < rem else 0)
        out.append((b, c))
        b += c
    return out


def run_steps_peer(phases, barrier, S, A, nsteps: int, x_valid: bool, slab: int = 0, nchunks: int = 1,
                   comm_ctas: int = 0) -> bool:
    """nsteps sharded Euler steps with the exchange fused into the producing passes (peer
    stores over NVLink).  The receive buffers X and Y live in ``phases``; ``barrier()`` orders
    every rank's earlier stream work before every rank's later stream work.  ``x_valid`` says X
    already holds the first inverse pass of S.  Returns the new x_valid.

    nchunks > 1 pipelines the real-space side by plane chunks: the peer-storing pass of chunk c
    (NVLink-bound; second stream, confined to ``comm_ctas`` SMs) runs under the HBM-bound
    inverse_mid / real_step kernels of chunk c+1."""
    parts = chunks(slab, nchunks) if nchunks > 1 and slab > 0 else None
    for _ in range(nsteps):
        if not x_valid:
            barrier()                        # nobody is still reading X
            phases.inverse_start_peer(S)     # S -> inverse axis 0 -> every rank's X
            barrier()
        if parts is None:
            phases.inverse_mid_x(A)          # X -> inverse axis 1 -> A
            phases.real_step(A)              # inverse axis 2, /N, g(c), forward axis 2
            phases.forward_mid_peer(A)       # forward axis 1 -> every rank's Y
        else:
            for i, (b, c) in enumerate(parts):
                # from the second chunk on the peer-storing pass of the previous chunk holds comm_ctas SMs
                phases.set_grid_cap(comm_ctas if i > 0 else 0, reserve=True)
                phases.inverse_mid_planes_x(A, b, c)
                phases.real_step_planes(A, b, c)
                phases.forward_mid_peer_planes(A, b, c, comm_ctas)
            phases.set_grid_cap(0)
            phases.exchange_join()
        barrier()
        phases.kspace_step_peer(S)           # Y: forward axis 0, Euler update of S, inverse axis 0 -> every rank's X
        barrier()
        phases.advance()
        x_valid = True
    return x_valid


def run_steps_dma(phases, barrier, S, A, SEND, nsteps: int, x_valid: bool, slab: int, nchunks: int = 4) -> bool:
    """nsteps sharded Euler steps with the exchange on the copy engines, pipelined by chunks:
    the DMA copies of chunk c (second stream, peer-mapped X / Y) run under the kernels of chunk
    c+1.  ``slab`` = n / world (planes per rank = spectrum columns k1l per rank)."""
    parts = chunks(slab, nchunks)
    for _ in range(nsteps):
        if not x_valid:
            barrier()
            phases.inverse_start(S, SEND)            # S -> inverse axis 0 -> SEND [k0][k1l][k2]
            phases.exchange_inverse(SEND, 0, slab)
            phases.exchange_join()
            barrier()
        for b, c in parts:                           # planes are independent up to the exchange
            phases.inverse_mid_planes_x(A, b, c)     # X -> inverse axis 1 -> A
            phases.real_step_planes(A, b, c)
            phases.forward_mid_planes(A, SEND, b, c)
            phases.exchange_forward(SEND, b, c)      # -> every rank's Y, under the next chunk's kernels
        phases.exchange_join()
        barrier()
        for b, c in parts:                           # columns are independent in k-space
            phases.kspace_step_cols_y(S, A, b, c)    # Y, S -> S; inverse axis 0 of the new S -> A
            phases.exchange_inverse(A, b, c)         # -> every rank's X
        phases.exchange_join()
        barrier()
        phases.advance()
        x_valid = True
    return x_valid


def upload_dma(phases, barrier, real_slab, S, SEND, slab: int):
    phases.forward_local(real_slab, SEND)
    phases.exchange_forward(SEND, 0, slab)
    phases.exchange_join()
    barrier()
    phases.forward_finish_peer(S)


def download_dma(phases, barrier, S, A, SEND, real_out, x_valid: bool, slab: int) -> bool:
    if not x_valid:
        barrier()
        phases.inverse_start(S, SEND)
        phases.exchange_inverse(SEND, 0, slab)
        phases.exchange_join()
        barrier()
    phases.inverse_mid_x(A)
    phases.inverse_finish(A, real_out)
    return True


def upload_peer(phases, barrier, real_slab, S):
    """real_slab [i0l][i1][i2] (destroyed) -> S = its share of the spectrum."""
    phases.forward_local_peer(real_slab)     # forward axes 2, 1 -> every rank's Y
    barrier()
    phases.forward_finish_peer(S)            # Y -> forward axis 0 -> S


def download_peer(phases, barrier, S, A, real_out, x_valid: bool) -> bool:
    """S -> real_out = this rank's slab of the real-space field.  X stays valid."""
    if not x_valid:
        barrier()
        phases.inverse_start_peer(S)
        barrier()
    phases.inverse_mid_x(A)
    phases.inverse_finish(A, real_out)
    return True


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
        try:
            import torch
            self.sm_count = int(torch.cuda.get_device_properties(device).multi_processor_count)
        except Exception:
            self.sm_count = 148

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

    # ---- peer-store exchange -------------------------------------------------------------
    def peer_alloc(self):
        check(lib().gopf_dist_peer_alloc(self._h))

    def peer_export(self, which: int) -> bytes:
        buf = ctypes.create_string_buffer(64)
        check(lib().gopf_dist_peer_export(self._h, int(which), buf))
        return buf.raw

    def peer_import(self, which: int, rank: int, handle: bytes):
        if len(handle) != 64:
            raise ValueError("CUDA IPC handle must be 64 bytes")
        check(lib().gopf_dist_peer_import(self._h, int(which), int(rank), ctypes.c_char_p(handle)))

    def peer_local(self, which: int) -> int:
        p = ctypes.c_void_p()
        check(lib().gopf_dist_peer_local(self._h, int(which), ctypes.byref(p)))
        return p.value

    def inverse_start_peer(self, S):
        check(lib().gopf_dist_inverse_start_peer(self._h, self._p(S)))

    def inverse_mid_x(self, W):
        check(lib().gopf_dist_inverse_mid(self._h, ctypes.c_void_p(self.peer_local(0)), self._p(W)))

    def forward_mid_peer(self, W):
        check(lib().gopf_dist_forward_mid_peer(self._h, self._p(W)))

    def forward_local_peer(self, W):
        check(lib().gopf_dist_forward_local_peer(self._h, self._p(W)))

    def forward_finish_peer(self, S):
        check(lib().gopf_dist_forward_finish_peer(self._h, self._p(S)))

    def kspace_step_peer(self, S):
        check(lib().gopf_dist_kspace_step_peer(self._h, self._p(S)))

    # ---- chunked phases + copy-engine exchange ----------------------------------------------
    def inverse_mid_planes_x(self, W, begin: int, count: int):
        check(lib().gopf_dist_inverse_mid_planes(self._h, ctypes.c_void_p(self.peer_local(0)), self._p(W), int(begin), int(count)))

    def real_step_planes(self, W, begin: int, count: int):
        check(lib().gopf_dist_real_step_planes(self._h, self._p(W), int(begin), int(count)))

    def forward_mid_planes(self, W, send, begin: int, count: int):
        check(lib().gopf_dist_forward_mid_planes(self._h, self._p(W), self._p(send), int(begin), int(count)))

    def kspace_step_cols_y(self, S, Tout, k1_begin: int, k1_count: int):
        check(lib().gopf_dist_kspace_step_cols(self._h, ctypes.c_void_p(self.peer_local(1)), self._p(S), self._p(Tout),
                                               int(k1_begin), int(k1_count)))

    def exchange_forward(self, send, begin: int, count: int):
        check(lib().gopf_dist_exchange_forward(self._h, self._p(send), int(begin), int(count)))

    def exchange_inverse(self, T, k1_begin: int, k1_count: int):
        check(lib().gopf_dist_exchange_inverse(self._h, self._p(T), int(k1_begin), int(k1_count)))

    def exchange_join(self):
        check(lib().gopf_dist_exchange_join(self._h))

    def set_grid_cap(self, ctas: int, reserve: bool = False):
        """Persistent compute-stream kernels launch at most ``ctas`` CTAs (0: one per SM); with ``reserve`` the
        argument is the number of SMs to leave free instead."""
        if reserve and ctas > 0:
            ctas = max(1, self.sm_count - ctas)
        check(lib().gopf_dist_set_grid_cap(self._h, int(ctas)))

    def peer_unmap(self):
        check(lib().gopf_dist_peer_unmap(self._h))

    def forward_mid_peer_planes(self, W, begin: int, count: int, max_ctas: int):
        check(lib().gopf_dist_forward_mid_peer_planes(self._h, self._p(W), int(begin), int(count), int(max_ctas)))

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

    def __init__(self, model, n: int, dt: float, device: int, group=None, exchange: str = "peer", nchunks: int = 8,
                 comm_ctas: int = 48):
        """exchange = "peer": the transpose is fused into the producing passes as stores into
        the peers' receive buffers over NVLink (CUDA IPC mappings), with a tiny NCCL all-reduce
        as the stream barrier; "dma": the same receive buffers filled by copy-engine copies
        pipelined under the kernels, ``nchunks`` chunks per exchange; "nccl": pack in the pass,
        ``all_to_all_single``, unpack in the next pass (the baseline)."""
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
        if exchange not in ("peer", "dma", "nccl"):
            raise ValueError("exchange must be 'peer', 'dma' or 'nccl'")
        self.exchange = exchange
        self.nchunks = nchunks
        self.comm_ctas = comm_ctas   # peer: SMs given to the NVLink-bound pass while it runs under the next chunk
        self.slab = n // self.world
        self.S, self.A = mk(), mk()
        self.B = mk() if exchange in ("nccl", "dma") else None
        self.a_valid = False   # nccl: A holds the first inverse pass of S; peer: X does
        self.on_device = False
        if exchange in ("peer", "dma") and not self._map_peers():
            self.exchange = "nccl"   # same kernels, classic all-to-all between the phases
            if self.B is None:
                self.B = mk()
        # One real stream for kernels, copies and the collective's stream dependencies.  (The
        # legacy default stream has handle 0, which the C ABI reads as "use the plan's stream".)
        self.stream = torch.cuda.Stream(device=self.device)
        self.phases.set_stream(self.stream.cuda_stream)

    def _map_peers(self) -> bool:
        """Allocate this rank's receive buffers and map every other rank's (CUDA IPC handles
        travel through one all_gather).  Returns False -- on EVERY rank -- when any rank could not
        export or map a buffer (e.g. a container without CUDA IPC), so that all ranks switch to the
        NCCL exchange together; the reason goes to stderr."""
        import sys
        torch, tdist = self.torch, self.tdist
        mine, err = bytes(128), None
        try:
            self.phases.peer_alloc()
            mine = self.phases.peer_export(0) + self.phases.peer_export(1)
        except GopfError as e:
            err = e
        if self.world > 1:
            send = torch.tensor(list(mine) + [0 if err is None else 1], dtype=torch.uint8, device=self.device)
            recv = [torch.empty_like(send) for _ in range(self.world)]
            tdist.all_gather(recv, send, group=self.group)
            rows = [bytes(t.cpu().tolist()) for t in recv]
        else:
            rows = [mine + bytes([0 if err is None else 1])]
        if err is None and not any(r[128] for r in rows):
            try:
                for q, h in enumerate(rows):
                    if q != self.rank:
                        self.phases.peer_import(0, q, h[:64])
                        self.phases.peer_import(1, q, h[64:128])
            except GopfError as e:
                err = e
        ok = torch.tensor([0 if (err is not None or any(r[128] for r in rows)) else 1], dtype=torch.int32, device=self.device)
        if self.world > 1:
            tdist.all_reduce(ok, op=tdist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            if err is not None:
                print(f"gopf_b200.dist: rank {self.rank}: peer mapping failed ({err}); all ranks use the NCCL exchange",
                      file=sys.stderr, flush=True)
            return False
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        return True

    def barrier(self):
        """Cross-rank barrier in stream order: the all-reduce cannot complete on any rank
        before every rank has reached it on its stream, i.e. finished its earlier kernels
        (whose peer stores are complete at kernel end)."""
        if self.world > 1:
            self.tdist.all_reduce(self._flag, group=self.group)

    def all_to_all(self, dst, src):
        if self.world == 1:
            dst.copy_(src)
            return
        self.tdist.all_to_all_single(self.torch.view_as_real(dst), self.torch.view_as_real(src), group=self.group)

    def Upload(self):
        with self.torch.cuda.stream(self.stream):
            host = self.torch.from_numpy(self.model.Fields[0].Data)
            self.A.copy_(host, non_blocking=True)
            if self.exchange == "peer":
                upload_peer(self.phases, self.barrier, self.A, self.S)
            elif self.exchange == "dma":
                upload_dma(self.phases, self.barrier, self.A, self.S, self.B, self.slab)
            else:
                upload(self.phases, self.all_to_all, self.A, self.S, self.B)
        self.a_valid = False
        self.on_device = True

    def StepDevice(self, nsteps: int):
        with self.torch.cuda.stream(self.stream):
            if self.exchange == "peer":
                self.a_valid = run_steps_peer(self.phases, self.barrier, self.S, self.A, nsteps, self.a_valid, self.slab,
                                              self.nchunks if self.world > 1 else 1, self.comm_ctas)
            elif self.exchange == "dma":
                self.a_valid = run_steps_dma(self.phases, self.barrier, self.S, self.A, self.B, nsteps, self.a_valid,
                                             self.slab, self.nchunks)
            else:
                self.a_valid = run_steps(self.phases, self.all_to_all, self.S, self.A, self.B, nsteps, self.a_valid)

    def Download(self):
        with self.torch.cuda.stream(self.stream):
            if self.exchange == "peer":
                self.a_valid = download_peer(self.phases, self.barrier, self.S, self.A, self.A, self.a_valid)
            elif self.exchange == "dma":
                self.a_valid = download_dma(self.phases, self.barrier, self.S, self.A, self.B, self.A, self.a_valid, self.slab)
            else:
                download(self.phases, self.all_to_all, self.S, self.A, self.B, self.A)  # last pass in place on A
                self.a_valid = False
            host = self.torch.from_numpy(self.model.Fields[0].Data)
            host.copy_(self.A)
        self.stream.synchronize()

    def Synchronize(self):
        self.stream.synchronize()

    def Propagate(self, nsteps: int):
        self.Upload()
        self.StepDevice(nsteps)
        self.Download()

    def close(self):
        """Collective tear-down: every rank closes its mappings of the other ranks' receive buffers, a barrier,
        then each rank frees its own (an exported allocation must outlive its importers' mappings)."""
        if self.phases is None:
            return
        self.torch.cuda.synchronize(self.device)
        if self.exchange in ("peer", "dma"):
            self.phases.peer_unmap()
            if self.world > 1:
                self.tdist.barrier(group=self.group)
        self.phases.close()
        self.phases = None
        self.S = self.A = self.B = None
