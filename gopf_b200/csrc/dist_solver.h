// DistSolver: slab-sharded single-field Euler step for cubic 3-D grids, one rank per GPU
// (SURVEY.md 8e).  The reference has no distributed path; this is the B200-native
// extension of the same step (pf/euler.go:16-47) to grids that do not fit, or do not run
// fast enough on, one GPU.
//
// Rank p of P owns planes i0 in [p*m, (p+1)*m), m = n/P, of the real-space array
// (contiguous in the reference's node numbering).  k-space lives TRANSPOSED: rank p owns
// k1 in [p*m, (p+1)*m) for every k0, k2, laid out [k0][k1_local][k2].  One all-to-all per
// distributed transform; the exchange itself is done by the caller (torch.distributed /
// NCCL) between the phases below, on buffers it owns.  Pack and unpack are folded into
// the passes on either side of the exchange through split row maps (fft_kernels.cuh).
//
//   send layout (forward)  [q][i0_local][k1_local(q)][k2]  -> all-to-all -> [k0][k1_local][k2]
//   [k0][k1_local][k2] = [q][k0_local(q)][k1_local][k2]    -> all-to-all -> [p][i0_local][k1_local(p)][k2]
#pragma once
#include <memory>

#include "fft_plan.h"
#include "fused_launch.h"
#include "model.h"
#include "solver.h"

namespace gopf {

class DistSolver {
public:
    DistSolver(Model* m, int n, int world, int rank, double dt, int device);

    int n() const { return n_; }
    int slab() const { return m_; }
    size_t local_cells() const { return (size_t)m_ * n_ * n_; }
    void set_stream(cudaStream_t s) { stream_ = s; }
    long long kernel_launches() const { return launches_; }
    void reset_launch_count() { launches_ = 0; }

    // ---- phases (all arrays: local_cells() complex128 on this rank's device) --------
    // upload: real slab W [i0l][i1][i2] -> forward axes 2, 1 -> send layout
    void forward_local(cplx* W, cplx* send);
    // after the exchange: forward axis 0 of T = [k0][k1l][k2], in place (T becomes the spectrum)
    void forward_finish(cplx* T);
    // first inverse pass of the spectrum: S -> T (axis 0)
    void inverse_start(const cplx* S, cplx* T);
    // after the exchange: inverse axis 1, recv (split layout) -> W [i0l][k1][k2]
    void inverse_mid(const cplx* recv, cplx* W);
    // last inverse pass, /N, derived function, first forward pass; W in place
    void real_step(cplx* W);
    // forward axis 1: W -> send layout
    void forward_mid(const cplx* W, cplx* send);
    // after the exchange: finish forward (axis 0) of T, Euler update of S, inverse axis 0 -> T
    void kspace_step(cplx* T, cplx* S);
    // download: last inverse pass + /N of W -> real slab
    void inverse_finish(cplx* W, cplx* real_out);
    // ---- peer-store exchange (B200: NVLink 5 / NVSwitch peer memory) ---------------------
    // Each rank owns two receive buffers, X ([p][i0l][k1l(p)][k2], input of inverse_mid) and
    // Y ([k0][k1l][k2], input of kspace_step / forward_finish).  Every rank maps every other
    // rank's X and Y (CUDA IPC) and the pass that produces exchange data stores each row
    // directly into its owner's buffer, so pack + send + unpack cost no pass and no separate
    // collective.  The caller separates a peer-writing phase from the phases that read the
    // buffers with a cross-rank barrier on the stream (gopf_b200/dist.py).
    void peer_alloc();                                         // cudaMalloc X, Y
    void peer_export(int which, void* handle64) const;         // which: 0 = X, 1 = Y
    void peer_import(int which, int rank, const void* handle64);
    cplx* peer_local(int which) const { return which == 0 ? X_[rank_] : Y_[rank_]; }
    bool peer_ready() const;
    // Close the mappings of the other ranks' buffers (the caller then runs a cross-rank barrier before any rank
    // frees its own: an exported allocation must outlive its importers' mappings).
    void peer_unmap();
    // compute-stream persistent kernels launch at most `ctas` CTAs from now on (0: no cap); see PassGeom::grid_cap
    void set_grid_cap(int ctas) { grid_cap_ = ctas < 0 ? 0 : ctas; }
    void inverse_start_peer(const cplx* S);   // S -> inverse axis 0 -> peers' X
    void forward_mid_peer(const cplx* W);     // W -> forward axis 1 -> peers' Y
    void kspace_step_peer(cplx* S);           // local Y, S -> S; inverse axis 0 of the new S -> peers' X
    void forward_local_peer(cplx* W);         // upload: forward axis 2 in place, then forward_mid_peer
    void forward_finish_peer(cplx* S);        // upload: local Y -> forward axis 0 -> S
    // ---- copy-engine exchange, pipelined by chunks ----------------------------------------
    // The same X / Y receive buffers, filled by DMA copies (cudaMemcpyAsync to the IPC-mapped
    // peer pointers) on a second stream while the compute stream works on the next chunk:
    // planes of the slab are independent through inverse_mid -> real_step -> forward_mid, and
    // columns (k1l) of the spectrum are independent through kspace_step.  The copy engines
    // reach the NVLink DMA rate and occupy no SM.
    void inverse_mid_planes(const cplx* recv, cplx* W, int begin, int count);
    void real_step_planes(cplx* W, int begin, int count);
    void forward_mid_planes(const cplx* W, cplx* send, int begin, int count);
    void kspace_step_cols(const cplx* Tin, cplx* S, cplx* Tout, int k1_begin, int k1_count);
    void exchange_forward(const cplx* send, int begin, int count);    // send planes -> every rank's Y (after the compute so far)
    void exchange_inverse(const cplx* T, int k1_begin, int k1_count);  // T columns -> every rank's X
    void exchange_join();                                              // compute stream waits for the copies issued so far
    // Peer-store exchange pipelined by plane chunks: forward axis 1 of planes [begin, begin+count)
    // -> peers' Y, launched on the second (high-priority) stream as a persistent kernel confined to
    // `max_ctas` SMs, so the NVLink-bound pass runs under the HBM-bound kernels of the next chunk.
    void forward_mid_peer_planes(const cplx* W, int begin, int count, int max_ctas);
    ~DistSolver();
    void advance() { steps_taken_++; current_step_++; }
    double get_time() const { return (double)current_step_ * dt_; }

private:
    Model* m_model_;
    int n_, world_, rank_, m_;
    double dt_;
    std::unique_ptr<FftPlan> plan_;  // tables and twiddles for length n
    cudaStream_t stream_ = nullptr;
    DevKProgram prog_;
    int derived_ = -1;
    long long steps_taken_ = 0, current_step_ = 0, launches_ = 0;
    int grid_cap_ = 0;
    // the uploaded slab had no imaginary part (checked on the device in forward_local*): the real-space kernel may
    // pair lines (PassGeom::real_pairs, tma_kernels.cuh); the answer is collected at the first real-space step
    bool field_real_ = false, real_check_pending_ = false;
    double* d_imag_max_ = nullptr;
    double* h_imag_max_ = nullptr;
    void check_real(const cplx* W);
    int real_pairs();
    cplx* X_[GOPF_MAX_PEERS];  // [rank]: mapped receive buffers (own rank: owned allocation)
    cplx* Y_[GOPF_MAX_PEERS];
    PeerOut peer_out(cplx* const* bufs, bool kspace_rows) const;
    cudaStream_t copy_stream_ = nullptr;
    cudaEvent_t ev_compute_ = nullptr, ev_copy_ = nullptr;
    void copy_after_compute();

    cudaStream_t stream() const { return stream_ ? stream_ : plan_->stream; }
    PassGeom slab_axis1(bool split_in, bool split_out) const;
    RowMap split_map() const;
    void check(cudaError_t e, const char* what);
};

}  // namespace gopf
