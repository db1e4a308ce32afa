// Solver: device-resident counterpart of pf.Solver + pf.Euler / pf.RK4
// (/root/reference/pf/solver.go:29-120, pf/euler.go:16-47, pf/rk4.go:29-127).
//
// State between host synchronisations is the k-space spectrum of every field,
// resident in HBM.  The reference re-transforms c every step (euler.go:19-21);
// FFT(IFFT(c^)/N) == c^ to rounding, so the persistent spectrum is algebraically
// the same step with one transform fewer (T_min = 2 for
// Cahn-Hilliard).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "fft_plan.h"
#include "jit.h"
#include "elastic.cuh"
#include "model.h"
#include "step_kernels.cuh"

namespace gopf {

enum class StepperKind { Euler, RK4, ImplicitEuler, SDD };

// Settings of the Newton-Krylov solve inside ImplicitEuler.Step (pf/implicitEuler.go:221-229
// DefaultNonLinSolver; oracle/pf.py NewtonKrylov states the algorithm).
struct NewtonKrylovOptions {
    int maxiter = 50;
    double step_size = 1e-3;
    double tol = 1e-7;
    int stencil = 6;       // central-difference points for J v: 2, 4 or 6 (DefaultNonLinSolver: 6, implicitEuler.go:226)
    int restart = 30;      // GMRES restart length
    double inner_tol = 1e-4;
    int max_restarts = 4;
};

// pf.SDD (pf/sdd.go:86-129): settings and monitor of the shrinking-dimer stepper
struct SddState {
    double alpha = 0.5;             // SDD.Alpha
    double tau_orientation = 1.0;   // TimeConstants.Orientation
    double tau_dimer_length = 1.0;  // TimeConstants.DimerLength
    double dt = 0.0;                // SDD.Dt: must be set explicitly (checkTimeStep, :131-135)
    double min_dimer_length = 0.0;
    double init_dimer_length = 0.0;
    long long current_step = 0;
    bool initialized = false;
    // SDDMonitor (:25-53)
    double max_force = 0.0, force_power_spectrum = 0.0, max_torque = 0.0, field_norm = 0.0, field_norm_change = 0.0;
};

// shared by Solver and DistSolver (solver.cu)
// allow_table: admit the tabulated form (DevKProgram::fast == 2); the caller then owns and fills P.dtab
void finalize_single_field_program(DevKProgram* prog, int n_fields, bool allow_table = false);
int single_field_derived_index(const Model& m);

struct KernelTimer {  // CUDA-event timing of one kernel class, accumulated over launches
    std::string name;
    double bytes_per_launch = 0.0;  // algorithmic HBM bytes
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    double total_ms = 0.0;
    long long launches = 0;
};

bool program_has_knoise(const DevKProgram& P);

class Solver {
public:
    Solver(Model* m, int rank, const int* n, double dt, int device);
    ~Solver();
    Solver(const Solver&) = delete;
    Solver& operator=(const Solver&) = delete;

    void set_stepper(const std::string& name);   // Solver.SetStepper (solver.go:88-103)
    void set_filter(const double* table, int n);  // TimeStepper.SetFilter with a tabulated ModalFilter
    void set_stream(cudaStream_t s) {
        user_stream_ = s;
        drop_graph();
    }
    void upload();              // host Field.Data -> device spectra
    void step(int nsteps);      // nsteps x Stepper.Step on device-resident state
    void download();            // device spectra -> real-space host Field.Data
    // Real part of field i only, optionally byte-swapped to big-endian on the device: the payload of
    // Field.SaveReal / Float64IO.SaveFields (pf/model.go:35-41, pf/fileIO.go:57-62) at half the D2H bytes
    void download_real(int field, double* host_out, bool big_endian);
    void propagate(int nsteps)  // Solver.Propagate on host buffers (solver.go:70-84)
    {
        upload();
        step(nsteps);
        download();
    }
    double get_time() const {
        if (stepper_ == StepperKind::SDD) return (double)sdd_.current_step * sdd_.dt;  // sdd.go:359-361
        return (double)current_step_ * dt_;
    }
    // pf.SDD (sdd.cu): SetInitialOrientation (sdd.go:413-427), settings / monitor by name, orientation
    void sdd_set_orientation(const double* orient, long long len);
    void sdd_set(const std::string& key, double value);
    double sdd_get(const std::string& key);
    void sdd_get_orientation(double* host_out);
    bool fused() const { return fused_; }
    void fused_form(int* form, int* derived_form);
    long long kernel_launches() const { return launches_; }
    void reset_launch_count() { launches_ = 0; }
    void set_profiling(bool on);
    const std::vector<KernelTimer>& collect_profile();
    cplx* spectrum(int i) {
        leave_blocked();
        return S_.s[i];
    }
    // Blocked k-space layout of the fused path on large 3-D grids (solver.cu): 0 = off,
    // s > 0 = on with blocks of 2^s axis-0 positions.  blocked_now(): the field spectrum currently sits in it.
    int blocked_log() const;
    bool blocked_now() const { return blocked_; }
    FftPlan& plan() { return *plan_; }
    void synchronize();
    void force_generic(bool on);
    double lp_multiplier(int slot);
    // ChargeTransport.Current (chargeTransport.go:121-146) of the term registered as `name`:
    // host_out[d*N + i] = -real(current_d[i]) from the device-resident spectrum
    void charge_current(const std::string& name, double* host_out);
    // IdealMixtureTerm.GetEnergy / PairCorrlationTerm.GetEnergy (pairCorrelationTerm.go:58-84, 185-193)
    double term_energy(const std::string& name);
    // Uint8IO.SaveFields payload (fileIO.go:29-44, util.go:108-117) + pfutil.MinReal / MaxReal
    void download_uint8(int field, unsigned char* host_out, double* mn, double* mx);
    void set_newton_krylov(const NewtonKrylovOptions& o) { nk_ = o; }
    bool last_step_converged() const { return ie_converged_; }
    // run-time specialisation (jit.h): switch, and how many kernels (registered functions, the
    // k-space update) currently run as compiled images
    void set_jit(bool on) { jit_on_ = on; }
    void set_jit_inpass(bool on) { jit_inpass_ = on; }
    int jit_kernels() const;
    const std::string& jit_log() const { return jit_log_; }
    long long residual_evaluations() const { return ie_residual_evals_; }

private:
    Model* m_;
    double dt_;
    std::unique_ptr<FftPlan> plan_;
    StepperKind stepper_ = StepperKind::Euler;
    long long current_step_ = 0;   // Euler.CurrentStep (RK4.Step never advances it: rk4.go:130-135)
    long long steps_taken_ = 0;    // counter for the noise stream
    cudaStream_t user_stream_ = nullptr;
    bool on_device_ = false;
    bool fused_ = false, allow_fused_ = true;
    bool w_valid_ = false;  // fused path: W holds the first inverse pass of the current spectrum
    // fused path: the uploaded field had no imaginary part (checked on the device at upload); with a real fast-form
    // program the real-space kernel may then pair lines (PassGeom::real_pairs)
    bool field_real_ = false, real_check_pending_ = false;
    double* d_imag_max_ = nullptr;
    double* h_imag_max_ = nullptr;  // page-locked
    void resolve_real_check();
    int fused_derived_ = -1;
    long long launches_ = 0;

    SpectraPtrs S_;            // [fields | derived | work] spectra
    RealPtrs R_;               // real-space fields (generic path)
    cplx* Rw_[GOPF_MAX_FIELDS];
    cplx* W_ = nullptr;        // fused work array
    // Blocked k-space layout (large 3-D grids, fused path).  Lines along axis 0 of a row-major [n0][n1][n2]
    // array have a row stride of n1*n2 cells: at 1024^3 every one of the 1024 rows of a tile sits in its own
    // 2-MB page 16 MB from the next, and the copy engine reads such tiles at 3.0 TB/s against 6.2 TB/s for the
    // 16-KB stride of the middle axis (scripts/tune_blocked.py).  So while fused steps run, the field spectrum
    // S and the work array between the middle-axis passes and the k-space kernel (W2_) are kept as
    // [n0 / 2^s][n1][2^s][n2]: 2^s consecutive axis-0 positions of one (i1, *) line are n2 cells apart.  The
    // k-space kernel then reads tiles made of n0/2^s runs of 2^s rows at the short stride (6.1 TB/s at s = 7),
    // and the middle-axis passes convert on the fly (row-major on the real-space side, blocked on the k-space
    // side) at no extra traffic.  Everything else sees the row-major S: leave_blocked() converts back.
    cplx* W2_ = nullptr;
    double* fused_dtab_ = nullptr;  // tabulated single-field form: filter / (1 - dt*den) per k-point
    bool blocked_ = false;
    int block_log_ = -1;  // decided at first use (-1: not yet)
    void enter_blocked();
    void leave_blocked();
    PassGeom blocked_axis0_geom() const;           // lines along axis 0 of a blocked array, tiles by (i1, i2)
    PassGeom blocked_axis1_geom(bool in_blocked, bool out_blocked) const;
    double* d_real_out_ = nullptr;  // staging for download_real
    double* d_filter_ = nullptr;
    int filter_n_ = 0;
    double* d_lp_state_ = nullptr;  // per VolumeConservingLP: multiplier, current integral, first flag
    cplx* rk_initial_[GOPF_MAX_FIELDS];
    cplx* rk_final_[GOPF_MAX_FIELDS];
    cplx* rk_k_[GOPF_MAX_FIELDS];
    cplx* sg_tmp_[3];
    cplx* implicit_tab_[GOPF_MAX_FIELDS];  // filter / (1 - dt*den) per field, when expensive per k (solver.cu)
    bool implicit_tab_dirty_ = true;
    // HomogeneousModulusLinElast (pf/homoLinElast.go): tabulated multiplier M(k) per term slot,
    // RK4's snapshot of the real-space field, and "OnStepFinished has run at least once"
    double* elast_mtab_[GOPF_MAX_SPECIAL];
    cplx* elast_phi_[GOPF_MAX_SPECIAL];
    bool elast_valid_ = false;
    double* d_table_[GOPF_MAX_SPECTRA];
    // ChargeTransport (catalog_terms.cu): conductivity tables per term slot, one current component
    double* ct_sigma_[GOPF_MAX_SPECIAL] = {nullptr, nullptr};
    cplx* ct_tmp_ = nullptr;
    double* obs_partial_ = nullptr;
    void observe(const cplx* a, const cplx* b, int mode, const double* q, double* sum, double* mn, double* mx);
    DevKProgram prog_;
    DevKProgram fused_prog_;
    bool prog_dirty_ = true;
    bool has_knoise_ = false;  // some term draws white noise at the k-point (model.h kspace_noise)

    // CUDA-graph replay of the fused step on small grids (launch-bound: 2-4 kernels of a few us)
    cudaGraphExec_t graph_exec_ = nullptr;
    int graph_steps_ = 0;
    bool graph_disabled_ = false;
    void drop_graph();
    bool graph_applicable() const;
    bool build_graph(int steps);

    // Registered functions compiled to straight-line sm_100a code at first use (jit.h); a NULL
    // entry keeps the interpreter kernel for that derived field
    std::vector<jit::Kernel*> jit_derived_;
    std::vector<char> jit_tried_;
    // ... or compiled into the load of their first forward pass (GOPF_JIT_INPASS)
    std::vector<jit::Kernel*> jit_pass_;
    std::vector<char> jit_pass_tried_;
    bool jit_inpass_ = false;
    jit::Kernel* jit_pass_kernel(int d, int N);
    bool jit_on_ = false;
    std::string jit_log_;
    void derived_pointwise(int d, cplx* out, unsigned long long step_no, cudaStream_t s);
    // k-space kernels compiled for the current program (one image, three entry points); NULL: generic kernels
    jit::Kernel* jit_kupdate_ = nullptr;
    jit::Kernel* jit_rk4_rhs_ = nullptr;
    jit::Kernel* jit_rk4_point_ = nullptr;
    std::string jit_kupdate_key_;
    bool ensure_jit_program(const DevKProgram& P, unsigned tab_mask);
    bool launch_update_jit(const DevKProgram& P, const ImplicitTab& tab);

    NewtonKrylovOptions nk_;
    bool ie_converged_ = true;
    long long ie_residual_evals_ = 0;
    cplx* ie_orig_[GOPF_MAX_FIELDS];      // spectra at the start of the step
    cplx* ie_rhs_prev_[GOPF_MAX_FIELDS];  // RHS at the start of the step
    cplx* ie_res_[GOPF_MAX_FIELDS];       // residual work spectra
    std::vector<double*> ie_vec_;         // x, F(x), b, s, w, tmp_p, tmp_m, then restart+1 Krylov vectors
    double* ie_partial_ = nullptr;        // reduction partials (device) 
    int ie_vec_restart_ = -1;

    // SDD (sdd.cu)
    SddState sdd_;
    double* sdd_orient_ = nullptr;           // orientation vector, F x N reals
    cplx* sdd_vhat_[GOPF_MAX_FIELDS] = {};   // FFT of the orientation blocks
    cplx* sdd_rs_[GOPF_MAX_FIELDS] = {};     // RHS at the start image, then the torque
    cplx* sdd_re_[GOPF_MAX_FIELDS] = {};     // RHS at the end image
    cplx* sdd_work_ = nullptr;
    cplx* sdd_d_ = nullptr;
    double* sdd_partial_ = nullptr;
    void sdd_step();
    void sdd_ensure_buffers();
    void sdd_free_buffers();
    double sdd_dimer_length(double t) const;
    void sdd_collect(int n_sums, unsigned blocks, double* sums, double* mx);

    bool profiling_ = false;
    std::vector<KernelTimer> timers_;

    cudaStream_t stream() const { return user_stream_ ? user_stream_ : plan_->stream; }
    void ensure_buffers();
    void rebuild_program();
    void decide_path();
    void euler_step_generic();
    void euler_update_generic();   // the step without the OnStepFinished hooks
    // ImplicitEuler (implicit_euler.cu)
    void implicit_euler_step();
    void ie_residual(const double* x, double* out);
    void ie_jac_vec(const double* x, const double* v, double* out);
    void ie_gmres(const double* x, const double* b, double* s);
    void ie_ensure_buffers();
    double ie_dot(const double* a, const double* b);
    double ie_max_abs(const double* a);
    void euler_step_fused();
    void rk4_step();
    void inverse_to_real(const cplx* spec, cplx* real_out);           // IFFT + /N, out of place
    void forward_derived(int d);                                       // derived d -> its spectrum
    void forward_in_place(cplx* data);
    void eval_real_fields();
    void squared_gradient_terms();
    void elastic_terms();
    void catalog_terms();            // ChargeTransport + point sources (catalog_terms.cu)
    void charge_transport_terms();
    void source_terms();
    void free_catalog_buffers();
    template <class Emit>
    void charge_current_components(const UserTerm& u, const cplx* rho, Emit emit);
    void elastic_hooks();
    bool has_elastic() const;
    void volume_lp_hooks();
    void launch_update(const DevKProgram& P);
    int tick(const char* name, double bytes);
    void tock(int id);
    FreqTabs freq_tabs() const;
};

}  // namespace gopf
