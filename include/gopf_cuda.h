/*
 * gopf_cuda.h -- C ABI of libgopfcuda.so, the sm_100a implementation of gopf's
 * spectral time-stepping hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes only, no torch / C++
 * types.  Every entry point names the reference interface it replaces
 * (file:line relative to the davidkleiven/gopf tree).  The Go side binds these
 * through cgo (see INTEGRATION.md and go/); tests and bench.py bind them through
 * ctypes.
 *
 * Conventions
 *   - complex128 arrays are interleaved (re, im) doubles, exactly Go's
 *     []complex128 / FFTW's fftw_complex; `double*` arguments named *_c128 point
 *     at 2*N doubles.
 *   - every function returns 0 on success and non-zero on failure;
 *     gopf_last_error() then describes the failure (thread local).  The
 *     reference panics on these paths (pf/solver.go:58, pf/rhsBuilder.go:29,147,
 *     pf/model.go:51); the Go shim turns a non-zero status back into panic().
 *   - handles are not thread-safe, matching FFTWWrapper (shared Data buffer,
 *     pfutil/fftWrap.go:8-13) and the single-goroutine reference.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with a non-zero status.
 */
#ifndef GOPF_CUDA_H
#define GOPF_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOPF_ABI_VERSION 1

typedef struct gopf_fft_plan gopf_fft_plan;
typedef struct gopf_model gopf_model;
typedef struct gopf_solver gopf_solver;

/* ---- library ------------------------------------------------------------- */
const char* gopf_last_error(void);
int gopf_abi_version(void);
/* number of visible CUDA devices (0 and status != 0 when the driver is absent) */
int gopf_device_count(int* count);
/* Launches that took the copy-engine-fed (TMA, warp-specialised) line kernels since the last reset:
 * the long-line variants of the axis passes and of the fused kernels behind gopf_fft_exec /
 * gopf_solver_step (tma_kernels.cuh; GOPF_TMA=0 in the environment keeps the register-resident
 * kernels).  Diagnostics for tests and the bench line; no reference counterpart. */
int gopf_tma_launch_count(int reset, int64_t* launches);

/* ---- pfutil index / k-table helpers (host side, integer-exact) ------------- */
/* pfutil.NodeIdx (pfutil/indexPositionConversion.go:4-22); pos = [row, col(, depth)] */
int gopf_node_idx(int rank, const int* domain_size, const int* pos, int64_t* node);
/* pfutil.Pos (pfutil/indexPositionConversion.go:24-44) */
int gopf_pos(int rank, const int* domain_size, int64_t node, int* pos_out);
/* FFTWWrapper.Freq (pfutil/fftWrap.go:57-74): out[rank] cycles/sample, Nyquist stays +0.5 */
int gopf_freq(int rank, const int* n, int64_t i, double* out);
/* FFTWWrapper.ConjugateNode (pfutil/fftWrap.go:78-95) */
int gopf_conjugate_node(int rank, const int* n, int64_t i, int64_t* out);

/* ---- transform level: pfutil.FFTWWrapper ---------------------------------- */
/* pfutil.NewFFTW(n) (pfutil/fftWrap.go:16-23).  rank 1..3, row-major, last axis
 * fastest.  device < 0 selects the current device. */
int gopf_fft_plan_create(int rank, const int* n, int device, gopf_fft_plan** out);
/* FFTWWrapper.FFT / IFFT (pfutil/fftWrap.go:26-39): in place on the caller's host
 * slice, unnormalised both ways.  sign = -1 forward, +1 inverse. */
int gopf_fft_exec(gopf_fft_plan* plan, double* host_inout_c128, int sign);
/* same on a device-resident array; stream is a cudaStream_t (NULL = plan stream) */
int gopf_fft_exec_device(gopf_fft_plan* plan, void* dev_inout_c128, int sign, void* stream);
/* one axis pass only (0 = slowest axis of the normalised 3-axis view), device arrays, out may
 * equal in; tile_cells = strided-tile width in cells (0: plan default).  Kernel tuning and tests. */
int gopf_fft_exec_axis_device(gopf_fft_plan* plan, const void* dev_in_c128, void* dev_out_c128, int sign, int axis,
                              int tile_cells, void* stream);
/* One strided line pass with explicit tile geometry (layout experiments, tests of the blocked k-space layout):
 * `slabs` x `cols` lines of the plan's extent along `axis`; cell (a, j, b) sits at element offset
 *   (a >> map[5]) * map[4] + (a & ((1 << map[5]) - 1)) * map[0] + (j >> map[3]) * map[2] + (j & ((1 << map[3]) - 1)) * map[1] + b
 * with map = {slab_stride, row_stride, row_split_stride, row_split_log (>= 31: none), slab_split_stride,
 * slab_split_log (>= 31: none)}, separately for input and output. */
int gopf_fft_exec_rows_device(gopf_fft_plan* plan, const void* dev_in_c128, void* dev_out_c128, int sign, int axis,
                              int64_t slabs, int64_t cols, const int64_t* in_map, const int64_t* out_map, int tile_cells,
                              void* stream);
/* Freq(i) for `count` node numbers evaluated ON THE DEVICE by the same code the
 * k-space kernels use (bit-exactness check of the device k-table). */
int gopf_fft_freq_device(gopf_fft_plan* plan, const int64_t* nodes, int64_t count, double* out);
int gopf_fft_plan_destroy(gopf_fft_plan* plan);

/* ---- gradient-based catalog terms at operator level ----------------------------------------
 * GradientCalculator, DivGrad, WeightedLaplacian (pf/gradientCalculator.go:10-172) and Advection
 * (pf/advection.go:17-98).  As shipped the reference cannot register these with a model (their
 * OnStepFinished(t) lacks the `bricks` argument of pf.PureTerm, pf/userDefinedTerm.go:35); they are used by
 * calling PrepareModel / Construct by hand, and these calls are that surface on a transform plan.  Arrays are
 * N complex128 (N = product of the plan's dimensions); the *_device forms take device pointers on the plan's
 * device and a stream (NULL: the plan's), the others host pointers (Field.Data / DerivedField.Data).
 *
 * gopf_gradient_calculate: GradientCalculator{FT, Comp, KeepNyquist}.Calculate(in, out)
 *   (gradientCalculator.go:19-31): out = IFFT(i 2 pi f_comp FFT(in)) / N, f = +1/2 zeroed unless keep_nyquist.
 * gopf_advection_construct: Advection{Field, VelocityFields}: what Construct's closure leaves in `field`
 *   (advection.go:87-94) after the derived fields of PrepareModel (:50-85) are up to date:
 *   -sum_d velocity[d] * GRAD_d(field), in real space (transformed = 0, as the reference's test reads it) or
 *   forward-transformed (transformed = 1, as a step sees derived fields).  n_velocity must equal the rank.
 * gopf_div_grad_construct: DivGrad{Field, F}.Construct (gradientCalculator.go:96-108) with the derived fields of
 *   PrepareModel (:72-93): sum_d i 2 pi f_d FFT(func_values * GRAD_d(field)); func_values = F(i, bricks) per node.
 * gopf_weighted_laplacian_construct: WeightedLaplacian.Construct (gradientCalculator.go:131-172):
 *   FFT( IFFT(L field_hat)/N * IFFT(prefactor_hat)/N ), L = -(2 pi |f|)^2; both inputs are spectra. */
int gopf_gradient_calculate(gopf_fft_plan* plan, const double* in, double* out, int comp, int keep_nyquist);
int gopf_gradient_calculate_device(gopf_fft_plan* plan, const void* in, void* out, int comp, int keep_nyquist, void* stream);
int gopf_advection_construct(gopf_fft_plan* plan, const double* field, const double* const* velocity, int n_velocity,
                             double* out, int transformed);
int gopf_advection_construct_device(gopf_fft_plan* plan, const void* field, const void* const* velocity, int n_velocity,
                                    void* out, int transformed, void* stream);
int gopf_div_grad_construct(gopf_fft_plan* plan, const double* field, const double* func_values, double* out);
int gopf_div_grad_construct_device(gopf_fft_plan* plan, const void* field, const void* func_values, void* out, void* stream);
int gopf_weighted_laplacian_construct(gopf_fft_plan* plan, const double* field_hat, const double* prefactor_hat, double* out);
int gopf_weighted_laplacian_construct_device(gopf_fft_plan* plan, const void* field_hat, const void* prefactor_hat, void* out,
                                             void* stream);

/* ---- step level: pf.Model ---------------------------------------------------
 * The model records what the Go API calls describe and compiles it, at
 * gopf_model_init / gopf_solver_create, into device programs.  Arbitrary Go
 * closures cannot run on the device: functions are given as expressions, terms
 * come from the reference's catalog.  Anything else fails here, loudly. */
/* pf.NewModel (pf/model.go:130-138) */
int gopf_model_create(gopf_model** out);
/* The C source the registered function `name` compiles to: `static inline double gopf_expr(r0, i0,
 * ..., r3, i3)` over the real / imaginary parts of fields 0..3 (kernel = 0), or the whole CUDA
 * translation unit handed to NVRTC (kernel = 1).  *needed = bytes including the terminator; buf
 * may be NULL to query.  Host only, no GPU. */
int gopf_model_function_source(gopf_model* m, const char* name, int kernel, char* buf, int64_t len, int64_t* needed);
/* The same for the forward pass with the function in its load (gopf_solver_set_jit_inpass) at line
 * length `line_length`; the compile also reports the mangled name of the kernel instance. */
int gopf_model_function_pass_source(gopf_model* m, const char* name, int line_length, char* buf, int64_t len, int64_t* needed);
int gopf_model_function_pass_compile(gopf_model* m, const char* name, int line_length, int64_t* cubin_bytes, char* lowered_name,
                                     int lowered_len);
/* NVRTC-compile that translation unit for sm_100a; *cubin_bytes = size of the image.  No GPU needed. */
int gopf_model_function_compile(gopf_model* m, const char* name, int64_t* cubin_bytes);
/* The CUDA translation unit the k-space update of this model (pf/euler.go:27-39) is specialised to
 * on an n[0] x n[1] (x n[2]) grid: the compiled term list as a constant image, grid geometry and
 * node count as literals.  tab_mask bit i: field i uses a tabulated implicit factor.  filter_addr /
 * filter_n: address and length of the modal-filter table to bake in (0: no filter); lp_addr: address
 * of the VolumeConservingLP state, three doubles per slot (0: none).  The solver passes its device
 * addresses; these entry points exist for inspection, compile checks and tests/host_emul (which
 * compiles the unit for the host and passes host addresses).  Calls Model.Init.  Host only, no GPU. */
int gopf_model_kupdate_source(gopf_model* m, int rank, const int* n, double dt, unsigned tab_mask, uint64_t filter_addr,
                              int filter_n, uint64_t lp_addr, char* buf, int64_t len, int64_t* needed);
int gopf_model_kupdate_compile(gopf_model* m, int rank, const int* n, double dt, unsigned tab_mask, uint64_t filter_addr,
                               int filter_n, uint64_t lp_addr, int64_t* cubin_bytes);
/* WhiteNoise fields (pf/noise.go:20-23) that enter an equation as a plain explicit term: draw their
 * spectrum directly at each k-point (Hermitian, E|xi^|^2 = N * 2 * Strength) instead of generating the
 * field in real space and transforming it every step -- one transform less per step; a model with one
 * field, one nonlinearity and such a noise term then takes the fused kernels.  The reference's random
 * stream is unpinned, parity is statistical either way.  Off by default; needs a library whose device
 * code was built with -DGOPF_KNOISE (gopf_solver_create fails otherwise).  Call before gopf_solver_create. */
int gopf_model_set_kspace_noise(gopf_model* m, int on);
/* 1 when the device code of this library carries the k-space noise generator (-DGOPF_KNOISE), else 0 */
int gopf_has_kspace_noise(void);
/* Raw images of what the model compiles to: the k-space program (struct DevKProgram of
 * gopf_b200/csrc/step_program.h, device pointers NULL) and derived field `index` (struct DevDerived;
 * index counts derived fields in registration order, spectrum index = number of fields + index).
 * For inspection and for tests/host_emul, which compiles the same headers for the host and runs the
 * evaluators against the oracle without a GPU.  *needed = size of the struct; buf may be NULL.
 * Both call Model.Init.  Host only. */
int gopf_model_program_image(gopf_model* m, int rank, double dt, void* buf, int64_t len, int64_t* needed);
int gopf_model_derived_image(gopf_model* m, int index, void* buf, int64_t len, int64_t* needed, int* used);
/* The program as the fused single-field kernels take it (spectrum 0 = the field, 1 = its one derived field,
 * real-polynomial fast form when it applies; solver.cu finalize_single_field_program, no filter set) and the
 * index of that derived field.  Fails when the model is not of that shape. */
int gopf_model_fused_program_image(gopf_model* m, int rank, double dt, void* buf, int64_t len, int64_t* needed, int* derived_index);
/* pf.NewField + Model.AddField (pf/model.go:44-57, 141-144).  host_c128 is the
 * caller-owned Field.Data backing array of n_nodes complex128; it is read by
 * gopf_solver_upload/propagate and written by gopf_solver_download/propagate,
 * never retained past the model's lifetime and never freed. */
int gopf_model_add_field(gopf_model* m, const char* name, int64_t n_nodes, double* host_c128);
/* pf.NewScalar + Model.AddScalar (pf/model.go:95-101, 147-149) */
int gopf_model_add_scalar(gopf_model* m, const char* name, double re, double im);
/* Model.AddEquation (pf/model.go:157-162), e.g. "dconc/dt = LAP conc^3 + m1*LAP conc" */
int gopf_model_add_equation(gopf_model* m, const char* equation);
/* Model.RegisterFunction (pf/model.go:400-412) with the GenericFunction given as a
 * real-valued expression over real parts of fields and scalars:
 * + - * / ^ ( ), H dH Landau dLandau exp log sin cos tanh sqrt abs negpart re() im();
 * negpart(x) = min(x, 0) expresses NegativeValuePenalty.Evaluate (pf/negative_value_penalty.go:15-27) */
int gopf_model_register_function(gopf_model* m, const char* name, const char* expression);
/* RegisterFunction(name, WhiteNoise{Strength}.Generate) (pf/noise.go:11-23): N(0, 2*Strength)
 * per node per step from a counter-based Philox stream (seed, step, node) */
int gopf_model_register_white_noise(gopf_model* m, const char* name, double strength, uint64_t seed);
/* RegisterDerivedField with prescribed real values: values[n_steps][n_nodes], step s reads
 * row s mod n_steps.  Lets a parity test inject one noise array into oracle and device. */
int gopf_model_register_table_field(gopf_model* m, const char* name, const double* values, int64_t n_steps);
/* RegisterImplicitTerm(name, &SpectralViscosity{Eps, DissipationThreshold, Power}) (pf/spectralViscosity.go:23-54) */
int gopf_model_register_spectral_viscosity(gopf_model* m, const char* name, double eps, double threshold, int power);
/* RegisterImplicitTerm(name, &PairCorrlationTerm{...}) / RegisterExplicitTerm(name,
 * &ExplicitPairCorrelationTerm{...}) (pf/pairCorrelationTerm.go:22-51, 89-110; pfc/pairCorrelation.go:8-37) */
int gopf_model_register_pair_correlation(gopf_model* m, const char* name, int explicit_term, const char* field,
                                         double prefactor, int laplacian, double eff_temp, int n_peaks,
                                         const double* plane_density, const double* location, const double* width,
                                         const int* num_planes);
/* RegisterMixedTerm(name, &IdealMixtureTerm{IdealMix{C3,C4}, Field, Prefactor, Laplacian}, dfields)
 * (pf/pairCorrelationTerm.go:116-178); register_derived != 0 also registers
 * IdealMixtureTerm.DerivedField (ideal_mixture_<field>_nonlin) */
int gopf_model_register_ideal_mixture(gopf_model* m, const char* name, const char* field, double c3, double c4,
                                      double prefactor, int laplacian, int register_derived);
/* RegisterExplicitTerm(name, &ConservativeNoise{UniquePrefix, Strength, Dim}, RequiredDerivedFields(N))
 * (pf/noise.go:25-100) */
int gopf_model_register_conservative_noise(gopf_model* m, const char* name, double strength, int dim,
                                           uint32_t unique_prefix, uint64_t seed);
/* same term over caller-registered current fields named "<unique_prefix>_current_<c>" */
int gopf_model_register_conservative_noise_term(gopf_model* m, const char* name, int dim, uint32_t unique_prefix);
/* RegisterExplicitTerm(name, &VolumeConservingLP{Field, Indicator, Dt}) (pf/volumeConserving.go:3-61) */
int gopf_model_register_volume_conserving_lp(gopf_model* m, const char* name, const char* field,
                                             const char* indicator, double dt);
/* RegisterExplicitTerm(name, &SquaredGradient{Field, Factor}) (pf/squareGradientTerm.go:14-68) */
int gopf_model_register_squared_gradient(gopf_model* m, const char* name, const char* field, double factor);
/* RegisterImplicitTerm(name, &TensorialHessian{Field, K}) (pf/tensorialHessian.go:17-74):
 * sum_ij K_ij d_i d_j acting on the equation's own field; K row-major, 4 (2-D) or 9 (3-D) values */
int gopf_model_register_tensorial_hessian(gopf_model* m, const char* name, const char* field, const double* k, int n_coeff);
/* RegisterExplicitTerm(name, NewHomogeneousModolus(fieldName, domainSize, matProp, misfit))
 * (pf/homoLinElast.go:30-150).  stiffness81 = elasticity.Rank4.Data (index i*27+j*9+k*3+l,
 * elasticity/rank4.go:10-24), misfit9 = the 3x3 misfit strain, row-major.  The term's private
 * real-space copy of the field starts as zeros and is refreshed after every step
 * (OnStepFinished, :130-134), so the term vanishes during the first step -- replicated. */
int gopf_model_register_homogeneous_modulus_lin_elast(gopf_model* m, const char* name, const char* field,
                                                      const double* stiffness81, const double* misfit9);
/* RegisterExplicitTerm(name, &ChargeTransport{Conductivity, ExternalField, Field, FT})
 * (pf/chargeTransport.go:29-119): minus the divergence of the current j = sigma (E_ind - E_ext) with
 * the induced field from Poisson's equation in k-space.  The Go closure Conductivity(i) is
 * tabulated by the caller: conductivity[v * n_nodes + i] = Conductivity(i)[v], n_voigt = 3 in 2-D
 * (s_xx, s_yy, s_xy) or 6 in 3-D (s_xx, s_yy, s_zz, s_xz, s_yz, s_xy); the table is copied.
 * external_field holds n_ext = rank values. */
int gopf_model_register_charge_transport(gopf_model* m, const char* name, const char* field,
                                         const double* conductivity, int n_voigt, int64_t n_nodes,
                                         const double* external_field, int n_ext);
/* Model.AddSource(eqNo, pf.NewSource(pos, f)) (pf/model.go:151-154, pf/sourceTerm.go:10-30): adds
 * f(t) * exp(-i 2 pi Freq(k) . pos) to the right-hand side of equation eq_no.  f is the reference's
 * TimeDepSource as a C callback; it is called on the host once per right-hand-side evaluation
 * with t = TimeStepper.GetTime() and `user`.  pos holds n_pos >= rank coordinates. */
typedef double (*gopf_time_fn)(double t, void* user);
int gopf_model_add_source(gopf_model* m, int eq_no, const double* pos, int n_pos, gopf_time_fn f, void* user);
/* Host evaluation of the per-k factors the ChargeTransport / Source kernels apply (same
 * __host__ __device__ code, gopf_b200/csrc/catalog_terms.cuh), for `count` frequency vectors
 * freq[count][rank]: field_mult[c][i] (pf/chargeTransport.go:64-73), div_mult[c][i] (:106-112),
 * component-major [rank][count]; either output may be NULL. */
int gopf_charge_transport_multipliers(int rank, const double* freq, int64_t count, double* field_mult,
                                      double* div_mult);
/* voigtIndex(i, j, dim) (pf/chargeTransport.go:151-171) */
int gopf_charge_transport_voigt_index(int i, int j, int dim, int* out);
/* Source.Eval on the host (pf/sourceTerm.go:25-30): out_c128[i] = amp * exp(-i 2 pi freq[i] . pos) */
int gopf_source_eval(int rank, const double* freq, int64_t count, const double* pos, double amp, double* out_c128);
/* elasticity.CubicMaterial / Isotropic / Rank4.Rotate / Rank4.ContractLast / EnergyDensity
 * (elasticity/rank4.go:39-128, linearElasticity.go:86-98): host-side tensor helpers */
int gopf_elasticity_cubic_material(double c11, double c12, double c44, double* out81);
int gopf_elasticity_isotropic(double bulk_mod, double poisson, double* out81);
int gopf_elasticity_rotate(double* inout81, const double* rot9);
int gopf_elasticity_contract_last(const double* stiffness81, const double* tensor9, double* out9);
int gopf_elasticity_energy_density(const double* stiffness81, const double* strain9, double* out);
/* elasticity.HomogeneousModulusEnergy(indicator, domainSize, misfit, matProp)
 * (elasticity/linearElasticity.go:101-165) on the device: elastic energy per unit precipitate volume of
 * the inclusion `indicator_c128` (host, N complex128, not modified; the reference transforms it in place
 * and back).  device = -1: current device.  Post-processing, not a step kernel. */
int gopf_elasticity_homogeneous_modulus_energy(int rank, const int* n, const double* indicator_c128,
                                               const double* misfit9, const double* stiffness81, int device,
                                               double* energy);
/* the real per-k factor s_ij(k) with eps^_ij = s_ij H^ used by that call (Displacements + Strain,
 * elasticity/linearElasticity.go:16-83), evaluated on the host for `count` padded frequency vectors
 * freq3[count][3] by the same __host__ __device__ code */
int gopf_elasticity_strain_factor(const double* stiffness81, const double* misfit9, const double* freq3, int64_t count,
                                  int i, int j, double* out);
/* The real multiplier M(k) the device applies between the transforms of the elastic term
 * (gopf_b200/csrc/elastic.cuh), evaluated on the host for `count` frequency triples
 * [f_row, f_col, f_depth] -- parity tests against elasticity.Displacements + Strain. */
int gopf_elasticity_multiplier(const double* stiffness81, const double* misfit9, int dim, const double* freq3,
                               int64_t count, double* out);
/* Model.Init (pf/model.go:244-260): parse + classify every term */
int gopf_model_init(gopf_model* m);
int gopf_model_num_fields(gopf_model* m, int* n);
int gopf_model_num_derived_fields(gopf_model* m, int* n);
int gopf_model_derived_field_name(gopf_model* m, int index, char* buf, int buf_len);
/* len(m.RHS[eq].Terms), len(m.RHS[eq].Denum) after Init (pf/rhsBuilder.go:19-22) */
int gopf_model_num_terms(gopf_model* m, int eq, int* n_terms, int* n_denum);
/* Model.EqNumber (pf/model.go:441-455) */
int gopf_model_eq_number(gopf_model* m, const char* field_name, int* eq);
int gopf_model_destroy(gopf_model* m);

/* pf.NewVandeven(order).Data (pf/vandeven.go:13-27): fills out[1000] */
int gopf_vandeven_table(int order, double* out, int n);

/* ---- step level: pf.Solver / pf.TimeStepper ------------------------------------ */
/* pf.NewSolver(m, domainSize, dt) (pf/solver.go:40-62): Init()s the model, plans the
 * transforms, default stepper Euler.  device < 0: current device. */
int gopf_solver_create(gopf_model* m, int rank, const int* domain_size, double dt, int device, gopf_solver** out);
/* Solver.SetStepper("euler" | "rk4") (pf/solver.go:88-103) */
int gopf_solver_set_stepper(gopf_solver* s, const char* name);
/* solver.Stepper = &pf.ImplicitEuler{Dt, FT} (pf/implicitEuler.go:20-229) is selected with
 * gopf_solver_set_stepper(s, "implicit_euler").  The non-linear solve is a Jacobian-free
 * Newton-Krylov iteration on the device (gopf_b200/csrc/implicit_euler.cu); the reference
 * delegates it to third-party modules that are not in its tree, so trajectories agree with it to
 * the solver tolerance only.  Defaults = DefaultNonLinSolver (implicitEuler.go:221-229:
 * Maxiter 50, StepSize 1e-3, Tol 1e-7, Stencil 6); GMRES restarts after `restart` vectors. */
int gopf_solver_set_newton_krylov(gopf_solver* s, int maxiter, double step_size, double tol, int stencil, int restart,
                                  double inner_tol, int max_restarts);
/* after a step: res.Converged of the last solve (the reference logs a warning when false,
 * implicitEuler.go:202-204) and the residual evaluations so far */
int gopf_solver_newton_krylov_status(gopf_solver* s, int* converged, int64_t* residual_evaluations);
/* TimeStepper.SetFilter for a tabulated ModalFilter (pf/util.go:120-132, pf/vandeven.go:30-40);
 * table == NULL removes the filter */
int gopf_solver_set_filter(gopf_solver* s, const double* table, int n);
/* run on the caller's cudaStream_t instead of the solver's own stream */
int gopf_solver_set_stream(gopf_solver* s, void* stream);
/* Solver.Propagate(nsteps) (pf/solver.go:70-84) on the host Field.Data arrays:
 * upload, nsteps x Stepper.Step (+ OnStepFinished hooks), download */
int gopf_solver_propagate(gopf_solver* s, int nsteps);
/* the same three phases separately, for device-resident stepping between callbacks */
int gopf_solver_upload(gopf_solver* s);
int gopf_solver_step(gopf_solver* s, int nsteps);
int gopf_solver_download(gopf_solver* s);
int gopf_solver_synchronize(gopf_solver* s);
/* TimeStepper.GetTime (pf/euler.go:50-52, pf/rk4.go:143-145) */
int gopf_solver_get_time(gopf_solver* s, double* t);
/* Blocked k-space layout of the fused path on large 3-D grids: *block_log = s when fused steps keep
 * the spectrum as [n0 / 2^s][n1][2^s][n2] between host synchronisations (0: row-major throughout), *active = 1
 * while the device spectrum currently sits in that layout.  Diagnostics and tests; every entry point that
 * exposes the spectrum converts back first. */
int gopf_solver_blocked_layout(gopf_solver* s, int* block_log, int* active);
/* How the fused single-field kernels evaluate the model (diagnostics, tests).  *form: 0 = not fused or the term
 * interpreter, 1 = polynomial form (every term a monomial with a real coefficient), 2 = tabulated form (polynomial
 * explicit side, optional k-space white noise; implicit side and modal filter tabulated once per k-point).
 * *derived_form: 0 = interpreter, 1 = integer power of the field, 2 = real polynomial of re(field). */
int gopf_solver_fused_form(gopf_solver* s, int* form, int* derived_form);
/* 1 when the single-field fused kernels are in use, 0 for the general path */
int gopf_solver_is_fused(gopf_solver* s, int* fused);
int gopf_solver_force_generic(gopf_solver* s, int on);
/* Run-time specialisation of registered functions (Model.RegisterFunction, pf/model.go:400-412):
 * the expression is compiled with NVRTC into a straight-line sm_100a kernel at first use instead
 * of being interpreted per cell.  On by default since round 2 (GOPF_JIT=0 in the environment when the solver is
 * created switches it off); gopf_solver_set_jit overrides that.  A function that fails to compile keeps the
 * interpreter kernel (both are device paths); gopf_solver_jit_log tells why. */
int gopf_solver_set_jit(gopf_solver* s, int on);
/* With the specialisation on, also compile each registered function into the load of the first
 * forward pass of its transform (a copy of the library's contiguous-axis pass kernel with a generated
 * loader): no pointwise kernel, no round trip of the function values.  On by default; GOPF_JIT_INPASS=0 switches it off. */
int gopf_solver_set_jit_inpass(gopf_solver* s, int on);
/* number of derived fields currently evaluated by compiled kernels */
int gopf_solver_jit_kernels(gopf_solver* s, int* count);
int gopf_solver_jit_log(gopf_solver* s, char* buf, int len);
/* kernels launched by this solver since creation / since the last reset */
int gopf_solver_kernel_launches(gopf_solver* s, int64_t* n, int reset);
/* k-space spectrum of field / derived field `index` -> host (debug / tests) */
int gopf_solver_get_spectrum(gopf_solver* s, int index, double* host_c128);
/* VolumeConservingLP.Multiplier of the slot-th registered term */
int gopf_solver_lp_multiplier(gopf_solver* s, int slot, double* value);
/* pf.SDD, the shrinking-dimer saddle-point stepper (pf/sdd.go:86-470), selected with
 * gopf_solver_set_stepper(s, "sdd") (the reference assigns solver.Stepper = &sdd; selecting it is
 * NewSDD: Alpha 0.5, time constants 1, orientation zeros, not initialised).
 * sdd_set_orientation = SDD.SetInitialOrientation (:413-427): n_fields * n_nodes reals, normalised,
 * their length becomes InitDimerLength; SDD.Init(init, final) (:390-408) is the same call with
 * real(final - init).  sdd_set / sdd_get address the struct's exported fields by their Go names:
 * "Alpha", "Dt" (must be set: a step with Dt < 1e-16 fails like the reference's panic),
 * "TimeConstants.Orientation", "TimeConstants.DimerLength", "MinDimerLength", "InitDimerLength",
 * "CurrentStep"; sdd_get also "DimerLength" (at GetTime()) and the SDDMonitor fields
 * "Monitor.MaxForce", "Monitor.ForcePowerSpectrum", "Monitor.MaxTorque", "Monitor.FieldNorm",
 * "Monitor.FieldNormChange" (:25-53).  A modal filter is refused (:431-433). */
int gopf_solver_sdd_set_orientation(gopf_solver* s, const double* orientation, int64_t len);
int gopf_solver_sdd_get_orientation(gopf_solver* s, double* host_out);
int gopf_solver_sdd_set(gopf_solver* s, const char* key, double value);
int gopf_solver_sdd_get(gopf_solver* s, const char* key, double* value);
/* ChargeTransport.Current(density, N, realspace) (pf/chargeTransport.go:121-146) for the term
 * registered as `name`, evaluated on the device-resident spectrum of its field:
 * host_out[d * N + i] = -real(current_d[i]), d < rank */
int gopf_solver_charge_current(gopf_solver* s, const char* name, double* host_out);
/* Real part of field `field_index` (N doubles) from the device-resident state, big-endian byte
 * order on request: the payload of Field.SaveReal / Float64IO.SaveFields (pf/model.go:35-41,
 * pf/fileIO.go:57-62, 85-95) for epoch callbacks of device-resident runs, at half the D2H bytes of
 * gopf_solver_download.  The swap to big endian runs on the device. */
int gopf_solver_download_real(gopf_solver* s, int field_index, double* host_out, int big_endian);
/* Uint8IO.SaveFields payload of one field (pf/fileIO.go:29-44): pfutil.MinReal / MaxReal of the
 * device-resident field (returned in min_real / max_real, either may be NULL) and RealPartAsUint8
 * (pf/util.go:108-117), uint8(255 * (re - min) / (max - min)), N bytes -- 1/16 of the D2H bytes of
 * gopf_solver_download.  This is the epoch callback of examples/cahnHilliard (config 1). */
int gopf_solver_download_uint8(gopf_solver* s, int field_index, uint8_t* host_out, double* min_real, double* max_real);
/* IdealMixtureTerm.GetEnergy (pf/pairCorrelationTerm.go:185-193) or PairCorrlationTerm.GetEnergy
 * (:58-84) of the term registered as `name`, evaluated on the device-resident state (the energy
 * observer of examples/pfcPhases, config 5) */
int gopf_solver_term_energy(gopf_solver* s, const char* name, double* energy);
/* per-kernel CUDA-event timing over the following gopf_solver_step calls */
int gopf_solver_profile_begin(gopf_solver* s);
int gopf_solver_profile_end(gopf_solver* s, int* n_kernels);
int gopf_solver_profile_get(gopf_solver* s, int i, char* name, int name_len, double* total_ms, int64_t* launches,
                            double* bytes_per_launch);
int gopf_solver_destroy(gopf_solver* s);

/* ---- slab-sharded step: one rank per GPU ------------------------------------------
 * No reference counterpart (gopf is single process); this extends pf.Euler.Step
 * (pf/euler.go:16-47) to a cubic n^3 grid split into `world` slabs of n/world planes
 * along the slowest axis.  The library runs the local phases; the caller performs the
 * all-to-all between them (NCCL through torch.distributed, or ncclSend/ncclRecv from Go)
 * on device buffers it owns, each n/world * n * n complex128.  Order of one step, with
 * work buffers A, B and spectrum S (see gopf_b200/dist.py and DESIGN.md):
 *   [first step only: inverse_start(S -> A)]
 *   all_to_all(A -> B); inverse_mid(B -> A); real_step(A); forward_mid(A -> B);
 *   all_to_all(B -> A); kspace_step(A, S); advance()
 * The model must be the single-field fused form (one field holding this rank's slab,
 * one equation, one nonlinear derived field). */
typedef struct gopf_dist_solver gopf_dist_solver;
int gopf_dist_solver_create(gopf_model* m, int n, int world, int rank, double dt, int device, gopf_dist_solver** out);
int gopf_dist_solver_set_stream(gopf_dist_solver* s, void* stream);
int gopf_dist_solver_local_cells(gopf_dist_solver* s, int64_t* cells);
/* upload side: real slab W -> forward axes 2,1 -> send layout; after the exchange forward axis 0 */
int gopf_dist_forward_local(gopf_dist_solver* s, void* w_c128, void* send_c128);
int gopf_dist_forward_finish(gopf_dist_solver* s, void* t_c128);
int gopf_dist_inverse_start(gopf_dist_solver* s, const void* spectrum_c128, void* t_c128);
int gopf_dist_inverse_mid(gopf_dist_solver* s, const void* recv_c128, void* w_c128);
int gopf_dist_real_step(gopf_dist_solver* s, void* w_c128);
int gopf_dist_forward_mid(gopf_dist_solver* s, const void* w_c128, void* send_c128);
int gopf_dist_kspace_step(gopf_dist_solver* s, void* t_c128, void* spectrum_c128);
/* download side: last inverse pass and 1/N, W -> real slab */
int gopf_dist_inverse_finish(gopf_dist_solver* s, void* w_c128, void* real_out_c128);
/* Peer-store exchange (B200 NVLink 5 / NVSwitch): the transpose is fused into the pass that
 * produces the data.  Every rank owns receive buffers X (which = 0; read by inverse_mid) and
 * Y (which = 1; read by kspace_step_peer / forward_finish_peer), exports them as 64-byte CUDA
 * IPC handles and imports every other rank's; the *_peer phases then store each row straight
 * into its owner's buffer.  The caller places a cross-rank barrier on the stream between a
 * peer-writing phase and the phases that read the buffers.  One step:
 *   [first step only: barrier; inverse_start_peer(S); barrier]
 *   inverse_mid(X_local -> A); real_step(A); forward_mid_peer(A); barrier;
 *   kspace_step_peer(S); barrier; advance()                                              */
int gopf_dist_peer_alloc(gopf_dist_solver* s);
int gopf_dist_peer_export(gopf_dist_solver* s, int which, void* handle64);
int gopf_dist_peer_import(gopf_dist_solver* s, int which, int rank, const void* handle64);
/* Close this rank's mappings of the other ranks' buffers.  Tear-down order: every rank unmaps, a cross-rank
 * barrier, then gopf_dist_solver_destroy frees the rank's own buffers (an exported allocation must outlive
 * its importers' mappings). */
int gopf_dist_peer_unmap(gopf_dist_solver* s);
/* The persistent (copy-engine-fed) kernels of the chunked compute-stream phases launch at most `ctas` CTAs
 * from now on (0: one per SM).  Set to SMs - max_ctas of forward_mid_peer_planes while that pass runs on the
 * second stream, so the statically partitioned tiles are not queued behind it. */
int gopf_dist_set_grid_cap(gopf_dist_solver* s, int ctas);
int gopf_dist_peer_local(gopf_dist_solver* s, int which, void** dev_ptr);
int gopf_dist_inverse_start_peer(gopf_dist_solver* s, const void* spectrum_c128);
int gopf_dist_forward_mid_peer(gopf_dist_solver* s, const void* w_c128);
int gopf_dist_forward_local_peer(gopf_dist_solver* s, void* w_c128);
int gopf_dist_forward_finish_peer(gopf_dist_solver* s, void* spectrum_c128);
int gopf_dist_kspace_step_peer(gopf_dist_solver* s, void* spectrum_c128);
/* Copy-engine exchange pipelined by chunks (same X / Y buffers and IPC mappings): planes
 * [begin, begin+count) of the slab are independent through inverse_mid -> real_step ->
 * forward_mid, columns k1l in [k1_begin, k1_begin+k1_count) through kspace_step, so the DMA
 * copies of one chunk (second stream, no SM) run under the kernels of the next.  One step:
 *   for each plane chunk: inverse_mid_planes(X_local -> A); real_step_planes(A);
 *                         forward_mid_planes(A -> SEND); exchange_forward(SEND)
 *   exchange_join; barrier
 *   for each column chunk: kspace_step_cols(Y_local, S -> T); exchange_inverse(T)
 *   exchange_join; barrier; advance()                                                     */
int gopf_dist_inverse_mid_planes(gopf_dist_solver* s, const void* recv_c128, void* w_c128, int begin, int count);
int gopf_dist_real_step_planes(gopf_dist_solver* s, void* w_c128, int begin, int count);
int gopf_dist_forward_mid_planes(gopf_dist_solver* s, const void* w_c128, void* send_c128, int begin, int count);
int gopf_dist_kspace_step_cols(gopf_dist_solver* s, const void* t_in_c128, void* spectrum_c128, void* t_out_c128,
                               int k1_begin, int k1_count);
int gopf_dist_exchange_forward(gopf_dist_solver* s, const void* send_c128, int begin, int count);
int gopf_dist_exchange_inverse(gopf_dist_solver* s, const void* t_c128, int k1_begin, int k1_count);
int gopf_dist_exchange_join(gopf_dist_solver* s);
/* Pipelined peer-store exchange: forward axis 1 of planes [begin, begin+count) -> every rank's Y,
 * on the library's second (high-priority) stream after the compute stream's work so far, as a
 * persistent kernel of at most max_ctas CTAs (0: no limit): the NVLink-bound pass of one chunk
 * runs under the HBM-bound inverse_mid_planes / real_step_planes of the next.  exchange_join
 * makes the compute stream wait for it.  One step:
 *   for each plane chunk: inverse_mid_planes(X_local -> A); real_step_planes(A);
 *                         forward_mid_peer_planes(A)
 *   exchange_join; barrier; kspace_step_peer(S); barrier; advance()                        */
int gopf_dist_forward_mid_peer_planes(gopf_dist_solver* s, const void* w_c128, int begin, int count, int max_ctas);
int gopf_dist_advance(gopf_dist_solver* s);
int gopf_dist_solver_get_time(gopf_dist_solver* s, double* t);
int gopf_dist_solver_kernel_launches(gopf_dist_solver* s, int64_t* n, int reset);
int gopf_dist_solver_destroy(gopf_dist_solver* s);

/* page-locked host memory for Field.Data (NewField adopts a caller slice, pf/model.go:44-57) */
int gopf_host_alloc(int64_t bytes, void** out);
int gopf_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* GOPF_CUDA_H */
