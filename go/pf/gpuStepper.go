// +build cuda

// GPU time stepper for davidkleiven/gopf: implements pf.TimeStepper (pf/solver.go:15-19) by
// mirroring the host Model into libgopfcuda.so (include/gopf_cuda.h) and stepping on the device.
// Usage, with the reference's API otherwise unchanged:
//
//	solver := pf.NewSolver(&model, domainSize, dt)
//	solver.Stepper = pf.NewGPUStepper(&model, domainSize, dt, "euler")   // Solver.Stepper is exported
//	solver.Solve(nepochs, nsteps)
//
// Step() uploads Field.Data, advances one step and downloads (API-exact, PCIe-bound);
// Propagate(n) does the same around n device-resident steps and is what a GPU-aware
// Solver.Propagate should call once per epoch (host data only needs to be current when callbacks
// and monitors fire, pf/solver.go:110-117).
//
// Registered Go closures cannot run on the device: functions must be given as expressions through
// RegisterFunctionExpr, user terms must be catalog types (type switch below).  Anything else panics
// at construction, not in the middle of a run.
//
// cgo note: a file that uses //export may only DECLARE C functions in its preamble; the two static
// helpers above are definitions, so in a real checkout they live in a second file of the package
// (gpuStepper_helpers.go with the same preamble minus the extern).  Kept here so the binding reads
// top to bottom.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no Go toolchain in the image); gopf_b200/pf.py is the
// binding the tests exercise, call for call.
package pf

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gopf_b200/lib -lgopfcuda
#include <stdlib.h>
#include <stdint.h>
#include "gopf_cuda.h"

// defined below with //export: calls the Go TimeDepSource registered under the index in `user`
extern double gopfSourceTrampoline(double t, void* user);
static gopf_time_fn gopf_source_trampoline_ptr(void) { return (gopf_time_fn)gopfSourceTrampoline; }
static void* gopf_index_as_ptr(uintptr_t i) { return (void*)i; }
*/
import "C"

import (
	"fmt"
	"os"
	"sync"
	"unsafe"
)

// GPUStepper is a TimeStepper backed by gopf_solver
type GPUStepper struct {
	model   *C.gopf_model
	solver  *C.gopf_solver
	Dt      float64
	sources []uintptr // handles of this stepper's TimeDepSource closures in gpuSourceFuncs
}

// TimeDepSource closures of the models' point sources (pf/sourceTerm.go:11); the C library calls
// them back on the host once per right-hand-side evaluation through gopfSourceTrampoline with the
// handle it was given at registration.  Handles belong to the stepper that registered them and are
// released by its Close; the table is guarded because steppers may live on different goroutines.
var (
	gpuSourceMu    sync.Mutex
	gpuSourceFuncs = map[uintptr]TimeDepSource{}
	gpuSourceNext  uintptr
)

func gpuRegisterSource(f TimeDepSource) uintptr {
	gpuSourceMu.Lock()
	defer gpuSourceMu.Unlock()
	gpuSourceNext++
	gpuSourceFuncs[gpuSourceNext] = f
	return gpuSourceNext
}

// GPUNoiseFuncs names the RegisterFunction entries that are WhiteNoise.Generate closures
// (pf/noise.go:20-23): a Go closure cannot run on the device, so NewGPUStepper registers the device
// generator of the same distribution for them (gopf_model_register_white_noise).  Set it before
// NewGPUStepper, e.g. pf.GPUNoiseFuncs = map[string]*pf.WhiteNoise{"NOISE": &noise}; GPUNoiseSeed
// seeds the counter-based stream (the reference's math/rand stream is unpinned).
var (
	GPUNoiseFuncs map[string]*WhiteNoise
	GPUNoiseSeed  uint64
)

// GPUKSpaceNoise asks NewGPUStepper to draw the spectrum of plain explicit WhiteNoise terms at the
// k-point (gopf_model_set_kspace_noise) instead of transforming a real-space noise field every step.
// The library must have been built with -DGOPF_KNOISE (C.gopf_has_kspace_noise() == 1).
var GPUKSpaceNoise = false

//export gopfSourceTrampoline
func gopfSourceTrampoline(t C.double, user unsafe.Pointer) C.double {
	gpuSourceMu.Lock()
	f := gpuSourceFuncs[uintptr(user)]
	gpuSourceMu.Unlock()
	return C.double(f(float64(t)))
}

func gpuCheck(status C.int) {
	if status != 0 {
		panic("gopfcuda: " + C.GoString(C.gopf_last_error()))
	}
}

func cstr(s string) *C.char { return C.CString(s) }

// FunctionExprs maps RegisterFunction names to device expressions, e.g.
// "CHEMICALPOT": "-((0.1*conc*(1-H(phase)) - 0.1*(1-conc)*H(phase))*1.0)"
type FunctionExprs map[string]string

// NewGPUStepper mirrors m (fields, scalars, equations, catalog terms) into the device library.
func NewGPUStepper(m *Model, domainSize []int, dt float64, scheme string, exprs FunctionExprs) *GPUStepper {
	st := &GPUStepper{Dt: dt}
	gpuCheck(C.gopf_model_create(&st.model))
	// cgo pointer rule: the library keeps the Field.Data pointer for the model's lifetime, so the
	// backing arrays must not be Go-heap memory.  Allocate them with gopf_host_alloc (page-locked,
	// C-owned) and hand them to pf.NewField(name, N, data), which adopts a caller slice
	// (pf/model.go:44-57); INTEGRATION.md section 2 shows the helper.
	for _, f := range m.Fields {
		name := cstr(f.Name)
		gpuCheck(C.gopf_model_add_field(st.model, name, C.int64_t(len(f.Data)), (*C.double)(unsafe.Pointer(&f.Data[0]))))
		C.free(unsafe.Pointer(name))
	}
	for name, b := range m.Bricks {
		if s, ok := b.(*Scalar); ok {
			cn := cstr(name)
			gpuCheck(C.gopf_model_add_scalar(st.model, cn, C.double(real(s.Value)), C.double(imag(s.Value))))
			C.free(unsafe.Pointer(cn))
		}
	}
	for name, expr := range exprs {
		cn, ce := cstr(name), cstr(expr)
		gpuCheck(C.gopf_model_register_function(st.model, cn, ce))
		C.free(unsafe.Pointer(cn))
		C.free(unsafe.Pointer(ce))
	}
	for name, noise := range GPUNoiseFuncs { // RegisterFunction(name, noise.Generate), pf/noise.go:20-23
		cn := cstr(name)
		gpuCheck(C.gopf_model_register_white_noise(st.model, cn, C.double(noise.Strength), C.uint64_t(GPUNoiseSeed)))
		C.free(unsafe.Pointer(cn))
	}
	register := func(name string, t interface{}) {
		cn := cstr(name)
		defer C.free(unsafe.Pointer(cn))
		switch v := t.(type) {
		case *SpectralViscosity:
			gpuCheck(C.gopf_model_register_spectral_viscosity(st.model, cn, C.double(v.Eps), C.double(v.DissipationThreshold), C.int(v.Power)))
		case *TensorialHessian:
			cf := cstr(v.Field)
			k := make([]C.double, len(v.K))
			for i, x := range v.K {
				k[i] = C.double(x)
			}
			gpuCheck(C.gopf_model_register_tensorial_hessian(st.model, cn, cf, &k[0], C.int(len(k))))
			C.free(unsafe.Pointer(cf))
		case *VolumeConservingLP:
			cf, ci := cstr(v.Field), cstr(v.Indicator)
			gpuCheck(C.gopf_model_register_volume_conserving_lp(st.model, cn, cf, ci, C.double(v.Dt)))
			C.free(unsafe.Pointer(cf))
			C.free(unsafe.Pointer(ci))
		case *SquaredGradient:
			cf := cstr(v.Field)
			gpuCheck(C.gopf_model_register_squared_gradient(st.model, cn, cf, C.double(v.Factor)))
			C.free(unsafe.Pointer(cf))
		case *HomogeneousModulusLinElast:
			// MatProp.Data is the 81-element rank-4 tensor (elasticity/rank4.go:10-24), Misfit a 3x3 mat.Dense
			cf := cstr(v.FieldName)
			misfit := make([]C.double, 9)
			for i := 0; i < 3; i++ {
				for j := 0; j < 3; j++ {
					misfit[3*i+j] = C.double(v.Misfit.At(i, j))
				}
			}
			stiff := make([]C.double, 81)
			for i, x := range v.MatProp.Data {
				stiff[i] = C.double(x)
			}
			gpuCheck(C.gopf_model_register_homogeneous_modulus_lin_elast(st.model, cn, cf, &stiff[0], &misfit[0]))
			C.free(unsafe.Pointer(cf))
		case *PairCorrlationTerm:
			registerPairCorrelation(st.model, cn, v, 0)
		case *ExplicitPairCorrelationTerm:
			registerPairCorrelation(st.model, cn, &v.PairCorrlationTerm, 1)
		case *IdealMixtureTerm:
			cf := cstr(v.Field)
			lap := 0
			if v.Laplacian {
				lap = 1
			}
			gpuCheck(C.gopf_model_register_ideal_mixture(st.model, cn, cf, C.double(v.IdealMix.C3), C.double(v.IdealMix.C4), C.double(v.Prefactor), C.int(lap), 1))
			C.free(unsafe.Pointer(cf))
		case *ConservativeNoise:
			gpuCheck(C.gopf_model_register_conservative_noise(st.model, cn, C.double(v.Strength), C.int(v.Dim), C.uint32_t(v.UniquePrefix), 0))
		case *ChargeTransport:
			// Conductivity(i) is a Go closure (pf/chargeTransport.go:29-37): tabulate it once,
			// component-major [n_voigt][N]; the library copies the table
			n := len(m.Fields[0].Data)
			nv := len(v.Conductivity(0))
			tab := make([]C.double, nv*n)
			for i := 0; i < n; i++ {
				for c, x := range v.Conductivity(i) {
					tab[c*n+i] = C.double(x)
				}
			}
			ext := make([]C.double, len(v.ExternalField))
			for i, x := range v.ExternalField {
				ext[i] = C.double(x)
			}
			cf := cstr(v.Field)
			gpuCheck(C.gopf_model_register_charge_transport(st.model, cn, cf, &tab[0], C.int(nv), C.int64_t(n), &ext[0], C.int(len(ext))))
			C.free(unsafe.Pointer(cf))
		default:
			panic(fmt.Sprintf("gopfcuda: term %s (%T) has no device implementation", name, t))
		}
	}
	for name, t := range m.ImplicitTerms {
		register(name, t)
	}
	for name, t := range m.ExplicitTerms {
		register(name, t)
	}
	for name, t := range m.MixedTerms {
		register(name, t)
	}
	for _, eq := range m.Equations {
		ce := cstr(eq)
		gpuCheck(C.gopf_model_add_equation(st.model, ce))
		C.free(unsafe.Pointer(ce))
	}
	for eqNo, srcs := range m.AllSources { // pf/model.go:151-154, 291-294
		for i := range srcs {
			pos := make([]C.double, len(srcs[i].Pos))
			for k, x := range srcs[i].Pos {
				pos[k] = C.double(x)
			}
			h := gpuRegisterSource(srcs[i].f)
			st.sources = append(st.sources, h)
			gpuCheck(C.gopf_model_add_source(st.model, C.int(eqNo), &pos[0], C.int(len(pos)),
				C.gopf_source_trampoline_ptr(), C.gopf_index_as_ptr(C.uintptr_t(h))))
		}
	}
	dims := make([]C.int, len(domainSize))
	for i, v := range domainSize {
		dims[i] = C.int(v)
	}
	if GPUKSpaceNoise {
		gpuCheck(C.gopf_model_set_kspace_noise(st.model, 1))
	}
	gpuCheck(C.gopf_solver_create(st.model, C.int(len(dims)), &dims[0], C.double(dt), -1, &st.solver))
	cs := cstr(scheme)
	gpuCheck(C.gopf_solver_set_stepper(st.solver, cs))
	C.free(unsafe.Pointer(cs))
	return st
}

func registerPairCorrelation(m *C.gopf_model, cn *C.char, v *PairCorrlationTerm, explicit C.int) {
	n := len(v.PairCorrFunc.Peaks)
	dens, loc, wid := make([]C.double, n), make([]C.double, n), make([]C.double, n)
	planes := make([]C.int, n)
	for i, p := range v.PairCorrFunc.Peaks {
		dens[i], loc[i], wid[i], planes[i] = C.double(p.PlaneDensity), C.double(p.Location), C.double(p.Width), C.int(p.NumPlanes)
	}
	cf := cstr(v.Field)
	lap := 0
	if v.Laplacian {
		lap = 1
	}
	gpuCheck(C.gopf_model_register_pair_correlation(m, cn, explicit, cf, C.double(v.Prefactor), C.int(lap),
		C.double(v.PairCorrFunc.EffTemp), C.int(n), &dens[0], &loc[0], &wid[0], &planes[0]))
	C.free(unsafe.Pointer(cf))
}

// Step performs one step on the host Field.Data arrays (pf.TimeStepper)
func (st *GPUStepper) Step(m *Model) { gpuCheck(C.gopf_solver_propagate(st.solver, 1)) }

// Propagate performs nsteps device-resident steps between one upload and one download
func (st *GPUStepper) Propagate(nsteps int, m *Model) { gpuCheck(C.gopf_solver_propagate(st.solver, C.int(nsteps))) }

// SetFilter sets a tabulated modal filter (pf.TimeStepper); only *Vandeven carries a table
func (st *GPUStepper) SetFilter(filter ModalFilter) {
	if v, ok := filter.(*Vandeven); ok {
		gpuCheck(C.gopf_solver_set_filter(st.solver, (*C.double)(unsafe.Pointer(&v.Data[0])), C.int(len(v.Data))))
		return
	}
	panic("gopfcuda: only tabulated filters (pf.Vandeven) run on the device")
}

// SetNewtonKrylov forwards the settings of ImplicitEuler.NonlinSolver (pf/implicitEuler.go:221-229)
// to the device solver; scheme "implicit_euler" replaces &pf.ImplicitEuler{Dt, FT}.  InnerMethod has
// no counterpart: the device runs its own restarted GMRES (restart 30, inner tolerance 1e-4).
func (st *GPUStepper) SetNewtonKrylov(maxiter int, stepSize, tol float64, stencil int) {
	gpuCheck(C.gopf_solver_set_newton_krylov(st.solver, C.int(maxiter), C.double(stepSize), C.double(tol), C.int(stencil), 30, 1e-4, 4))
}

// Converged reports res.Converged of the last implicit step (the reference logs a warning when
// it is false, pf/implicitEuler.go:202-204)
func (st *GPUStepper) Converged() bool {
	var c C.int
	var n C.int64_t
	gpuCheck(C.gopf_solver_newton_krylov_status(st.solver, &c, &n))
	return c != 0
}

// SaveReal writes the real part of field i of the device-resident state as big-endian float64,
// byte for byte what Field.SaveReal writes (pf/model.go:35-41, pf/fileIO.go:85-95); the byte swap
// runs on the device and only 8 bytes per cell cross PCIe.
func (st *GPUStepper) SaveReal(i int, n int, fname string) {
	buf := make([]byte, 8*n)
	gpuCheck(C.gopf_solver_download_real(st.solver, C.int(i), (*C.double)(unsafe.Pointer(&buf[0])), 1))
	if err := os.WriteFile(fname, buf, 0644); err != nil {
		panic(err)
	}
}

// NewGPUSDD mirrors a configured pf.SDD (pf/sdd.go:86-129) onto the device: scheme "sdd", the
// orientation vector (already normalised by Init / SetInitialOrientation) and the exported settings.
// Call SyncSDD(&sdd) after changing a setting; Monitor / CurrentStep / orientation are read back by
// PullSDD(&sdd) (e.g. from a Solver callback).
func NewGPUSDD(m *Model, domainSize []int, sdd *SDD, exprs FunctionExprs) *GPUStepper {
	st := NewGPUStepper(m, domainSize, sdd.Dt, "sdd", exprs)
	if sdd.initialized {
		gpuCheck(C.gopf_solver_sdd_set_orientation(st.solver, (*C.double)(unsafe.Pointer(&sdd.orientation[0])), C.int64_t(len(sdd.orientation))))
	}
	st.SyncSDD(sdd)
	return st
}

func (st *GPUStepper) sddSet(key string, v float64) {
	ck := cstr(key)
	defer C.free(unsafe.Pointer(ck))
	gpuCheck(C.gopf_solver_sdd_set(st.solver, ck, C.double(v)))
}

func (st *GPUStepper) sddGet(key string) float64 {
	ck := cstr(key)
	defer C.free(unsafe.Pointer(ck))
	var v C.double
	gpuCheck(C.gopf_solver_sdd_get(st.solver, ck, &v))
	return float64(v)
}

// SyncSDD pushes the exported fields of the SDD struct (set_orientation overwrote InitDimerLength
// with the norm of the unit vector, so it is pushed again here)
func (st *GPUStepper) SyncSDD(sdd *SDD) {
	st.sddSet("Alpha", sdd.Alpha)
	st.sddSet("Dt", sdd.Dt)
	st.sddSet("TimeConstants.Orientation", sdd.TimeConstants.Orientation)
	st.sddSet("TimeConstants.DimerLength", sdd.TimeConstants.DimerLength)
	st.sddSet("MinDimerLength", sdd.MinDimerLength)
	st.sddSet("InitDimerLength", sdd.InitDimerLength)
	st.sddSet("CurrentStep", float64(sdd.CurrentStep))
}

// PullSDD refreshes Monitor, CurrentStep and the orientation vector from the device
func (st *GPUStepper) PullSDD(sdd *SDD) {
	sdd.CurrentStep = int(st.sddGet("CurrentStep"))
	sdd.Monitor.MaxForce = st.sddGet("Monitor.MaxForce")
	sdd.Monitor.ForcePowerSpectrum = st.sddGet("Monitor.ForcePowerSpectrum")
	sdd.Monitor.MaxTorque = st.sddGet("Monitor.MaxTorque")
	sdd.Monitor.FieldNorm = st.sddGet("Monitor.FieldNorm")
	sdd.Monitor.FieldNormChange = st.sddGet("Monitor.FieldNormChange")
	if sdd.initialized {
		gpuCheck(C.gopf_solver_sdd_get_orientation(st.solver, (*C.double)(unsafe.Pointer(&sdd.orientation[0]))))
	}
}

// SaveUint8 is Uint8IO.SaveFields for field i (pf/fileIO.go:29-44) from the device-resident state:
// min / max reduction and the 0..255 scaling run on the device, n bytes cross PCIe.
func (st *GPUStepper) SaveUint8(i int, n int, fname string) {
	buf := make([]byte, n)
	gpuCheck(C.gopf_solver_download_uint8(st.solver, C.int(i), (*C.uint8_t)(unsafe.Pointer(&buf[0])), nil, nil))
	if err := os.WriteFile(fname, buf, 0644); err != nil {
		panic(err)
	}
}

// TermEnergy is IdealMixtureTerm.GetEnergy / PairCorrlationTerm.GetEnergy
// (pf/pairCorrelationTerm.go:58-84, 185-193) of the term registered as name, on the device state.
func (st *GPUStepper) TermEnergy(name string) float64 {
	cn := cstr(name)
	defer C.free(unsafe.Pointer(cn))
	var e C.double
	gpuCheck(C.gopf_solver_term_energy(st.solver, cn, &e))
	return float64(e)
}

// ChargeCurrent is ChargeTransport.Current (pf/chargeTransport.go:121-146) for the term registered
// as name, evaluated on the device-resident state: res[d][i] = -real(current_d[i]).
func (st *GPUStepper) ChargeCurrent(name string, dim int, n int) [][]float64 {
	flat := make([]float64, dim*n)
	cn := cstr(name)
	defer C.free(unsafe.Pointer(cn))
	gpuCheck(C.gopf_solver_charge_current(st.solver, cn, (*C.double)(unsafe.Pointer(&flat[0]))))
	res := make([][]float64, dim)
	for d := 0; d < dim; d++ {
		res[d] = flat[d*n : (d+1)*n]
	}
	return res
}

// SetJit switches the run-time specialisation on or off: registered functions and the k-space
// update are then compiled for this model by NVRTC at the next step instead of being interpreted.
// Default: on (GOPF_JIT=0 in the environment switches it off).
func (st *GPUStepper) SetJit(on bool) {
	v := C.int(0)
	if on {
		v = 1
	}
	gpuCheck(C.gopf_solver_set_jit(st.solver, v))
}

// JitKernels returns how many kernels currently run as compiled images (0: interpreter kernels).
func (st *GPUStepper) JitKernels() int {
	var n C.int
	gpuCheck(C.gopf_solver_jit_kernels(st.solver, &n))
	return int(n)
}

// GetTime returns the current time (pf.TimeStepper)
func (st *GPUStepper) GetTime() float64 {
	var t C.double
	gpuCheck(C.gopf_solver_get_time(st.solver, &t))
	return float64(t)
}

// Close releases the device resources
func (st *GPUStepper) Close() {
	C.gopf_solver_destroy(st.solver)
	C.gopf_model_destroy(st.model)
	gpuSourceMu.Lock()
	for _, h := range st.sources {
		delete(gpuSourceFuncs, h)
	}
	gpuSourceMu.Unlock()
	st.sources = nil
}
