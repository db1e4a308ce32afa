// +build cuda

// Drop-in replacement for pfutil/fftWrap.go of davidkleiven/gopf: the same exported type,
// constructor, fields and methods, executing on a B200 through libgopfcuda.so (include/gopf_cuda.h)
// instead of FFTW.  Build the reference with `-tags cuda` and this file instead of fftWrap.go.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  The ctypes binding in
// gopf_b200/pfutil.py makes exactly these calls and is what the tests exercise.
package pfutil

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gopf_b200/lib -lgopfcuda
#include <stdlib.h>
#include "gopf_cuda.h"
*/
import "C"

import (
	"runtime"
	"unsafe"
)

// FFTWWrapper implements the pf.FourierTransform interface on the GPU
type FFTWWrapper struct {
	plan       *C.gopf_fft_plan
	Data       []complex128 // kept for API compatibility (pfutil/fftWrap.go:11); unused by the device path
	Dimensions []int
}

func check(status C.int) {
	if status != 0 {
		panic("gopfcuda: " + C.GoString(C.gopf_last_error()))
	}
}

func cInts(n []int) []C.int {
	out := make([]C.int, len(n))
	for i, v := range n {
		out[i] = C.int(v)
	}
	return out
}

// NewFFTW returns a new FFTWWrapper (pfutil/fftWrap.go:16-23)
func NewFFTW(n []int) *FFTWWrapper {
	fw := &FFTWWrapper{Dimensions: n}
	dims := cInts(n)
	check(C.gopf_fft_plan_create(C.int(len(n)), &dims[0], -1, &fw.plan))
	runtime.SetFinalizer(fw, func(f *FFTWWrapper) { C.gopf_fft_plan_destroy(f.plan) })
	return fw
}

// FFT performs forward fourier transform in place (pfutil/fftWrap.go:26-31)
func (fw *FFTWWrapper) FFT(data []complex128) []complex128 {
	check(C.gopf_fft_exec(fw.plan, (*C.double)(unsafe.Pointer(&data[0])), -1))
	return data
}

// IFFT performs the unnormalised inverse fourier transform in place (pfutil/fftWrap.go:34-39)
func (fw *FFTWWrapper) IFFT(data []complex128) []complex128 {
	check(C.gopf_fft_exec(fw.plan, (*C.double)(unsafe.Pointer(&data[0])), 1))
	return data
}

// Freq returns the frequency corresponding to site i (pfutil/fftWrap.go:57-74)
func (fw *FFTWWrapper) Freq(i int) []float64 {
	res := make([]float64, len(fw.Dimensions))
	dims := cInts(fw.Dimensions)
	check(C.gopf_freq(C.int(len(dims)), &dims[0], C.int64_t(i), (*C.double)(unsafe.Pointer(&res[0]))))
	return res
}

// ConjugateNode returns the node of the negative frequency (pfutil/fftWrap.go:78-95)
func (fw *FFTWWrapper) ConjugateNode(i int) int {
	var out C.int64_t
	dims := cInts(fw.Dimensions)
	check(C.gopf_conjugate_node(C.int(len(dims)), &dims[0], C.int64_t(i), &out))
	return int(out)
}

// GradientCalculate is pf.GradientCalculator{FT: fw, Comp: comp, KeepNyquist: keepNyquist}.Calculate(indata, data)
// (pf/gradientCalculator.go:19-31) on the device: data = IFFT(i 2 pi f_comp FFT(indata)) / N.
func (fw *FFTWWrapper) GradientCalculate(indata []complex128, data []complex128, comp int, keepNyquist bool) {
	keep := C.int(0)
	if keepNyquist {
		keep = 1
	}
	check(C.gopf_gradient_calculate(fw.plan, (*C.double)(unsafe.Pointer(&indata[0])), (*C.double)(unsafe.Pointer(&data[0])),
		C.int(comp), keep))
}

// DivGradConstruct is what pf.DivGrad{Field, F}.Construct's closure leaves in `out` once the derived fields of
// PrepareModel hold F * grad field (pf/gradientCalculator.go:72-108); funcValues[i] = F(i, bricks).
func (fw *FFTWWrapper) DivGradConstruct(field []complex128, funcValues []complex128, out []complex128) {
	check(C.gopf_div_grad_construct(fw.plan, (*C.double)(unsafe.Pointer(&field[0])), (*C.double)(unsafe.Pointer(&funcValues[0])),
		(*C.double)(unsafe.Pointer(&out[0]))))
}

// WeightedLaplacianConstruct is pf.WeightedLaplacian.Construct's closure (pf/gradientCalculator.go:131-172); both
// inputs are spectra.
func (fw *FFTWWrapper) WeightedLaplacianConstruct(fieldHat []complex128, prefactorHat []complex128, out []complex128) {
	check(C.gopf_weighted_laplacian_construct(fw.plan, (*C.double)(unsafe.Pointer(&fieldHat[0])),
		(*C.double)(unsafe.Pointer(&prefactorHat[0])), (*C.double)(unsafe.Pointer(&out[0]))))
}
