"""CPU oracle for the gopf spectral time-stepping hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy / scipy.fft (pocketfft, fp64)
restatement of the reference algorithm (davidkleiven/gopf, Go + FFTW).  It is the
checker the CUDA path is compared against; it is never the thing shipped or
measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``gopf_b200/`` imports it, and the product path raises when ``libgopfcuda.so`` is
missing instead of falling back to this code.

Parity status: PINNED against every known-answer test the reference holds for
this path -- see ``tests/test_oracle_*.py``, each of
which cites the Go test it restates.  The reference stores no golden arrays and
no multi-step trajectory, and the Go toolchain / libfftw3 are absent from this
image, so multi-step Cahn-Hilliard parity is anchored on this restatement only.
Third-party arithmetic restated from its published definition:
  * github.com/barnex/fftw (v0.0.0-20181125072904-b800f77a10de) -> libfftw3
    c2c DFT, sign -1 forward / +1 inverse, unnormalised      -> scipy.fft
  * github.com/davidkleiven/gosfft v1.0.2 (same DFT)          -> scipy.fft
  * gonum v0.9.0 floats.Norm (2-norm), mat.Dense.Solve (LU)   -> numpy
  * Go math/cmplx.Pow (polar form)                            -> ``go_cpow``

All ``file:line`` citations are relative to /root/reference.
"""
