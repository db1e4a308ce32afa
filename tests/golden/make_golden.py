"""Generates tests/golden/*.npz from the oracle restatement (oracle/), which is pinned to the
reference's own known-answer tests (tests/test_oracle_*.py).  The reference (Go + FFTW) cannot
run in this image, so these vectors are the oracle's outputs frozen at commit time: they guard
the oracle against drift (tests/test_golden_cpu.py) and give the CUDA path a fixture that does
not depend on the oracle's code at test time (tests/test_golden_gpu.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gopf_b200 import synthetic, workloads  # noqa: E402  (inputs only: seeded streams and model definitions)
from oracle import elasticity as oel  # noqa: E402
from oracle import pf as opf  # noqa: E402
from oracle import pfutil as opfutil  # noqa: E402
from oracle import sdd as osdd  # noqa: E402
from oracle import terms as oterms  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ch(dims, steps_list, stepper="euler"):
    n = opfutil.prod_int(dims)
    m = opf.NewModel()
    f = opf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = opf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT)
    s.SetStepper(stepper)
    out, done = {}, 0
    for k in steps_list:
        s.Propagate(k - done)
        done = k
        out[f"after_{k}"] = f.Data.copy()
    return out


def main():
    # k-tables and index maps (pfutil/fftWrap.go:42-95, indexPositionConversion.go:4-44)
    kt = {}
    for dims in ([8, 16], [9, 9], [8, 8, 8], [9, 9, 9]):
        ft = opfutil.NewFFTW(dims)
        tag = "x".join(map(str, dims))
        n = opfutil.prod_int(dims)
        kt[f"freq_{tag}"] = ft.freq_table()
        kt[f"conj_{tag}"] = np.array([ft.ConjugateNode(i) for i in range(n)], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "ktable.npz"), **kt)

    # transform of the ramp data[i] = i on 8x16 (pfutil/fftwWrap_test.go:11-57)
    ramp = np.arange(8 * 16, dtype=np.float64).astype(np.complex128)
    fwd = ramp.copy()
    opfutil.NewFFTW([8, 16]).FFT(fwd)
    np.savez_compressed(os.path.join(OUT, "fft_ramp_8x16.npz"), input=ramp, forward=fwd)

    # Cahn-Hilliard trajectories (cfg 1 structure, small)
    np.savez_compressed(os.path.join(OUT, "ch_2d_32x32_euler.npz"), **ch([32, 32], [1, 10, 100]))
    np.savez_compressed(os.path.join(OUT, "ch_3d_16_euler.npz"), **ch([16, 16, 16], [1, 10, 50]))
    np.savez_compressed(os.path.join(OUT, "ch_2d_32x32_rk4.npz"), **ch([32, 32], [5], "rk4"))

    # cfg 4: precipitate with elasticity + volume constraint
    for dims in ([32, 32], [16, 16, 16]):
        m, conc, phase, s, vol = workloads.build_precipitate(opf, oterms, oel, dims, expressions=False)
        s.Solve(2, 5)
        np.savez_compressed(os.path.join(OUT, f"precipitate_{'x'.join(map(str, dims))}.npz"), conc=conc.Data, phase=phase.Data,
                            multiplier=np.array([vol.Multiplier]))

    # cfg 5 without noise (deterministic), with Vandeven(5)
    m, f, s = workloads.build_pfc(opf, oterms, [32, 32], noise=None, filt_order=5)
    s.Solve(2, 5)
    np.savez_compressed(os.path.join(OUT, "pfc_32x32_vandeven5.npz"), density=f.Data)

    # Khachaturyan chain on random H^ (elasticity.Displacements + Strain contraction)
    dims = [8, 8, 8]
    f3 = oel.pad3(opfutil.NewFFTW(dims).freq_table())
    C = oel.CubicMaterial(110.0, 60.0, 30.0)
    mis = workloads.PRECIPITATE_MISFIT
    rng = np.random.default_rng(5)
    H = rng.standard_normal(512) + 1j * rng.standard_normal(512)
    eff = oel.EffectiveForce(C, mis)
    force = np.stack([eff.Get(c, f3, H) for c in range(3)], axis=1)
    disp = oel.Displacements(force, f3, C)
    A = C.ContractLast(mis)
    tot = np.zeros(512, dtype=np.complex128)
    for i in range(3):
        for j in range(i, 3):
            tot += (1.0 if i == j else 2.0) * A[i, j] * oel.Strain(disp, f3, i, j)
    np.savez_compressed(os.path.join(OUT, "khachaturyan_8.npz"), indicator_hat=H, contracted_strain_hat=tot, freq3=f3)
    # ChargeTransport, point sources, SDD, the epoch observers
    dims = [32, 32]
    m, f, term, s = workloads.build_charge(opf, oterms, dims, opfutil.NewFFTW(dims))
    s.Solve(2, 3)
    cur = term.Current(f, 32 * 32, True)
    np.savez_compressed(os.path.join(OUT, "charge_transport_32x32.npz"), density=f.Data, current=np.stack(cur))
    for dims in ([16, 32], [16, 16, 16]):
        m, f, s = workloads.build_sourced_diffusion(opf, oterms, dims)
        s.Solve(2, 5)
        np.savez_compressed(os.path.join(OUT, f"sources_{'x'.join(map(str, dims))}.npz"), conc=f.Data)
    m, phi, sdd, s = workloads.build_sdd_nucleation(opf, osdd.NewSDD, 64, expressions=False)
    s.Solve(1, 60)
    np.savez_compressed(os.path.join(OUT, "sdd_nucleation_64x64.npz"), phi=phi.Data, orientation=sdd.orientation,
                        monitor=np.array([sdd.Monitor.MaxForce, sdd.Monitor.ForcePowerSpectrum, sdd.Monitor.MaxTorque,
                                          sdd.Monitor.FieldNorm, sdd.Monitor.FieldNormChange]))
    # Uint8IO payload of the 2-D Cahn-Hilliard field after 10 steps and the pfcPhases energy observer
    traj = ch([32, 32], [10])
    m, f, s = workloads.build_pfc(opf, oterms, [32, 32], noise=None, filt_order=None)
    s.Solve(1, 5)
    np.savez_compressed(os.path.join(OUT, "observers.npz"), ch_uint8=opf.uint8_payload(traj["after_10"]),
                        pfc_energy=np.array([m.MixedTerms["IDEAL"].GetEnergy(m.Bricks, 1024),
                                             m.ImplicitTerms["EXCESS"].GetEnergy(m.Bricks, s.FT, [32, 32])]))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
