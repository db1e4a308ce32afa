"""Oracle restatement of the reference's term catalog on the hot path
(SURVEY.md 8a rows a11, a13, a16-a19).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations: /root/reference.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np

from . import pfutil
from .pf import DerivedField, LaplacianN, as_frequency


# ==========================================================================
# pf/squareGradientTerm.go
# ==========================================================================
class SquaredGradient:
    """pf/squareGradientTerm.go:14-68: Factor * sum_d FFT((IFFT(i 2pi f_d u^)/N)^2);
    only the +0.5 Nyquist is zeroed (:48).  The reference uses gosfft here; the
    DFT is the same, so the oracle uses its own FFTWWrapper."""

    def __init__(self, field: str, domain_size, workers: int = 1):
        if len(domain_size) not in (2, 3):
            raise ValueError("squaregradient: Domain size has to be of length 2 or 3")
        self.Field = field
        self.Factor = 1.0
        self.FT = pfutil.NewFFTW(domain_size, workers)

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            n = field.shape[0]
            field[:] = 0.0
            f = as_frequency(freq).table(n)
            dim = f.shape[1]
            idx = np.arange(n)
            for d in range(dim):
                fd = f[:, d].copy()
                fd[np.abs(fd - 0.5) < 1e-10] = 0.0
                work = np.asarray(bricks[self.Field].Get(idx)) * (1j * (2.0 * math.pi * fd))
                work = np.ascontiguousarray(work, dtype=np.complex128)
                self.FT.IFFT(work)
                work = pfutil.go_cpow(work / complex(float(n), 0.0), 2.0)
                work = np.ascontiguousarray(work)
                self.FT.FFT(work)
                field += work * complex(self.Factor, 0.0)

        return fn

    def OnStepFinished(self, t, bricks):
        pass


def NewSquareGradient(field: str, domain_size, workers: int = 1) -> SquaredGradient:
    return SquaredGradient(field, domain_size, workers)


# ==========================================================================
# pf/vandeven.go
# ==========================================================================
class Vandeven:
    """pf/vandeven.go:8-40: 1000-point trapezoid table + linear interpolation."""

    def __init__(self, order: int):
        data = np.zeros(1000, dtype=np.float64)
        data[0] = 1.0
        prefactor = math.gamma(float(2 * order)) / (math.gamma(float(order)) * math.gamma(float(order)))
        dx = 1.0 / 999.0
        for i in range(1, 1000):
            x = float(i) * dx
            i2 = math.pow(x * (1 - x), float(order - 1))
            x1 = x - dx
            i1 = math.pow(x1 * (1.0 - x1), float(order - 1))
            data[i] = data[i - 1] - prefactor * 0.5 * (i1 + i2) * dx
        self.Data = data

    def Eval(self, x: float) -> float:
        n = float(len(self.Data) - 1)
        idx = int(x * n)
        if idx >= len(self.Data) - 1:
            return float(self.Data[-1])
        dx = 1.0 / n
        dy = self.Data[idx + 1] - self.Data[idx]
        x0 = float(idx) * dx
        return float(self.Data[idx] + (x - x0) * dy / dx)

    def eval_array(self, x: np.ndarray) -> np.ndarray:
        n = float(len(self.Data) - 1)
        idx = (x * n).astype(np.int64)  # Go int() truncates toward zero; x >= 0
        last = idx >= len(self.Data) - 1
        idc = np.where(last, 0, idx)
        dx = 1.0 / n
        dy = self.Data[idc + 1] - self.Data[idc]
        x0 = idc.astype(np.float64) * dx
        val = self.Data[idc] + (x - x0) * dy / dx
        return np.where(last, self.Data[-1], val)


def NewVandeven(order: int) -> Vandeven:
    return Vandeven(order)


# ==========================================================================
# pf/spectralViscosity.go
# ==========================================================================
def interpolant(f: float, peak_position: float) -> float:
    """pf/spectralViscosity.go:31-40 (2x^2 - 3x^3 as written; pinned by
    spectralViscosity_test.go:10-35)."""
    frac = 1.0 / 3.0
    if f < frac * peak_position:
        return 0.0
    if f > peak_position:
        return 1.0
    x = 1.5 * (f - frac * peak_position) / peak_position
    return 2.0 * x * x - 3.0 * x * x * x


def interpolant_array(f: np.ndarray, peak_position: float) -> np.ndarray:
    frac = 1.0 / 3.0
    x = 1.5 * (f - frac * peak_position) / peak_position
    mid = 2.0 * x * x - 3.0 * x * x * x
    return np.where(f < frac * peak_position, 0.0, np.where(f > peak_position, 1.0, mid))


class SpectralViscosity:
    """pf/spectralViscosity.go:23-56: implicit term -Eps*Q(|f|)*|f|^Power."""

    def __init__(self, Eps: float, DissipationThreshold: float, Power: int):
        self.Eps = Eps
        self.DissipationThreshold = DissipationThreshold
        self.Power = Power

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            f = as_frequency(freq).table(field.shape[0])
            f_rad = np.sqrt(np.sum(f * f, axis=1))
            value = interpolant_array(f_rad, self.DissipationThreshold)
            field[:] = -self.Eps * value * np.power(f_rad, float(self.Power))

        return fn

    def OnStepFinished(self, t, bricks):
        pass


class TensorialHessian:
    """pf/tensorialHessian.go:17-74: implicit term sum_ij K_ij d_i d_j, i.e. the multiplier
    -4 pi^2 sum_j f_j^2 K_jj - 8 pi^2 sum_{j<k} f_j f_k K_jk."""

    def __init__(self, K, Field: str = ""):
        self.Field = Field
        self.K = [float(v) for v in K]

    def GetCoeff(self, i: int, j: int) -> float:
        d = 3 if len(self.K) == 9 else 2  # :65-71
        return self.K[i * d + j]

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            f = as_frequency(freq).table(field.shape[0])
            dim = f.shape[1]
            acc = np.zeros(field.shape[0], dtype=np.float64)
            for j in range(dim):
                acc += -4.0 * math.pi * math.pi * f[:, j] * f[:, j] * self.GetCoeff(j, j)
            for j in range(dim):
                for k in range(j + 1, dim):
                    acc += -8.0 * math.pi * math.pi * f[:, j] * f[:, k] * self.GetCoeff(j, k)
            field[:] = acc

        return fn

    def OnStepFinished(self, t, bricks):
        pass


# ==========================================================================
# pf/noise.go
# ==========================================================================
class WhiteNoise:
    """pf/noise.go:11-23.  The Go math/rand stream cannot be reproduced
    (unseeded global source), so the normal draws come from an injectable
    ``normal(n) -> array`` callable; parity tests share one pre-generated array."""

    def __init__(self, Strength: float, normal: Optional[Callable[[int], np.ndarray]] = None):
        self.Strength = Strength
        self.normal = normal or (lambda n: np.random.default_rng().standard_normal(n))

    def Generate(self, i, bricks):
        std = math.sqrt(2.0 * self.Strength)
        n = i.shape[0] if isinstance(i, np.ndarray) else 1
        return (self.normal(n) * std).astype(np.complex128)


class ConservativeNoise:
    """pf/noise.go:25-100: sum_c 2i sin(pi f_c) xi^_c, modes with ||f_c|-1/2| <= 1e-6
    skipped (:70)."""

    def __init__(self, strength: float, dim: int, unique_prefix: int = 0,
                 normal: Optional[Callable[[int], np.ndarray]] = None):
        self.UniquePrefix = unique_prefix
        self.Strength = strength
        self.Dim = dim
        self.normal = normal or (lambda n: np.random.default_rng().standard_normal(n))

    def GetCurrentName(self, comp: int) -> str:
        return f"{self.UniquePrefix}_current_{comp}"

    def CurrentFieldsAreRegistered(self, bricks) -> bool:
        return all(self.GetCurrentName(c) in bricks for c in range(self.Dim))

    def Construct(self, bricks):
        if not self.CurrentFieldsAreRegistered(bricks):
            raise RuntimeError("ConservativeCurrent: Current fields are not register.")

        def fn(freq, t, field: np.ndarray):
            n = field.shape[0]
            field[:] = 0.0
            f = as_frequency(freq).table(n)
            idx = np.arange(n)
            for comp in range(self.Dim):
                fc = f[:, comp]
                keep = np.abs(np.abs(fc) - 0.5) > 1e-6
                contrib = (1j * 2.0 * np.sin(math.pi * fc)) * np.asarray(bricks[self.GetCurrentName(comp)].Get(idx))
                field += np.where(keep, contrib, 0.0)

        return fn

    def OnStepFinished(self, t, bricks):
        pass

    def RequiredDerivedFields(self, num_nodes: int) -> List[DerivedField]:
        out = []
        for i in range(self.Dim):
            def calc(data, self=self):
                std = math.sqrt(2.0 * self.Strength)
                data[:] = self.normal(data.shape[0]) * std

            out.append(DerivedField(np.zeros(num_nodes, dtype=np.complex128), self.GetCurrentName(i), calc))
        return out


# ==========================================================================
# pf/volumeConserving.go
# ==========================================================================
class VolumeConservingLP:
    """pf/volumeConserving.go:3-61."""

    def __init__(self, field_name: str, indicator: str, dt: float, num_nodes: int):
        self.Multiplier = 0.0
        self.Indicator = indicator
        self.Field = field_name
        self.CurrentIntegral = 0.0
        self.Dt = dt
        self.NumNodes = num_nodes
        self.IsFirstUpdate = True

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            if self.Indicator not in bricks:
                raise RuntimeError("VolumeConservingLP: Indicator is not a derived field")
            idx = np.arange(field.shape[0])
            field[:] = np.asarray(bricks[self.Indicator].Get(idx)) * complex(self.Multiplier, 0.0)

        return fn

    def OnStepFinished(self, t, bricks):
        idx = np.arange(self.NumNodes)
        # sequential left-to-right sum in the reference; pairwise here (diff ~1e-16 rel)
        field_integral = float(np.sum(np.asarray(bricks[self.Field].Get(idx)).real))
        indicator_integral = float(np.real(bricks[self.Indicator].Get(0)))
        if self.IsFirstUpdate:
            self.CurrentIntegral = field_integral
            self.IsFirstUpdate = False
        else:
            delta = field_integral - self.CurrentIntegral
            self.CurrentIntegral = field_integral
            self.Multiplier = self.Multiplier - delta / (self.Dt * indicator_integral)


def NewVolumeConservingLP(field_name, indicator, dt, num_nodes) -> VolumeConservingLP:
    return VolumeConservingLP(field_name, indicator, dt, num_nodes)


# ==========================================================================
# pfc/pairCorrelation.go, pfc/ideal.go, pf/pairCorrelationTerm.go
# ==========================================================================
class Peak:
    def __init__(self, PlaneDensity: float, Location: float, Width: float, NumPlanes: int):
        self.PlaneDensity = PlaneDensity
        self.Location = Location
        self.Width = Width
        self.NumPlanes = NumPlanes


class ReciprocalSpacePairCorrelation:
    """pfc/pairCorrelation.go:18-37."""

    def __init__(self, EffTemp: float, Peaks: List[Peak]):
        self.EffTemp = EffTemp
        self.Peaks = Peaks

    def eval_array(self, k: np.ndarray) -> np.ndarray:
        result = np.zeros_like(k)
        for p in self.Peaks:
            pref = np.exp(-self.EffTemp * self.EffTemp * k * k / (2.0 * p.PlaneDensity * float(p.NumPlanes)))
            value = pref * np.exp(-0.5 * np.power((k - p.Location) / p.Width, 2))
            result = np.where(value > result, value, result)
        return result

    def Eval(self, k: float) -> float:
        return float(self.eval_array(np.array([k], dtype=np.float64))[0])


def SquareLattice2D(width: float, a: float) -> List[Peak]:
    a2 = a / math.sqrt(2.0)
    return [Peak(1.0, 2.0 * math.pi / a, width, 4), Peak(1.0 / math.sqrt(2.0), 2.0 * math.pi / a2, width, 4)]


def TriangularLattice2D(width: float, a: float) -> List[Peak]:
    return [Peak(2.0, 2.0 * math.pi / a, width, 3)]


class IdealMix:
    """pfc/ideal.go:17-50."""

    def __init__(self, C3: float, C4: float):
        self.C3 = C3
        self.C4 = C4

    def QuadraticPrefactor(self):
        return 0.5

    def ThirdOrderPrefactor(self):
        return -self.C3 / 6.0

    def FourthOrderPrefactor(self):
        return self.C4 / 12.0

    def Eval(self, n):
        return self.QuadraticPrefactor() * n * n + self.ThirdOrderPrefactor() * n * n * n + self.FourthOrderPrefactor() * n * n * n * n

    def Deriv(self, n):
        return 2.0 * self.QuadraticPrefactor() * n + 3.0 * self.ThirdOrderPrefactor() * n * n + 4.0 * self.FourthOrderPrefactor() * n * n * n


class PairCorrlationTerm:
    """pf/pairCorrelationTerm.go:22-84 (implicit): -Prefactor*C2(2pi|f|) [* (-k^2)]."""

    def __init__(self, PairCorrFunc, Field: str, Prefactor: float, Laplacian: bool):
        self.PairCorrFunc = PairCorrFunc
        self.Field = Field
        self.Prefactor = Prefactor
        self.Laplacian = Laplacian

    def _multiplier(self, freq, n):
        f = as_frequency(freq).table(n)
        f_rad = np.sqrt(np.sum(f * f, axis=1))
        return self.Prefactor * self.PairCorrFunc.eval_array(2.0 * math.pi * f_rad)

    def Construct(self, bricks):
        def fn(freq, t, out: np.ndarray):
            out[:] = -self._multiplier(freq, out.shape[0])
            if self.Laplacian:
                LaplacianN(1).Eval(freq, out)

        return fn

    def OnStepFinished(self, t, bricks):
        pass

    def GetEnergy(self, bricks, ft, domain_size) -> float:
        n = pfutil.prod_int(domain_size)
        b = np.asarray(bricks[self.Field].Get(np.arange(n)))
        field = np.array(b, dtype=np.complex128)
        ft.FFT(field)
        field *= self._multiplier(ft.Freq, n)
        ft.IFFT(field)
        pfutil.div_real_scalar(field, float(n))
        return -0.5 * float(np.sum((field * b).real))


class ExplicitPairCorrelationTerm(PairCorrlationTerm):
    """pf/pairCorrelationTerm.go:89-110."""

    def Construct(self, bricks):
        def fn(freq, t, out: np.ndarray):
            n = out.shape[0]
            out[:] = -self._multiplier(freq, n) * np.asarray(bricks[self.Field].Get(np.arange(n)))
            if self.Laplacian:
                LaplacianN(1).Eval(freq, out)

        return fn


class IdealMixtureTerm:
    """pf/pairCorrelationTerm.go:116-193 (mixed term)."""

    def __init__(self, IdealMix_: IdealMix, Field: str, Prefactor: float, Laplacian: bool):
        self.IdealMix = IdealMix_
        self.Field = Field
        self.Prefactor = Prefactor
        self.Laplacian = Laplacian

    def Eval(self, i, bricks):
        value = np.real(bricks[self.Field].Get(i))
        return (self.Prefactor * self.IdealMix.Deriv(value)) + 0j

    def ConstructLinear(self, bricks):
        def fn(freq, t, field: np.ndarray):
            field[:] = complex(self.Prefactor, 0.0)
            if self.Laplacian:
                LaplacianN(1).Eval(freq, field)

        return fn

    def nonLinearDerivedFieldName(self) -> str:
        return f"ideal_mixture_{self.Field}_nonlin"

    def DerivedField(self, num_nodes: int, bricks) -> DerivedField:
        def calc(out: np.ndarray):
            v = np.real(bricks[self.Field].Get(np.arange(out.shape[0])))
            out[:] = 3.0 * self.IdealMix.ThirdOrderPrefactor() * v * v + 4.0 * self.IdealMix.FourthOrderPrefactor() * v * v * v

        return DerivedField(np.zeros(num_nodes, dtype=np.complex128), self.nonLinearDerivedFieldName(), calc)

    def ConstructNonLinear(self, bricks):
        def fn(freq, t, field: np.ndarray):
            name = self.nonLinearDerivedFieldName()
            if name not in bricks:
                raise RuntimeError(f"Missing derived field {name}.")
            field[:] = bricks[name].Get(np.arange(field.shape[0]))
            if self.Laplacian:
                LaplacianN(1).Eval(freq, field)

        return fn

    def OnStepFinished(self, t, bricks):
        pass

    def GetEnergy(self, bricks, nodes: int) -> float:
        v = np.real(bricks[self.Field].Get(np.arange(nodes)))
        return float(np.sum(self.Prefactor * self.IdealMix.Eval(v)))


# ==========================================================================
# pf/homoLinElast.go
# ==========================================================================
def Indicator(x):
    """pf/homoLinElast.go:10-12."""
    return 3.0 * x * x - 2.0 * x * x * x


def IndicatorDeriv(x):
    """pf/homoLinElast.go:15-17."""
    return 6.0 * x - 6.0 * x * x


class HomogeneousModulusLinElast:
    """pf/homoLinElast.go:30-150 (Khachaturyan homogeneous-modulus driving force).
    ``Field`` is the real-space phase field of the PREVIOUS OnStepFinished (zeros before the
    first one, :145), so the term is identically zero during the first step."""

    def __init__(self, field_name: str, domain_size, mat_prop, misfit, workers: int = 1):
        from . import elasticity as el
        self._el = el
        self.FieldName = field_name
        self.Dim = len(domain_size)
        self.N = pfutil.prod_int(domain_size)
        self.MatProp = mat_prop
        self.Misfit = np.asarray(misfit, dtype=np.float64)
        self.EffForce = el.EffectiveForce(mat_prop, self.Misfit)
        self.Field = np.zeros(self.N, dtype=np.float64)
        self.Disps = el.Displacements
        self.FT = pfutil.NewFFTW(domain_size, workers)

    def Freq3(self) -> np.ndarray:
        return self._el.pad3(self.FT.freq_table())

    def Force(self, indicator: np.ndarray) -> np.ndarray:
        """:114-127 -- components i < Dim, frequencies padded to 3."""
        f3 = self.Freq3()
        res = np.zeros((self.N, 3), dtype=np.complex128)
        for i in range(self.Dim):
            res[:, i] = self.EffForce.Get(i, f3, indicator)
        return res

    def Construct(self, bricks):
        el = self._el

        def fn(freq, t, field: np.ndarray):
            field[:] = 0.0
            work = Indicator(self.Field).astype(np.complex128)
            self.FT.FFT(work)
            force = self.Force(work)
            f3 = self.Freq3()
            disp = self.Disps(force, f3, self.MatProp)
            dwork = IndicatorDeriv(self.Field).astype(np.complex128)
            A = self.MatProp.ContractLast(self.Misfit)
            for i in range(self.Dim):
                for j in range(i, self.Dim):
                    strains = np.ascontiguousarray(el.Strain(disp, f3, i, j))
                    self.FT.IFFT(strains)
                    strains /= complex(float(self.N), 0.0)
                    strains *= dwork
                    self.FT.FFT(strains)
                    factor = 1.0 if i == j else 2.0
                    field += complex(factor * A[i, j], 0.0) * strains
            e_density = el.EnergyDensity(self.MatProp, self.Misfit)
            work = (Indicator(self.Field) * IndicatorDeriv(self.Field)).astype(np.complex128)
            self.FT.FFT(work)
            field -= complex(2.0 * e_density, 0.0) * work

        return fn

    def OnStepFinished(self, t, bricks):
        self.Field[:] = np.real(bricks[self.FieldName].Get(np.arange(self.N)))


def NewHomogeneousModolus(field_name, domain_size, mat_prop, misfit, workers: int = 1):
    return HomogeneousModulusLinElast(field_name, domain_size, mat_prop, misfit, workers)


# ==========================================================================
# the remaining catalog (oracle first; no device implementation yet)
# pf/gradientCalculator.go, pf/advection.go, pf/sourceTerm.go, pf/negative_value_penalty.go
# ==========================================================================
class GradientCalculator:
    """pf/gradientCalculator.go:19-58: d/dx_Comp of a real-space array, pseudo-spectrally; the
    +0.5 Nyquist frequency is zeroed unless KeepNyquist."""

    def __init__(self, FT, Comp: int, KeepNyquist: bool = False):
        self.FT, self.Comp, self.KeepNyquist = FT, Comp, KeepNyquist

    def Calculate(self, indata: np.ndarray, data: np.ndarray):
        data[:] = indata
        self.FT.FFT(data)
        f = as_frequency(self.FT.Freq).table(data.shape[0])[:, self.Comp].copy()
        if not self.KeepNyquist:
            f[np.abs(f - 0.5) < 1e-10] = 0.0
        data *= 1j * 2.0 * math.pi * f
        self.FT.IFFT(data)
        data /= float(data.shape[0])

    def ToDerivedField(self, name: str, N: int, brick):
        from .pf import DerivedField  # local import: pf imports nothing from terms

        def calc(data: np.ndarray):
            data[:] = brick.Get(np.arange(N))
            self.Calculate(data.copy(), data)

        return DerivedField(np.zeros(N, dtype=np.complex128), name, calc)


class DivGrad:
    """pf/gradientCalculator.go:60-134: div(F grad field), F a function of the bricks."""

    def __init__(self, Field: str, F):
        self.Field, self.F = Field, F

    def FuncName(self) -> str:
        return f"DivGrad_{self.Field}_Func"

    def GradName(self, comp: int) -> str:
        return f"GRAD_{self.Field}_{comp}"

    def PrepareModel(self, N: int, m, FT):
        from .pf import DerivedField
        dim = len(FT.Freq(0))
        for d in range(dim):
            grad = GradientCalculator(FT, d, False)
            m.RegisterDerivedField(grad.ToDerivedField(self.GradName(d), N, m.Bricks[self.Field]))

            def calc(data: np.ndarray, d=d):
                idx = np.arange(N)
                data[:] = self.F(idx, m.Bricks) * m.Bricks[self.GradName(d)].Get(idx)

            m.RegisterDerivedField(DerivedField(np.zeros(N, dtype=np.complex128), self.FuncName() + self.GradName(d), calc))

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            f = as_frequency(freq).table(field.shape[0])
            field[:] = 0.0
            for d in range(f.shape[1]):
                field += (1j * 2.0 * math.pi * f[:, d]) * bricks[self.FuncName() + self.GradName(d)].Get(np.arange(field.shape[0]))

        return fn

    def OnStepFinished(self, t, bricks=None):
        pass


class WeightedLaplacian:
    """pf/gradientCalculator.go:114-172: F(c) LAP field.  Field and PreFactor name bricks that hold SPECTRA when the
    closure runs (inside a step every field and derived field is transformed, euler.go:18-26)."""

    def __init__(self, Field: str, PreFactor: str, FT):
        self.Field, self.PreFactor, self.FT = Field, PreFactor, FT

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            n = field.shape[0]
            idx = np.arange(n)
            field[:] = bricks[self.Field].Get(idx)
            f = as_frequency(freq).table(n)
            field *= -(2.0 * math.pi * np.sqrt(np.sum(f * f, axis=1))) ** 2  # LaplacianN{Power: 1}.Eval, diffOp.go:25-30
            self.FT.IFFT(field)
            field /= float(n)
            work = np.array(bricks[self.PreFactor].Get(idx), dtype=np.complex128)
            self.FT.IFFT(work)
            work /= float(n)
            field *= work
            self.FT.FFT(field)

        return fn

    def OnStepFinished(self, t, bricks=None):
        pass


class Advection:
    """pf/advection.go:10-96: -(v . grad field) through gradient derived fields."""

    def __init__(self, Field: str, VelocityFields):
        self.Field, self.VelocityFields = Field, list(VelocityFields)

    def GradName(self, comp: int) -> str:
        return f"{self.Field}_{comp}"

    def GetName(self) -> str:
        return "".join(self.VelocityFields) + "DotGrad" + self.Field

    def AllFieldsExist(self, m) -> bool:
        return all(m.IsBrickName(v) for v in self.VelocityFields) and m.IsBrickName(self.Field)

    def PrepareModel(self, N: int, m, FT):
        from .pf import DerivedField
        dim = len(FT.Freq(0))
        if len(self.VelocityFields) != dim:
            raise RuntimeError("Advection: Inconsistent number of velocity fields")
        if not self.AllFieldsExist(m):
            raise RuntimeError("Advection: Make sure that field and all the velocity fields are added to the model")
        for d in range(dim):
            m.RegisterDerivedField(GradientCalculator(FT, d).ToDerivedField(self.GradName(d), N, m.Bricks[self.Field]))

        def calc(data: np.ndarray):
            idx = np.arange(N)
            data[:] = 0.0
            for d in range(len(self.VelocityFields)):
                data += m.Bricks[self.VelocityFields[d]].Get(idx) * m.Bricks[self.GradName(d)].Get(idx)

        m.RegisterDerivedField(DerivedField(np.zeros(N, dtype=np.complex128), self.GetName(), calc))

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            field[:] = -bricks[self.GetName()].Get(np.arange(field.shape[0]))

        return fn

    def OnStepFinished(self, t, bricks=None):
        pass


class Source:
    """pf/sourceTerm.go:10-30: f(t) exp(-2 pi i k . pos) -- the transform of a point source."""

    def __init__(self, pos, f):
        self.Pos = [float(p) for p in pos]
        self.f = f

    def Eval(self, freq, t: float, data: np.ndarray):
        k = as_frequency(freq).table(data.shape[0])
        data[:] = complex(self.f(t), 0.0) * np.exp(-1j * 2.0 * math.pi * (k @ np.asarray(self.Pos)))


def NewSource(pos, f) -> Source:
    return Source(pos, f)


class NegativeValuePenalty:
    """pf/negative_value_penalty.go:5-38."""

    def __init__(self, Prefactor: float, Exponent: int, Field: str):
        self.Prefactor, self.Exponent, self.Field = Prefactor, Exponent, Field

    def Penalty(self, x):
        x = np.asarray(x, dtype=np.float64)
        p = float(self.Exponent)
        with np.errstate(invalid="ignore"):
            val = -2.0 * self.Prefactor * p * np.power(np.where(x > 0.0, 1.0, x), p - 1.0)
        return np.where(x > 0.0, 0.0, val)

    def Evaluate(self, i, bricks):
        return self.Penalty(np.real(bricks[self.Field].Get(i))).astype(np.complex128)


def NewDefaultNegativeValuePenalty(field: str) -> NegativeValuePenalty:
    return NegativeValuePenalty(1500.0, 3, field)


# ==========================================================================
# pf/chargeTransport.go
# ==========================================================================
VOIGT_3D = ((0, 5, 4), (5, 1, 3), (4, 3, 2))
VOIGT_2D = ((0, 2), (2, 1))


def voigtIndex(i: int, j: int, dim: int) -> int:
    """pf/chargeTransport.go:151-171."""
    return VOIGT_2D[i][j] if dim == 2 else VOIGT_3D[i][j]


class ChargeTransport:
    """pf/chargeTransport.go:9-149: d rho/dt += div(sigma (grad phi - E_ext)) with the potential from
    Poisson's equation in k-space.  ``Conductivity(i)`` returns the Voigt components at node i
    (vectorised here: an (N, n_voigt) array for an index array)."""

    def __init__(self, Conductivity, ExternalField, Field: str, FT):
        self.Conductivity, self.ExternalField, self.Field, self.FT = Conductivity, list(ExternalField), Field, FT

    def _sigma(self, N: int, dim: int) -> np.ndarray:
        s = self.Conductivity(np.arange(N))
        s = np.asarray(s, dtype=np.float64)
        if s.ndim == 1:  # a constant tensor
            s = np.broadcast_to(s, (N, s.shape[0]))
        return s

    def current(self, brick, N: int) -> np.ndarray:
        """:38-71 -- returns effCurrent as (dim, N)."""
        k = as_frequency(self.FT.Freq).table(N)
        dim = k.shape[1]
        k_sq = np.sum(k * k, axis=1)
        sigma = self._sigma(N, dim)
        rho = brick.Get(np.arange(N))
        eff_current = np.zeros((dim, N), dtype=np.complex128)
        for d in range(dim):
            keep = np.abs(np.abs(k[:, d]) - 0.5) > 1e-10
            eff_field = np.where(keep, rho * (1j * k[:, d] / (2.0 * math.pi * k_sq + 1e-16)), 0.0).astype(np.complex128)
            self.FT.IFFT(eff_field)
            eff_field /= float(N)
            eff_field -= complex(self.ExternalField[d], 0.0)
            for d2 in range(dim):
                eff_current[d2] += sigma[:, voigtIndex(d, d2, dim)] * eff_field
        return eff_current

    def Construct(self, bricks):
        def fn(freq, t, field: np.ndarray):
            N = field.shape[0]
            k = as_frequency(freq).table(N)
            dim = k.shape[1]
            field[:] = 0.0
            eff_current = self.current(bricks[self.Field], N)
            for d2 in range(dim):
                work = np.ascontiguousarray(eff_current[d2])
                self.FT.FFT(work)
                keep = np.abs(np.abs(k[:, d2]) - 0.5) > 1e-10
                field += np.where(keep, (1j * 2.0 * math.pi * k[:, d2]) * work, 0.0)

        return fn

    def Current(self, density, N: int, realspace: bool):
        """:104-146."""
        if realspace:
            from .pf import NewField
            r = np.array(density.Get(np.arange(N)), dtype=np.complex128)
            self.FT.FFT(r)
            density = NewField("ftDensity", N, r)
        cur = self.current(density, N)
        return [-cur[d].real for d in range(cur.shape[0])]

    def OnStepFinished(self, t, bricks):
        pass
