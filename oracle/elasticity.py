"""Oracle restatement of the reference's ``elasticity`` package pieces behind
``pf.HomogeneousModulusLinElast``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations: /root/reference.
gonum ``mat.Dense.Solve`` (LU, partial pivoting) is restated by ``numpy.linalg.solve``.
"""
from __future__ import annotations

import math

import numpy as np

from . import pfutil


class Rank4:
    """elasticity/rank4.go:10-75: 81 doubles, index i*27 + j*9 + k*3 + l."""

    def __init__(self, data=None):
        self.Data = np.zeros(81, dtype=np.float64) if data is None else np.array(data, dtype=np.float64)

    def At(self, i, j, k, l):
        return float(self.Data[i * 27 + j * 9 + k * 3 + l])

    def Set(self, i, j, k, l, v):
        self.Data[i * 27 + j * 9 + k * 3 + l] = v

    def tensor(self) -> np.ndarray:
        return self.Data.reshape(3, 3, 3, 3)

    def Rotate(self, rot: np.ndarray):
        """rank4.go:39-60."""
        c = self.tensor()
        self.Data = np.einsum("im,jn,kp,lq,mnpq->ijkl", rot, rot, rot, rot, c).reshape(81)

    def ContractLast(self, tensor: np.ndarray) -> np.ndarray:
        """rank4.go:62-75: out_ij = sum_kl C_ijkl t_kl."""
        return np.einsum("ijkl,kl->ij", self.tensor(), np.asarray(tensor, dtype=np.float64))


def Shear(bulk_mod: float, poisson: float) -> float:
    return 3.0 * bulk_mod * (1.0 - 2.0 * poisson) / (2.0 * (1.0 + poisson))


def Isotropic(bulk_mod: float, poisson: float) -> Rank4:
    """rank4.go:84-110."""
    shear = Shear(bulk_mod, poisson)
    t = Rank4()
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    v = 0.0
                    if i == j and k == l:
                        v += bulk_mod - 2.0 * shear / 3.0
                    if i == k and j == l:
                        v += shear
                    if i == l and j == k:
                        v += shear
                    t.Set(i, j, k, l, v)
    return t


def CubicMaterial(c11: float, c12: float, c44: float) -> Rank4:
    """rank4.go:113-128 (as written: only C_ijij and C_jiji carry c44)."""
    t = Rank4()
    for i in range(3):
        t.Set(i, i, i, i, c11)
    for i in range(3):
        for j in range(i + 1, 3):
            t.Set(i, i, j, j, c12)
            t.Set(j, j, i, i, c12)
            t.Set(i, j, i, j, c44)
            t.Set(j, i, j, i, c44)
    return t


def RotationMatrix(angle: float, axis: int) -> np.ndarray:
    """rank4.go:157-182."""
    c, s = math.cos(angle), math.sin(angle)
    rot = np.zeros((3, 3))
    if axis == 0:
        rot[0, 0] = 1.0; rot[1, 1] = c; rot[2, 2] = c; rot[1, 2] = s; rot[2, 1] = -s
    elif axis == 1:
        rot[1, 1] = 1.0; rot[0, 0] = c; rot[2, 2] = c; rot[0, 2] = s; rot[2, 0] = -s
    elif axis == 2:
        rot[2, 2] = 1.0; rot[0, 0] = c; rot[1, 1] = c; rot[0, 1] = s; rot[1, 0] = -s
    return rot


def pad3(f: np.ndarray) -> np.ndarray:
    """HomogeneousModulusLinElast.Freq (pf/homoLinElast.go:103-111): frequencies padded to 3."""
    if f.shape[1] == 3:
        return f
    out = np.zeros((f.shape[0], 3), dtype=np.float64)
    out[:, :f.shape[1]] = f
    return out


def Displacements(ft_body_force: np.ndarray, freq3: np.ndarray, mat_prop: Rank4) -> np.ndarray:
    """elasticity/linearElasticity.go:16-63.  ft_body_force: (N, 3) complex; freq3: (N, 3).
    G_mn = (2 pi)^2 sum_jl C_mjnl f_j f_l; solve G u = F (real and imaginary parts as two
    right-hand sides); modes with every |f_c| < 1e-10 give u = 0."""
    c = mat_prop.tensor()
    g = np.einsum("mjnl,kj,kl->kmn", c, freq3, freq3) * math.pow(2.0 * math.pi, 2)
    zero = np.all(np.abs(freq3) < 1e-10, axis=1)
    g[zero] = np.eye(3)
    rhs = np.stack([ft_body_force.real, ft_body_force.imag], axis=2)  # (N, 3, 2)
    sol = np.linalg.solve(g, rhs)
    disp = sol[:, :, 0] + 1j * sol[:, :, 1]
    disp[zero] = 0.0
    return disp


def Strain(ft_disp: np.ndarray, freq: np.ndarray, m: int, n: int) -> np.ndarray:
    """elasticity/linearElasticity.go:67-83: i pi (f_n u_m + f_m u_n), |f| = 1/2 zeroed."""
    fm = freq[:, m].copy()
    fn = freq[:, n].copy()
    fm[np.abs(np.abs(fm) - 0.5) < 1e-10] = 0.0
    fn[np.abs(np.abs(fn) - 0.5) < 1e-10] = 0.0
    return (1j * math.pi * fn) * ft_disp[:, m] + (1j * math.pi * fm) * ft_disp[:, n]


def EnergyDensity(mat_prop: Rank4, strain: np.ndarray) -> float:
    """elasticity/linearElasticity.go:86-98."""
    return 0.5 * float(np.einsum("ijkl,ij,kl->", mat_prop.tensor(), strain, strain))


class EffectiveForce:
    """elasticity/effectiveForce.go:9-35."""

    def __init__(self, mat_prop: Rank4, misfit: np.ndarray):
        self.EffStress = mat_prop.ContractLast(misfit)

    def Get(self, comp: int, freq: np.ndarray, indicator: np.ndarray) -> np.ndarray:
        force = np.zeros(indicator.shape[0], dtype=np.complex128)
        for j in range(freq.shape[1]):
            force += (1j * (-self.EffStress[comp, j] * 2.0 * math.pi * freq[:, j])) * indicator
        return force


def Ellipsoid(N: int, a: float, b: float, c: float) -> np.ndarray:
    """elasticity/shapes.go:9-26 (voxel indicator, flattened row-major)."""
    i, j, k = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    v = ((i - N // 2) / a) ** 2 + ((j - N // 2) / b) ** 2 + ((k - N // 2) / c) ** 2
    return (v <= 1.0).astype(np.complex128).reshape(-1)


def HomogeneousModulusEnergy(indicator: np.ndarray, domain_size, misfit: np.ndarray, mat_prop: Rank4) -> float:
    """elasticity/linearElasticity.go:101-165 (energy per unit precipitate volume)."""
    indicator = np.array(indicator, dtype=np.complex128)
    volume = float(indicator.real.sum())
    eff = EffectiveForce(mat_prop, misfit)
    ft = pfutil.NewFFTW(domain_size)
    ft.FFT(indicator)
    f = ft.freq_table()
    force = np.stack([eff.Get(c, f, indicator) for c in range(3)], axis=1)
    f3 = pad3(f)
    disp = Displacements(force, f3, mat_prop)
    ft.IFFT(indicator)
    indicator /= indicator.shape[0]
    n = indicator.shape[0]
    strains = np.zeros((n, 3, 3))
    inside = indicator.real > 0.5
    for i in range(3):
        for j in range(i, 3):
            s = Strain(disp, f3, i, j)
            ft.IFFT(s)
            re = s.real / float(n)
            re = re - np.where(inside, misfit[i, j], 0.0)
            strains[:, i, j] = re
            strains[:, j, i] = re
    energy = 0.5 * float(np.einsum("ijkl,nij,nkl->", mat_prop.tensor(), strains, strains))
    return energy / volume


def EshelbyEnergyDensityDilatational(poisson, shear, misfit):
    """elasticity/linearElasticity_test.go:109-111."""
    return 2.0 * (1.0 + poisson) * shear * misfit * misfit / (1.0 - poisson)
