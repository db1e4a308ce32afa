"""Mirror of the pieces of the reference's ``elasticity`` package that parametrise
``pf.HomogeneousModulusLinElast`` (elasticity/rank4.go, linearElasticity.go:86-98), over the
C ABI (include/gopf_cuda.h).  Host-side tensor bookkeeping only; the Khachaturyan operator
itself runs inside the device step (gopf_b200/csrc/elastic.cuh)."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, lib

_PD = ctypes.POINTER(ctypes.c_double)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_PD)


class Rank4:
    """elasticity.Rank4 (rank4.go:10-75): Data[i*27 + j*9 + k*3 + l]."""

    def __init__(self, data=None):
        self.Data = np.zeros(81, dtype=np.float64) if data is None else np.ascontiguousarray(data, dtype=np.float64)

    def At(self, i, j, k, l):
        return float(self.Data[i * 27 + j * 9 + k * 3 + l])

    def Set(self, i, j, k, l, v):
        self.Data[i * 27 + j * 9 + k * 3 + l] = v

    def Rotate(self, rot):
        r = np.ascontiguousarray(rot, dtype=np.float64).reshape(9)
        check(lib().gopf_elasticity_rotate(_p(self.Data), _p(r)))

    def ContractLast(self, tensor) -> np.ndarray:
        t = np.ascontiguousarray(tensor, dtype=np.float64).reshape(9)
        out = np.zeros(9, dtype=np.float64)
        check(lib().gopf_elasticity_contract_last(_p(self.Data), _p(t), _p(out)))
        return out.reshape(3, 3)


def CubicMaterial(c11: float, c12: float, c44: float) -> Rank4:
    t = Rank4()
    check(lib().gopf_elasticity_cubic_material(ctypes.c_double(c11), ctypes.c_double(c12), ctypes.c_double(c44), _p(t.Data)))
    return t


def Isotropic(bulkMod: float, poisson: float) -> Rank4:
    t = Rank4()
    check(lib().gopf_elasticity_isotropic(ctypes.c_double(bulkMod), ctypes.c_double(poisson), _p(t.Data)))
    return t


def EnergyDensity(matProp: Rank4, strain) -> float:
    e = np.ascontiguousarray(strain, dtype=np.float64).reshape(9)
    out = ctypes.c_double()
    check(lib().gopf_elasticity_energy_density(_p(matProp.Data), _p(e), ctypes.byref(out)))
    return out.value


def KhachaturyanMultiplier(matProp: Rank4, misfit, dim: int, freq3) -> np.ndarray:
    """M(k) (csrc/elastic.cuh) for an (n, 3) array of [f_row, f_col, f_depth]."""
    f = np.ascontiguousarray(freq3, dtype=np.float64).reshape(-1, 3)
    mis = np.ascontiguousarray(misfit, dtype=np.float64).reshape(9)
    out = np.zeros(f.shape[0], dtype=np.float64)
    check(lib().gopf_elasticity_multiplier(_p(matProp.Data), _p(mis), int(dim), _p(f), ctypes.c_int64(f.shape[0]), _p(out)))
    return out


def StrainFactor(matProp: Rank4, misfit, freq3, i: int, j: int) -> np.ndarray:
    """s_ij(k) with eps^_ij = s_ij H^ (csrc/elastic_energy.cu; Displacements + Strain,
    elasticity/linearElasticity.go:16-83) for an (n, 3) array of padded frequencies, on the host."""
    f = np.ascontiguousarray(freq3, dtype=np.float64).reshape(-1, 3)
    mis = np.ascontiguousarray(misfit, dtype=np.float64).reshape(9)
    out = np.zeros(f.shape[0], dtype=np.float64)
    check(lib().gopf_elasticity_strain_factor(_p(matProp.Data), _p(mis), _p(f), ctypes.c_int64(f.shape[0]), int(i), int(j), _p(out)))
    return out


def HomogeneousModulusEnergy(indicator, domainSize, misfit, matProp: Rank4, device: int = -1) -> float:
    """elasticity.HomogeneousModulusEnergy (elasticity/linearElasticity.go:101-165) on the device."""
    ind = np.ascontiguousarray(indicator, dtype=np.complex128)
    mis = np.ascontiguousarray(misfit, dtype=np.float64).reshape(9)
    dims = (ctypes.c_int * len(domainSize))(*[int(v) for v in domainSize])
    out = ctypes.c_double()
    check(lib().gopf_elasticity_homogeneous_modulus_energy(len(domainSize), dims, _p(ind.view(np.float64)), _p(mis),
                                                           _p(matProp.Data), int(device), ctypes.byref(out)))
    return out.value
