"""TEST INFRASTRUCTURE -- CPU restatement of pf/sdd.go (Shrinking-Dimer-Dynamics saddle-point
stepper).  numpy + the oracle's FFTWWrapper.  Every function cites the
reference lines it follows.  Pinned by tests/test_oracle_sdd.py to pf/sdd_test.go.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nothing under gopf_b200/ does.
"""
from __future__ import annotations

import math
from typing import List

import numpy as np

from . import pfutil


class SDDTimeConstants:
    """pf/sdd.go:14-23."""

    def __init__(self, Orientation: float = 1.0, DimerLength: float = 1.0):
        self.Orientation, self.DimerLength = Orientation, DimerLength


class SDDMonitor:
    """pf/sdd.go:25-53 (the csv logger, :55-84, is host I/O and not restated)."""

    def __init__(self):
        self.MaxForce = 0.0
        self.ForcePowerSpectrum = 0.0
        self.MaxTorque = 0.0
        self.FieldNorm = 0.0
        self.FieldNormChange = 0.0


class diagonalShermannMorrison:
    """pf/sdd.go:445-470: (D + u v^T)^-1 b with diagonal D."""

    def __init__(self, invDiagonal, u, v):
        self.invDiagonal, self.u, self.v = invDiagonal, u, v

    def dot(self, vec: np.ndarray):
        denum = complex(1.0, 0.0) + np.sum(self.u * self.invDiagonal * self.v)
        v_dot = np.sum(self.v * self.invDiagonal * vec)
        vec[:] = self.invDiagonal * vec - self.invDiagonal * self.u * v_dot / denum


class SDD:
    """pf/sdd.go:86-443."""

    def __init__(self, domain_size, model):
        # NewSDD (:120-129)
        self.TimeConstants = SDDTimeConstants(1.0, 1.0)
        self.Alpha = 0.5
        self.Dt = 0.0
        self.CurrentStep = 0
        self.MinDimerLength = 0.0
        self.Monitor = SDDMonitor()
        self.InitDimerLength = 0.0
        self.orientation = np.zeros(model.NumNodes() * len(model.Fields), dtype=np.float64)
        self.ft = pfutil.NewFFTW(domain_size)
        self.initialized = False

    # :131-135
    def checkTimeStep(self):
        if self.Dt < 1e-16:
            raise RuntimeError("Timestep not set in SDD. Make sure that the Dt attribute has explicitly been set.")

    # :139-147
    def fft(self, m):
        m.SyncDerivedFields()
        for f in m.Fields:
            self.ft.FFT(f.Data)
        for f in m.DerivedFields:
            self.ft.FFT(f.Data)

    # :150-155
    def ifft(self, m):
        for f in m.Fields:
            self.ft.IFFT(f.Data)
            f.Data /= float(f.Data.shape[0])

    # :158-300
    def Step(self, m):
        if not self.initialized:
            raise RuntimeError("SDD: The method have to be initialized first. See SDD.Init\n")
        self.checkTimeStep()
        N = m.NumNodes()
        F = len(m.Fields)

        fnorm = self.FieldNorm(m.Fields)
        self.Monitor.FieldNormChange = self.Monitor.FieldNorm - fnorm
        self.Monitor.FieldNorm = fnorm

        ft_orientation = self.orientation.astype(np.complex128)
        for i in range(F):
            self.ft.FFT(ft_orientation[i * N:(i + 1) * N])

        rhs_start = np.zeros(F * N, dtype=np.complex128)
        rhs_end = np.zeros(F * N, dtype=np.complex128)
        l = self.DimerLength(self.GetTime())

        self.ShiftFieldsAlongDimer(m.Fields, -0.5 * l)  # to the start image
        self.fft(m)
        self.extractRHS(m, rhs_start)

        self.ifft(m)  # to the end image
        self.ShiftFieldsAlongDimer(m.Fields, l)
        self.fft(m)
        self.extractRHS(m, rhs_end)

        self.ifft(m)  # back to the centre
        self.ShiftFieldsAlongDimer(m.Fields, -0.5 * l)
        self.fft(m)

        self.Monitor.MaxForce = 0.0
        self.Monitor.ForcePowerSpectrum = 0.0
        c_dt = complex(self.Dt, 0.0)
        for i in range(F):
            orig_field = m.Fields[i].Data.copy()
            # :204-206 -- as written the weighted force always reads the FIRST field's block
            # (index j, not i*N + j)
            work = complex(self.Alpha, 0.0) * rhs_start[:N] + complex(1.0 - self.Alpha, 0.0) * rhs_end[:N]
            d = m.Fields[i].Data
            active = ft_orientation[i * N:(i + 1) * N]
            self.householder(work, active, 2.0, N)
            d += c_dt * work  # :212-214
            work = m.GetDenum(i, self.ft.Freq, self.GetTime())
            dsm = self.householderDenum(work, active, 2.0)
            dsm.dot(d)
            diff = np.abs((d - orig_field) / c_dt)  # :219-225
            self.Monitor.ForcePowerSpectrum += float(np.sum(diff * diff / float(N)))
            self.Monitor.MaxForce = max(self.Monitor.MaxForce, float(np.max(diff)))
        self.Monitor.ForcePowerSpectrum /= float(F * N)
        self.Monitor.ForcePowerSpectrum = math.sqrt(self.Monitor.ForcePowerSpectrum)

        torque = rhs_start  # shares storage (:232)
        torque -= rhs_end
        for i in range(F):  # :241-248
            denum = m.GetDenum(i, self.ft.Freq, self.GetTime())
            torque[i * N:(i + 1) * N] -= denum * ft_orientation[i * N:(i + 1) * N] * complex(l, 0.0)

        self.householder(torque, ft_orientation, 1.0, N)  # :260
        for i in range(F):
            self.ft.IFFT(torque[i * N:(i + 1) * N])
        torque /= float(N)

        # :277-283
        coef = self.Dt / (self.DimerLength(self.GetTime()) * self.TimeConstants.Orientation)
        self.orientation -= coef * torque.real
        self.Monitor.MaxTorque = float(np.max(np.abs(torque.real))) if torque.size else 0.0
        length = math.sqrt(float(np.dot(self.orientation, self.orientation)))  # :286-289
        self.orientation /= length

        self.ifft(m)
        self.CurrentStep += 1

    # :302-307
    def extractRHS(self, m, rhs: np.ndarray):
        N = m.NumNodes()
        for i in range(len(m.Fields)):
            rhs[i * N:(i + 1) * N] = m.GetRHS(i, self.ft.Freq, self.GetTime())

    # :312-324
    def householder(self, data: np.ndarray, ft_orientation: np.ndarray, sigma: float, num_nodes: int):
        dot = np.sum(data * np.conj(ft_orientation)) / complex(float(num_nodes), 0.0)
        data -= complex(sigma, 0.0) * ft_orientation * dot

    # :327-344
    def householderDenum(self, data: np.ndarray, ft_orientation: np.ndarray, sigma: float) -> diagonalShermannMorrison:
        c_dt = complex(self.Dt, 0.0)
        inv_diag = 1.0 / (1.0 - c_dt * data)
        v = complex(sigma, 0.0) * data * c_dt * np.conj(ft_orientation) / complex(float(data.shape[0]), 0.0)
        return diagonalShermannMorrison(inv_diag, ft_orientation, v)

    # :348-356
    def ShiftFieldsAlongDimer(self, fields, scale: float):
        outer = 0
        for f in fields:
            n = f.Data.shape[0]
            f.Data += scale * self.orientation[outer:outer + n]
            outer += n

    # :359-361
    def GetTime(self) -> float:
        return float(self.CurrentStep) * self.Dt

    # :364-370
    def DimerLength(self, t: float) -> float:
        l = self.InitDimerLength * math.exp(-t / self.TimeConstants.DimerLength)
        return self.MinDimerLength if l < self.MinDimerLength else l

    # :373-375
    def RequiredDimerLengthTime(self, l: float) -> float:
        return self.TimeConstants.DimerLength * math.log(self.InitDimerLength / l)

    # :378-387
    def FieldNorm(self, fields) -> float:
        return float(sum(np.sum(np.abs(f.Data) ** 2) for f in fields))

    # :390-408
    def Init(self, init: List, final: List):
        outer = 0
        for a, b in zip(init, final):
            n = a.Data.shape[0]
            self.orientation[outer:outer + n] = (b.Data - a.Data).real
            outer += n
        self.InitDimerLength = math.sqrt(float(np.dot(self.orientation, self.orientation)))
        self.orientation /= self.InitDimerLength
        self.initialized = True

    # :413-427
    def SetInitialOrientation(self, orient):
        orient = np.asarray(orient, dtype=np.float64)
        if orient.shape[0] != self.orientation.shape[0]:
            raise RuntimeError("Inconsistent length of the passed orientaiton vector")
        self.InitDimerLength = math.sqrt(float(np.dot(orient, orient)))
        self.orientation[:] = orient / self.InitDimerLength
        self.initialized = True

    # :431-433
    def SetFilter(self, filt):
        raise RuntimeError("SDD: Does not support modal filters")


def NewSDD(domain_size, model) -> SDD:
    return SDD(domain_size, model)
