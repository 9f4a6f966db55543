"""Synthetic initial conditions of the reference's test drivers (closed-form,
no external data), vectorised over arrays of points.

  ShallowWaterTestCase2   reference test/shallowwater_sphere/SWTest2.cpp:33-159
  BaroclinicWaveJWTest    reference test/nonhydro_sphere/BaroclinicWaveJWTest.cpp:27-413

Interface follows TestCase (reference src/atm/TestCase.h:33-173):
evaluate_topography(phys, lon, lat) and evaluate_pointwise_state(phys, z, lon,
lat) -> list of components in the order of the EquationSet (U, V zonal /
meridional in m/s; the grid converts them to covariant components).
"""
import math

import numpy as np


class ShallowWaterTestCase2:
    """Williamson et al. (1992) test 2: steady geostrophic flow."""

    equation_set = "shallow_water"
    ztop = 1.0

    def __init__(self, h0=2998.104995, u0=38.61068277, alpha=0.0):
        self.h0, self.u0, self.alpha = h0, u0, alpha

    def evaluate_topography(self, phys, lon, lat):
        return np.zeros_like(lon)

    def evaluate_pointwise_state(self, phys, z, lon, lat):
        lat = np.where(np.abs(lat - 0.5 * math.pi) < 1.0e-12, lat - 1.0e-12, lat)
        lat = np.where(np.abs(lat + 0.5 * math.pi) < 1.0e-12, lat + 1.0e-12, lat)
        a = self.alpha
        u = self.u0 * np.cos(lat) * (math.cos(a) + np.cos(lon) * np.tan(lat) * math.sin(a))
        v = -self.u0 * np.sin(lon) * math.sin(a)
        htrig = -np.cos(lon) * np.cos(lat) * math.sin(a) + np.sin(lat) * math.cos(a)
        h = self.h0 - (phys.earth_radius * phys.omega + 0.5 * self.u0) \
            * self.u0 * htrig * htrig / phys.g
        return [u, v, h]


class BaroclinicWaveJWTest:
    """Jablonowski & Williamson (2006) baroclinic wave."""

    equation_set = "primitive_nonhydro"

    def __init__(self, alpha=0.0, ztop=30000.0, perturbation="exp"):
        self.alpha = alpha
        self.ztop = ztop
        self.perturbation = perturbation.lower()
        self.eta0 = 0.252
        self.tropopause_eta = 0.2
        self.t0 = 288.0
        self.delta_t = 4.8e5
        self.lapse_rate = 0.005
        self.u0 = 35.0
        self.up = 1.0
        self.pert_lon = math.pi / 9.0
        self.pert_lat = 2.0 * math.pi / 9.0
        self.pert_r = 0.1

    def _profiles(self, phys, aux_eta, lat):
        s = np.sin(lat)
        s2 = s * s
        s6 = s2 * s2 * s2
        c = np.cos(lat)
        c2 = c * c
        c3 = c2 * c
        p1 = self.u0 * np.cos(aux_eta) ** 1.5 * (-2.0 * s6 * (c2 + 1.0 / 3.0) + 10.0 / 63.0)
        p2 = phys.earth_radius * phys.omega * (8.0 / 5.0 * c3 * (s2 + 2.0 / 3.0) - 0.25 * math.pi)
        return p1, p2

    def evaluate_topography(self, phys, lon, lat):
        aux = 0.5 * math.pi * (1.0 - self.eta0)
        p1, p2 = self._profiles(phys, aux, lat)
        return self.u0 * math.cos(aux) ** 1.5 * (p1 + p2) / phys.g

    def geopotential_temperature(self, phys, eta, lat):
        aux = 0.5 * math.pi * (eta - self.eta0)
        expo = phys.R * self.lapse_rate / phys.g
        tavg = self.t0 * eta ** expo
        strat = eta < self.tropopause_eta
        te = self.tropopause_eta
        tavg = tavg + np.where(strat, self.delta_t * np.maximum(te - eta, 0.0) ** 5.0, 0.0)
        p1, p2 = self._profiles(phys, aux, lat)
        temp = 2.0 * p1 + p2
        temp = tavg + 0.75 * eta * math.pi * self.u0 / phys.R \
            * np.sin(aux) * np.sqrt(np.cos(aux)) * temp
        gavg = self.t0 * phys.g / self.lapse_rate * (1.0 - eta ** expo)
        es = np.where(strat, eta, te)
        corr = phys.R * self.delta_t * (
            (np.log(es / te) + 137.0 / 60.0) * te ** 5
            - 5.0 * te ** 4 * es + 5.0 * te ** 3 * es ** 2
            - (10.0 / 3.0) * te ** 2 * es ** 3 + 5.0 / 4.0 * te * es ** 4
            - 1.0 / 5.0 * es ** 5)
        gavg = gavg - np.where(strat, corr, 0.0)
        geo = gavg + self.u0 * np.cos(aux) ** 1.5 * (p1 + p2)
        return geo, temp

    def eta_from_rll(self, phys, z, lat):
        """Newton iteration of EtaFromRLL (:297-345)."""
        eta = np.full(np.broadcast(z, lat).shape, 1.0e-7)
        lat = np.broadcast_to(lat, eta.shape)
        z = np.broadcast_to(z, eta.shape)
        done = np.zeros(eta.shape, dtype=bool)
        geo = temp = None
        for _ in range(25):
            geo, temp = self.geopotential_temperature(phys, eta, lat)
            f = -phys.g * z + geo
            df = -phys.R / eta * temp
            new = eta - f / df
            conv = np.abs(eta - new) < 1.0e-13
            eta = np.where(done, eta, new)
            done = done | conv
            if done.all():
                break
        if not done.all():
            raise RuntimeError("Maximum number of iterations exceeded.")
        # the reference returns geopotential / temperature of the last
        # evaluated iterate
        return eta, temp

    def evaluate_reference_state(self, phys, z, lon, lat):
        eta, temp = self.eta_from_rll(phys, z, lat)
        ulon = self.u0 * np.cos(0.5 * math.pi * (eta - self.eta0)) ** 1.5 \
            * np.sin(2.0 * lat) * np.sin(2.0 * lat)
        p = phys.p0 * eta
        rho = p / (phys.R * temp)
        rhotheta = phys.rho_theta_from_pressure(p)
        zero = np.zeros_like(ulon)
        return [ulon, zero, rhotheta / rho, zero, rho]

    def evaluate_pointwise_state(self, phys, z, lon, lat):
        st = self.evaluate_reference_state(phys, z, lon, lat)
        if self.perturbation == "exp":
            lonb = np.broadcast_to(lon, st[0].shape)
            latb = np.broadcast_to(lat, st[0].shape)
            r = np.arccos(np.clip(
                math.sin(self.pert_lat) * np.sin(latb)
                + math.cos(self.pert_lat) * np.cos(latb) * np.cos(lonb - self.pert_lon),
                -1.0, 1.0))
            r = r / self.pert_r
            st[0] = st[0] + np.where(r < 1.0, self.up * np.exp(-r * r), 0.0)
        return st
