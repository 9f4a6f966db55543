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


class _Numpy:
    """array namespace used by the closed-form initial conditions; a torch
    twin (below) evaluates the same formulas on the GPU for large grids."""
    where = staticmethod(np.where)
    sin = staticmethod(np.sin)
    cos = staticmethod(np.cos)
    tan = staticmethod(np.tan)
    log = staticmethod(np.log)
    exp = staticmethod(np.exp)
    sqrt = staticmethod(np.sqrt)
    abs = staticmethod(np.abs)
    arccos = staticmethod(np.arccos)
    zeros_like = staticmethod(np.zeros_like)

    @staticmethod
    def clip(a, lo, hi):
        return np.clip(a, lo, hi)

    @staticmethod
    def full_like_broadcast(a, b, value):
        return np.full(np.broadcast(a, b).shape, value)

    @staticmethod
    def broadcast_to(a, shape):
        return np.broadcast_to(a, shape)

    @staticmethod
    def all(a):
        return bool(np.all(a))

    @staticmethod
    def zeros_bool(shape):
        return np.zeros(shape, dtype=bool)


class _Torch:
    def __init__(self):
        import torch
        self.t = torch
        for name in ("where", "sin", "cos", "tan", "log", "exp", "sqrt", "abs",
                     "arccos", "zeros_like", "broadcast_to"):
            setattr(self, name, getattr(torch, name))

    def clip(self, a, lo, hi):
        return self.t.clamp(a, lo, hi)

    def full_like_broadcast(self, a, b, value):
        shape = self.t.broadcast_shapes(a.shape, b.shape)
        return self.t.full(shape, value, dtype=self.t.float64, device=a.device)

    def all(self, a):
        return bool(a.all().item())

    def zeros_bool(self, shape):
        return self.t.zeros(shape, dtype=self.t.bool, device="cuda")


NUMPY = _Numpy()


class ShallowWaterTestCase2:
    """Williamson et al. (1992) test 2: steady geostrophic flow."""

    equation_set = "shallow_water"
    ztop = 1.0

    def __init__(self, h0=2998.104995, u0=38.61068277, alpha=0.0):
        self.h0, self.u0, self.alpha = h0, u0, alpha

    def evaluate_topography(self, phys, lon, lat):
        return np.zeros_like(lon)

    def evaluate_pointwise_state(self, phys, z, lon, lat, xp=NUMPY):
        lat = xp.where(xp.abs(lat - 0.5 * math.pi) < 1.0e-12, lat - 1.0e-12, lat)
        lat = xp.where(xp.abs(lat + 0.5 * math.pi) < 1.0e-12, lat + 1.0e-12, lat)
        a = self.alpha
        u = self.u0 * xp.cos(lat) * (math.cos(a) + xp.cos(lon) * xp.tan(lat) * math.sin(a))
        v = -self.u0 * xp.sin(lon) * math.sin(a) + 0.0 * lat
        htrig = -xp.cos(lon) * xp.cos(lat) * math.sin(a) + xp.sin(lat) * math.cos(a)
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

    def _profiles(self, phys, aux_eta, lat, xp=NUMPY):
        s = xp.sin(lat)
        s2 = s * s
        s6 = s2 * s2 * s2
        c = xp.cos(lat)
        c2 = c * c
        c3 = c2 * c
        p1 = self.u0 * xp.cos(aux_eta) ** 1.5 * (-2.0 * s6 * (c2 + 1.0 / 3.0) + 10.0 / 63.0)
        p2 = phys.earth_radius * phys.omega * (8.0 / 5.0 * c3 * (s2 + 2.0 / 3.0) - 0.25 * math.pi)
        return p1, p2

    def evaluate_topography(self, phys, lon, lat):
        aux = 0.5 * math.pi * (1.0 - self.eta0)
        p1, p2 = self._profiles(phys, np.float64(aux), lat)
        return self.u0 * math.cos(aux) ** 1.5 * (p1 + p2) / phys.g

    def geopotential_temperature(self, phys, eta, lat, xp=NUMPY):
        aux = 0.5 * math.pi * (eta - self.eta0)
        expo = phys.R * self.lapse_rate / phys.g
        tavg = self.t0 * eta ** expo
        strat = eta < self.tropopause_eta
        te = self.tropopause_eta
        zero = xp.zeros_like(eta)
        tavg = tavg + xp.where(strat, self.delta_t * xp.clip(te - eta, 0.0, 1.0e30) ** 5.0, zero)
        p1, p2 = self._profiles(phys, aux, lat, xp)
        temp = 2.0 * p1 + p2
        temp = tavg + 0.75 * eta * math.pi * self.u0 / phys.R \
            * xp.sin(aux) * xp.sqrt(xp.cos(aux)) * temp
        gavg = self.t0 * phys.g / self.lapse_rate * (1.0 - eta ** expo)
        es = xp.where(strat, eta, zero + te)
        corr = phys.R * self.delta_t * (
            (xp.log(es / te) + 137.0 / 60.0) * te ** 5
            - 5.0 * te ** 4 * es + 5.0 * te ** 3 * es ** 2
            - (10.0 / 3.0) * te ** 2 * es ** 3 + 5.0 / 4.0 * te * es ** 4
            - 1.0 / 5.0 * es ** 5)
        gavg = gavg - xp.where(strat, corr, zero)
        geo = gavg + self.u0 * xp.cos(aux) ** 1.5 * (p1 + p2)
        return geo, temp

    def eta_from_rll(self, phys, z, lat, xp=NUMPY):
        """Newton iteration of EtaFromRLL (:297-345)."""
        eta = xp.full_like_broadcast(z, lat, 1.0e-7)
        lat = xp.broadcast_to(lat, eta.shape)
        z = xp.broadcast_to(z, eta.shape)
        done = xp.zeros_bool(eta.shape)
        geo = temp = None
        for _ in range(25):
            geo, temp = self.geopotential_temperature(phys, eta, lat, xp)
            f = -phys.g * z + geo
            df = -phys.R / eta * temp
            new = eta - f / df
            conv = xp.abs(eta - new) < 1.0e-13
            eta = xp.where(done, eta, new)
            done = done | conv
            if xp.all(done):
                break
        if not xp.all(done):
            raise RuntimeError("Maximum number of iterations exceeded.")
        # the reference returns geopotential / temperature of the last
        # evaluated iterate
        return eta, temp

    def evaluate_reference_state(self, phys, z, lon, lat, xp=NUMPY):
        eta, temp = self.eta_from_rll(phys, z, lat, xp)
        ulon = self.u0 * xp.cos(0.5 * math.pi * (eta - self.eta0)) ** 1.5 \
            * xp.sin(2.0 * lat) * xp.sin(2.0 * lat)
        p = phys.p0 * eta
        rho = p / (phys.R * temp)
        # PhysicalConstants::RhoThetaFromPressure (PhysicalConstants.h:389-391)
        rhotheta = xp.exp(xp.log(p / phys.pressure_scaling) / phys.gamma)
        zero = xp.zeros_like(ulon)
        return [ulon, zero, rhotheta / rho, zero, rho]

    def evaluate_pointwise_state(self, phys, z, lon, lat, xp=NUMPY):
        st = self.evaluate_reference_state(phys, z, lon, lat, xp)
        if self.perturbation == "exp":
            lonb = xp.broadcast_to(lon, st[0].shape)
            latb = xp.broadcast_to(lat, st[0].shape)
            r = xp.arccos(xp.clip(
                math.sin(self.pert_lat) * xp.sin(latb)
                + math.cos(self.pert_lat) * xp.cos(latb) * xp.cos(lonb - self.pert_lon),
                -1.0, 1.0))
            r = r / self.pert_r
            st[0] = st[0] + xp.where(r < 1.0, self.up * xp.exp(-r * r), xp.zeros_like(r))
        return st


class ThermalBubbleCartesianTest:
    """Rising thermal bubble on an x-z slice
    (reference test/nonhydro_xz/ThermalBubbleCartesianTest.cpp:75-285)."""

    equation_set = "primitive_nonhydro"
    # x0, x1, y0, y1, z0, z1 (:93-99)
    dims = (0.0, 1000.0, -500.0, 500.0, 0.0, 1000.0)
    ztop = 1000.0

    def __init__(self, theta_bar=300.0, theta_c=0.5, r_c=250.0, x_c=500.0, z_c=350.0,
                 pi_c=3.14159265):
        self.theta_bar, self.theta_c = theta_bar, theta_c
        self.r_c, self.x_c, self.z_c, self.pi_c = r_c, x_c, z_c, pi_c

    def evaluate_topography(self, phys, x, y):
        return np.zeros_like(x)

    def evaluate_pointwise_state(self, phys, z, x, y, xp=NUMPY):
        """-> [u, v, theta, w, rho] (:236-268); u, v are the covariant components
        themselves on the Cartesian grid."""
        rp = xp.sqrt((x - self.x_c) * (x - self.x_c) + (z - self.z_c) * (z - self.z_c))
        tprime = xp.where(rp <= self.r_c,
                          0.5 * self.theta_c * (1.0 + xp.cos(self.pi_c * rp / self.r_c)),
                          0.0 * rp)
        theta = self.theta_bar + tprime
        exner = -phys.g / (phys.cp * self.theta_bar) * z + 1.0
        rho = phys.p0 / (phys.R * self.theta_bar) * exner ** (phys.cv / phys.R) + 0.0 * x
        zero = 0.0 * rho
        return [zero, zero, theta, zero, rho]


class BaroclinicWaveJWTracerTest(BaroclinicWaveJWTest):
    """The JW wave carrying `ntracers` analytic tracer densities rho * q:
    q_0 a smooth global field decaying with height, q_c (c >= 1) cosine bells.
    The same closed forms as the oracle's dump hook uses for its tracer
    fixtures (oracle/ref_dump.cpp, JWTracerTest), i.e. a dry stand-in for the
    five-tracer configuration 4 of SURVEY 8 (the DCMIP-2016 driver itself
    needs gfortran)."""

    def __init__(self, ntracers=3, **kw):
        super().__init__(**kw)
        self.ntracers = int(ntracers)

    def evaluate_tracers(self, phys, z, lon, lat, rho, xp=NUMPY):
        out = []
        for c in range(self.ntracers):
            if c == 0:
                q = 1.0e-3 * (2.0 + xp.sin(lon) * xp.cos(lat)) * xp.exp(-z / 8000.0)
            else:
                lon0 = 0.5 + 1.7 * c
                lat0 = 0.6 - 0.5 * c
                z0 = 4000.0 + 3000.0 * c
                r = xp.arccos(math.sin(lat0) * xp.sin(lat)
                              + math.cos(lat0) * xp.cos(lat) * xp.cos(lon - lon0))
                rz = xp.abs(z - z0) / 6000.0
                dd = xp.sqrt(r * r / (0.9 * 0.9) + rz * rz)
                q = xp.where(dd < 1.0, 0.5e-2 * (1.0 + xp.cos(math.pi * dd)), 0.0 * dd)
            out.append(rho * q)
        return out


class BaroclinicWaveJWMoistTest(BaroclinicWaveJWTracerTest):
    """The JW wave with moisture: tracers 0, 1, 2 are rho qv, rho qc, rho qr (the
    Kessler categories), the rest the analytic tracers of the stand-in case.  The
    specific humidity is the DCMIP-2016 profile q0 exp(-(lat / latw)^4)
    exp(-((eta - 1) p0 / pw)^2) with eta taken as the isothermal-atmosphere value
    exp(-z / H): test data for the moist configuration 4, not a reference test."""

    def __init__(self, ntracers=5, q0=0.018, **kw):
        super().__init__(ntracers=ntracers, **kw)
        if self.ntracers < 3:
            raise ValueError("the moist case carries at least rho qv, rho qc, rho qr")
        self.q0 = q0

    def evaluate_tracers(self, phys, z, lon, lat, rho, xp=NUMPY):
        rest = super().evaluate_tracers(phys, z, lon, lat, rho, xp)
        eta = xp.exp(-z / 8000.0)
        latw = 2.0 * math.pi / 9.0
        pw = 34000.0
        q = self.q0 * xp.exp(-(lat / latw) ** 4) * xp.exp(-((eta - 1.0) * phys.p0 / pw) ** 2)
        zero = 0.0 * q
        return [rho * q, rho * zero, rho * zero] + rest[3:]
