"""TEST INFRASTRUCTURE ONLY -- import shim for the UNMODIFIED reference (build container only).

Imports /root/reference/spacecraft_model.py as it lies on disk, after registering
empty stub modules for the UI-only third-party packages it names at import time
(astropy, poliastro, matplotlib -- none of which the hot path touches; see
SURVEY.md section 8(c)).  Nothing here is shipped or used by the product path, and
nothing here runs on the GPU box (there is no /root/reference there): it exists so
that `oracle/make_golden.py` can generate golden vectors from the real reference
and so that the C restatement (oracle/reentry_oracle.c) can be pinned against it.

Two fixed-step constructions of the reference path are provided:

* O2 -- scipy's own RK45 stepper pinned to a constant step through its public
        options, RHS = the unmodified `SpacecraftModel.equations_of_motion`
        (spacecraft_model.py:437-493, call at :516).
* O3 -- an @njit fixed-step Dormand-Prince driver (scipy tableau, rk.py:541-550)
        whose RHS calls the reference's imported, unmodified njit leaves; only
        the glue of spacecraft_model.py:437-470 and :181-188 is restated.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("REENTRY_REFERENCE_DIR", "/root/reference")

_loaded = None


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


class _Time:
    """Minimal stand-in for astropy.time.Time: only `.jd` is read (spacecraft_model.py:398)."""

    def __init__(self, value="2024-01-01 00:00:00", *a, jd=None, **k):
        if jd is not None:
            self.jd = float(jd)
        elif isinstance(value, (int, float)):
            self.jd = float(value)
        elif str(value).startswith("2024-01-01 00:00:00"):
            self.jd = 2460310.5
        else:
            raise ValueError("stub Time only knows the reference's default epoch")


class Epoch:
    """Any object with .jd is a valid `epoch` for the reference constructor."""

    def __init__(self, jd):
        self.jd = float(jd)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "spacecraft_model.py"))


def load():
    """Return (spacecraft_model, coordinate_converter, constants) modules of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_DIR}")
    if "astropy" not in sys.modules:
        a = _stub("astropy")
        a.units = _stub("astropy.units")
        a.time = _stub("astropy.time", Time=_Time)
    if "poliastro" not in sys.modules:
        p = _stub("poliastro")
        p.core = _stub("poliastro.core")
        p.core.elements = _stub("poliastro.core.elements", coe2rv=None)
        p.twobody = _stub("poliastro.twobody", Orbit=type("Orbit", (), {}))
    if "matplotlib" not in sys.modules:
        m = _stub("matplotlib")
        m.colors = _stub("matplotlib.colors")
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # the module prints a curve_fit result at import
        import spacecraft_model as sm  # noqa
        import coordinate_converter as cc  # noqa
        import constants as const  # noqa
    _loaded = (sm, cc, const)
    return _loaded


def make_model(Cd, A, m, epoch_jd, gmst0):
    """Reference model; dt=10, iter_fact=2 avoids the thermal-loop crash."""
    sm, _, _ = load()
    return sm.SpacecraftModel(Cd=Cd, A=A, m=m, epoch=Epoch(epoch_jd), gmst0=gmst0, dt=10, iter_fact=2)


# --------------------------------------------------------------------------- O2
def o2_scipy_pinned(model, y0, t0, dt, n_steps):
    """scipy RK45 with step-size control disabled; RHS = unmodified equations_of_motion.

    Returns dict(t, y[6,n], t_event or None, y_event or None, n_fev, uniform).
    """
    from scipy.integrate import solve_ivp

    sm, _, _ = load()

    def rhs(t, y):  # spacecraft_model.py:491-493
        dy = model.equations_of_motion(t, y)
        return np.concatenate((dy["velocity"], dy["acceleration"]))

    def altitude_event(t, y):  # spacecraft_model.py:496-501
        return sm.euclidean_norm(y[0:3]) - sm.EARTH_R - 1000.0

    altitude_event.terminal = True
    altitude_event.direction = -1
    tf = t0 + n_steps * dt
    sol = solve_ivp(rhs, (t0, tf), np.asarray(y0, dtype=np.float64), method="RK45",
                    first_step=dt, max_step=dt, rtol=1e3, atol=1e30,
                    events=[altitude_event], dense_output=False,
                    t_eval=t0 + dt * np.arange(n_steps + 1))
    te = sol.t_events[0]
    return dict(t=sol.t, y=sol.y, t_event=(float(te[0]) if len(te) else None),
                y_event=(sol.y_events[0][0] if len(te) else None), n_fev=sol.nfev)


# --------------------------------------------------------------------------- O3
_o3 = None


def o3_build():
    """Build the njit fixed-step DP5 driver over the reference's own njit leaves."""
    global _o3
    if _o3 is not None:
        return _o3
    import numba
    from numba import njit

    sm, cc, const = load()
    euclidean_norm = sm.euclidean_norm
    eci_to_ecef = cc.eci_to_ecef
    ecef_to_eci = cc.ecef_to_eci
    ecef_to_geodetic = cc.ecef_to_geodetic
    J2 = sm.J2_perturbation_numba
    moon_pos = sm.moon_position_vector
    sun_pos = sm.sun_position_vector
    third = sm.third_body_acceleration
    nrl = sm.simplified_nrlmsise_00
    drag = sm.atmospheric_drag
    EARTH_OMEGA = const.EARTH_OMEGA
    EARTH_MU = const.EARTH_MU
    EARTH_J2 = const.EARTH_J2
    EARTH_R = const.EARTH_R
    MOON_K = const.MOON_K
    SUN_K = const.SUN_K

    @njit
    def accel(t, y, epoch0, gmst0, Cd, A, m):
        # glue of spacecraft_model.py:437-470 (thermal call at :468 omitted: not part of dy)
        r_eci = y[0:3].copy()
        v_eci = y[3:6].copy()
        r_norm = euclidean_norm(r_eci)
        epoch = epoch0 + t / 86400.0
        gmst = gmst0 + EARTH_OMEGA * t
        r_ecef = eci_to_ecef(r_eci, gmst)
        v_ground = eci_to_ecef(v_eci, gmst)
        v_rel = v_ground - np.array([-EARTH_OMEGA * r_ecef[1], EARTH_OMEGA * r_ecef[0], 0.0])
        a_grav = -EARTH_MU * r_eci / (r_norm ** 3)
        a_J2 = J2(r_eci, EARTH_MU, EARTH_J2, EARTH_R)
        moon_r = moon_pos(epoch)
        sun_r = sun_pos(epoch)
        a_moon = third(r_eci, moon_r, MOON_K)
        a_sun = third(r_eci, sun_r, SUN_K)
        altitude = r_norm - EARTH_R
        latitude, _, _ = ecef_to_geodetic(r_ecef[0], r_ecef[1], r_ecef[2], 100, 1e-6)
        if altitude <= 0:  # atmosphere_model wrapper, spacecraft_model.py:181-188
            rho = 1.225
        else:
            rho, _T = nrl(altitude, latitude, epoch)
        a_drag_ecef = drag(Cd, A, rho, v_rel, m)
        a_drag = ecef_to_eci(a_drag_ecef, gmst)
        a_total = a_grav + a_J2 + a_moon + a_sun + a_drag
        out = np.empty(6)
        out[0:3] = v_eci
        out[3:6] = a_total
        return out

    C = np.array([0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0])
    Amat = np.zeros((6, 5))
    Amat[1, :1] = [1 / 5]
    Amat[2, :2] = [3 / 40, 9 / 40]
    Amat[3, :3] = [44 / 45, -56 / 15, 32 / 9]
    Amat[4, :4] = [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729]
    Amat[5, :5] = [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]
    B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])

    @njit
    def propagate_one(y0, t0, dt, n_steps, stride, epoch0, gmst0, Cd, A, m, ys, K_event):
        """Fixed-step DP5 (FSAL).  ys[k] <- y at step k*stride.  Returns (impact_step, n_done).

        Event rule of scipy/ivp.py:151 for direction -1: g_old >= 0 and g_new <= 0.
        On impact, K_event[0:7] holds K0..K6 of the crossing step and K_event[7] its y_old.
        """
        y = y0.copy()
        t = t0
        K = np.empty((7, 6))
        K[0] = accel(t, y, epoch0, gmst0, Cd, A, m)
        g_old = euclidean_norm(y[0:3]) - EARTH_R - 1000.0
        ys[0] = y
        for n in range(n_steps):
            for s in range(1, 6):
                dy = np.zeros(6)
                for j in range(s):
                    dy += Amat[s, j] * K[j]
                K[s] = accel(t + C[s] * dt, y + dy * dt, epoch0, gmst0, Cd, A, m)
            acc = np.zeros(6)
            for j in range(6):
                acc += B[j] * K[j]
            y_new = y + dt * acc
            t_new = t + dt
            K[6] = accel(t_new, y_new, epoch0, gmst0, Cd, A, m)
            g_new = euclidean_norm(y_new[0:3]) - EARTH_R - 1000.0
            if g_old >= 0.0 and g_new <= 0.0:
                K_event[0:7] = K
                K_event[7] = y
                return n + 1, n
            y = y_new
            t = t_new
            g_old = g_new
            K[0] = K[6]
            if (n + 1) % stride == 0:
                ys[(n + 1) // stride] = y
        return -1, n_steps

    @njit(parallel=True)
    def propagate_many(y0, t0, dt, n_steps, stride, epoch0, gmst0, Cd, A, m, ys, impact_step, n_done, K_event):
        for i in numba.prange(y0.shape[0]):
            a, b = propagate_one(y0[i], t0, dt, n_steps, stride, epoch0, gmst0, Cd[i], A[i], m[i], ys[i], K_event[i])
            impact_step[i] = a
            n_done[i] = b

    _o3 = dict(accel=accel, propagate_one=propagate_one, propagate_many=propagate_many)
    return _o3


# scipy RK45 dense-output matrix (scipy/integrate/_ivp/rk.py:554-565), read from scipy itself
def dense_P():
    from scipy.integrate._ivp.rk import RK45

    return np.array(RK45.P, dtype=np.float64)


def event_time_from_K(K_event, t_old, dt):
    """scipy's event refinement (ivp.py:52-77): brentq on the dense-output interpolant."""
    from scipy.optimize import brentq

    _, _, const = load()
    EPS = np.finfo(float).eps
    K = K_event[0:7]
    y_old = K_event[7]
    Q = K.T.dot(dense_P())  # rk.py:173

    def sol(t):  # rk.py:723-737
        x = (t - t_old) / dt
        p = np.cumprod(np.tile(x, 4))
        return y_old + dt * np.dot(Q, p)

    def g(t):
        yy = sol(t)
        return np.sqrt(yy[0] ** 2 + yy[1] ** 2 + yy[2] ** 2) - const.EARTH_R - 1000.0

    te = brentq(g, t_old, t_old + dt, xtol=4 * EPS, rtol=4 * EPS)
    return te, sol(te)


def o3_run(y0, Cd, A, m, epoch_jd, gmst0, t0, dt, n_steps, stride=1):
    """Run O3 on y0[N,6]; returns dict with ys[N,n_out,6] (NaN after impact), impact_step, t_event, y_event."""
    o3 = o3_build()
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1, 6)
    N = y0.shape[0]
    Cd = np.broadcast_to(np.asarray(Cd, dtype=np.float64), (N,)).copy()
    A = np.broadcast_to(np.asarray(A, dtype=np.float64), (N,)).copy()
    m = np.broadcast_to(np.asarray(m, dtype=np.float64), (N,)).copy()
    n_out = n_steps // stride + 1
    ys = np.full((N, n_out, 6), np.nan)
    impact_step = np.zeros(N, dtype=np.int64)
    n_done = np.zeros(N, dtype=np.int64)
    K_event = np.zeros((N, 8, 6))
    o3["propagate_many"](y0, float(t0), float(dt), int(n_steps), int(stride), float(epoch_jd), float(gmst0),
                         Cd, A, m, ys, impact_step, n_done, K_event)
    t_event = np.full(N, np.nan)
    y_event = np.full((N, 6), np.nan)
    for i in range(N):
        if impact_step[i] > 0:
            t_old = t0 + dt * (impact_step[i] - 1)
            t_event[i], y_event[i] = event_time_from_K(K_event[i], t_old, dt)
    return dict(ys=ys, impact_step=impact_step, n_done=n_done, t_event=t_event, y_event=y_event)
