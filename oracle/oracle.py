"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/reentry_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.  PARITY PINNED against golden
vectors made from the unmodified reference (tests/golden/, oracle/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libreentry_oracle.so")

_lib = None

c_dp = C.POINTER(C.c_double)


class RhsParts(C.Structure):
    _fields_ = [
        ("a_total", C.c_double * 3), ("a_grav", C.c_double * 3), ("a_j2", C.c_double * 3),
        ("a_moon", C.c_double * 3), ("a_sun", C.c_double * 3), ("a_drag", C.c_double * 3),
        ("altitude", C.c_double), ("latitude_deg", C.c_double), ("rho", C.c_double), ("T", C.c_double),
        ("v_rel", C.c_double * 3), ("geodetic_iters", C.c_int),
    ]


FMA_LIB_PATH = os.path.join(HERE, "libreentry_oracle_fma.so")
_lib_fma = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "reentry_oracle.c")
    hdr = os.path.join(os.path.dirname(HERE), "include", "reentry_constants.h")
    for path in (LIB_PATH, FMA_LIB_PATH):
        stale = (not os.path.exists(path)) or any(os.path.getmtime(p) > os.path.getmtime(path) for p in (src, hdr))
        if force or stale:
            subprocess.run(["make", "-C", HERE, "-B", os.path.basename(path)], check=True, capture_output=True)
    return LIB_PATH


def lib_fma():
    """The same C source compiled with FMA contraction: an equally valid second rounding of the algorithm (see Makefile).
    Only rbo_propagate is used from it."""
    global _lib_fma
    if _lib_fma is None:
        if not os.path.exists(FMA_LIB_PATH):
            build()
        _lib_fma = C.CDLL(FMA_LIB_PATH)
        _lib_fma.rbo_propagate.restype = C.c_int
    return _lib_fma


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.rbo_euclidean_norm.restype = C.c_double
        _lib.rbo_solar_activity_factor.restype = C.c_double
        _lib.rbo_solar_factor_time.restype = C.c_double
        _lib.rbo_propagate.restype = C.c_int
        _lib.rbo_max_threads.restype = C.c_int
        _lib.rbo_haversine_distance.restype = C.c_double
        _lib.rbo_ground_track.restype = C.c_int
        _lib.rbo_propagate_adaptive.restype = C.c_int
    return _lib


def _v3(a):
    return (C.c_double * 3)(*[float(x) for x in a])


def _out3():
    return (C.c_double * 3)()


def euclidean_norm(v):
    return lib().rbo_euclidean_norm(_v3(v))


def eci_to_ecef(v, gmst):
    o = _out3(); lib().rbo_eci_to_ecef(_v3(v), C.c_double(gmst), o); return np.array(o)


def ecef_to_eci(v, gmst):
    o = _out3(); lib().rbo_ecef_to_eci(_v3(v), C.c_double(gmst), o); return np.array(o)


def geodetic_to_spheroid(lat, lon, alt):
    o = _out3(); lib().rbo_geodetic_to_spheroid(C.c_double(lat), C.c_double(lon), C.c_double(alt), o)
    return tuple(o)


def ecef_to_geodetic(x, y, z, return_iters=False):
    o = _out3(); it = C.c_int(0)
    lib().rbo_ecef_to_geodetic(C.c_double(x), C.c_double(y), C.c_double(z), o, C.byref(it))
    return (tuple(o), it.value) if return_iters else tuple(o)


def enu_to_ecef(v, lat, lon):
    o = _out3(); lib().rbo_enu_to_ecef(_v3(v), C.c_double(lat), C.c_double(lon), o); return np.array(o)


def solar_activity_factor(jd, altitude):
    return lib().rbo_solar_activity_factor(C.c_double(jd), C.c_double(altitude))


def atmosphere_model(altitude, latitude, jd):
    rho = C.c_double(); T = C.c_double()
    lib().rbo_atmosphere_model(C.c_double(altitude), C.c_double(latitude), C.c_double(jd), C.byref(rho), C.byref(T))
    return rho.value, T.value


def atmospheric_drag(Cd, A, rho, v, mass):
    o = _out3()
    lib().rbo_atmospheric_drag(C.c_double(Cd), C.c_double(A), C.c_double(rho), _v3(v), C.c_double(mass), o)
    return np.array(o)


def moon_position_vector(jd):
    o = _out3(); lib().rbo_moon_position_vector(C.c_double(jd), o); return np.array(o)


def sun_position_vector(jd):
    o = _out3(); lib().rbo_sun_position_vector(C.c_double(jd), o); return np.array(o)


def third_body_acceleration(sat, body, k):
    o = _out3(); lib().rbo_third_body_acceleration(_v3(sat), _v3(body), C.c_double(k), o); return np.array(o)


def j2_perturbation(r, k, J2, R):
    o = _out3(); lib().rbo_j2_perturbation(_v3(r), C.c_double(k), C.c_double(J2), C.c_double(R), o)
    return np.array(o)


def get_initial_state(v, lat, lon, alt, azimuth, gamma, gmst=0.0):
    o = (C.c_double * 6)()
    lib().rbo_get_initial_state(*[C.c_double(float(x)) for x in (v, lat, lon, alt, azimuth, gamma, gmst)], o)
    return np.array(o)


def get_initial_states(v, lat, lon, alt, azimuth, gamma, gmst=0.0):
    args = np.broadcast_arrays(*[np.asarray(a, dtype=np.float64) for a in (v, lat, lon, alt, azimuth, gamma)])
    out = np.empty((args[0].size, 6))
    for i, row in enumerate(zip(*[a.ravel() for a in args])):
        out[i] = get_initial_state(*row, gmst)
    return out


def equations_of_motion(t, y, epoch_jd, gmst0, Cd, A, m):
    """Returns (dy[6], parts dict) -- the dict of spacecraft_model.py:472-480 minus thermal entries."""
    yy = (C.c_double * 6)(*[float(x) for x in y]); dy = (C.c_double * 6)(); p = RhsParts()
    lib().rbo_equations_of_motion(C.c_double(t), yy, C.c_double(epoch_jd), C.c_double(gmst0), C.c_double(Cd),
                                  C.c_double(A), C.c_double(m), dy, C.byref(p))
    parts = dict(
        velocity=np.array(y[3:6], dtype=np.float64), acceleration=np.array(p.a_total),
        gravitational_acceleration=np.array(p.a_grav), J2_acceleration=np.array(p.a_j2),
        moon_acceleration=np.array(p.a_moon), sun_acceleration=np.array(p.a_sun),
        drag_acceleration=np.array(p.a_drag), altitude=p.altitude, latitude_deg=p.latitude_deg,
        rho=p.rho, T=p.T, v_rel=np.array(p.v_rel), geodetic_iters=p.geodetic_iters)
    return np.array(dy), parts


def time_table_entry(t, epoch_jd, gmst0):
    o = (C.c_double * 16)()
    lib().rbo_time_table_entry(C.c_double(t), C.c_double(epoch_jd), C.c_double(gmst0), o)
    return np.array(o)


def tableau():
    c = (C.c_double * 6)(); a = (C.c_double * 30)(); b = (C.c_double * 6)(); p = (C.c_double * 28)()
    lib().rbo_tableau(c, a, b, p)
    return np.array(c), np.array(a).reshape(6, 5), np.array(b), np.array(p).reshape(7, 4)


def max_threads() -> int:
    return lib().rbo_max_threads()


def propagate(y0, Cd, A, m, epoch_jd, gmst0, t0, dt, n_steps, stride=1, want_samples=True, n_threads=0, fma=False):
    """Fixed-step DP5 ensemble on the CPU.  y0 [N,6].  fma=True: the FMA-contracted build (second rounding).  Returns dict:
    samples [N, n_out, 8] (x,y,z,vx,vy,vz,altitude,speed; NaN after the event) or None,
    impact_step [N] i32 (-1 none), impact_time [N], final_state [N,6], status [N] u8,
    n_samples [N] i64, threads (int)."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1, 6)
    N = y0.shape[0]
    bc = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (N,)))
    Cd, A, m = bc(Cd), bc(A), bc(m)
    n_out = int(n_steps) // int(stride) + 1
    samples = np.full((N, n_out, 8), np.nan) if want_samples else None
    impact_step = np.empty(N, dtype=np.int32)
    impact_time = np.empty(N)
    final_state = np.empty((N, 6))
    status = np.empty(N, dtype=np.uint8)
    n_samples = np.empty(N, dtype=np.int64)
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    used = (lib_fma() if fma else lib()).rbo_propagate(
        C.c_int64(N), P(y0), P(Cd), P(A), P(m), C.c_double(epoch_jd), C.c_double(gmst0), C.c_double(t0),
        C.c_double(dt), C.c_int64(n_steps), C.c_int64(stride), P(samples), P(impact_step), P(impact_time),
        P(final_state), P(status), P(n_samples), C.c_int(n_threads))
    return dict(samples=samples, impact_step=impact_step, impact_time=impact_time, final_state=final_state,
                status=status, n_samples=n_samples, threads=used)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().rbo_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def dispersed_params(n, seed, nominal, sigma, first_index=0):
    """N3 generator spec (counter-based): params [n, 6] = nominal + sigma * z."""
    out = np.empty((n, 6))
    lib().rbo_dispersed_params(C.c_int64(n), C.c_int64(first_index), C.c_uint64(seed), (C.c_double * 6)(*nominal),
                               (C.c_double * 6)(*sigma), out.ctypes.data_as(C.c_void_p))
    return out


def haversine_distance(lat1, lon1, lat2, lon2):
    return lib().rbo_haversine_distance(*[C.c_double(float(x)) for x in (lat1, lon1, lat2, lon2)])


def ground_track(r_eci, t, altitude, gmst0, threshold=100000.0):
    """app.py:225-243, :302-309 for one trajectory: r_eci [n,3].  Returns geo [n,3] (lat deg, lon deg, h),
    downrange [n], crossing indices."""
    r = np.ascontiguousarray(r_eci, dtype=np.float64); n = r.shape[0]
    t = np.ascontiguousarray(t, dtype=np.float64); alt = np.ascontiguousarray(altitude, dtype=np.float64)
    geo = np.empty((n, 3)); dr = np.empty(n); ci = np.empty(64, dtype=np.int64)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    nc = lib().rbo_ground_track(C.c_int64(n), P(r), P(t), P(alt), C.c_double(gmst0), C.c_double(threshold), P(geo), P(dr),
                                P(ci), C.c_int(64))
    return geo, dr, ci[:min(nc, 64)].copy()


def spacecraft_temperature(v_norm, atmo_T, a_drag_norm, capsule_length, dt, thermal_conductivity, specific_heat_capacity,
                           emissivity, ablation_efficiency, iter_fact=2, spacecraft_m=500):
    """spacecraft_model.py:243-255; returns (Qc, Qr, Q_net, Q, T_s, dT) in the reference's return order."""
    o = (C.c_double * 6)()
    lib().rbo_spacecraft_temperature(*[C.c_double(float(x)) for x in (
        v_norm, atmo_T, a_drag_norm, capsule_length, dt, thermal_conductivity, specific_heat_capacity, emissivity,
        ablation_efficiency, iter_fact, spacecraft_m)], o)
    return tuple(o)


def propagate_adaptive(y0, Cd, A, m, epoch_jd, gmst0, t_span, t_eval0, t_eval_dt, n_eval, rtol=1e-8, atol=1e-10,
                       max_steps=100000, method="RK45"):
    """The reference's run_simulation as shipped (scipy adaptive RK45, spacecraft_model.py:516) for ONE trajectory.
    Returns dict: t [n], y [6,n], altitude [n], speed [n], t_event (NaN if none), y_final [6], status, n_accepted,
    n_rejected, nfev, step_t (accepted step end times)."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    samples = np.full((int(n_eval), 8), np.nan)
    ns = C.c_int64(); te = C.c_double(); ye = (C.c_double * 6)(); na = C.c_int64(); nr = C.c_int64(); nf = C.c_int64()
    step_t = np.full(max_steps + 1, np.nan)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().rbo_propagate_adaptive_method.restype = C.c_int
    st = lib().rbo_propagate_adaptive_method(C.c_int({"RK45": 0, "RK23": 1, "DOP853": 2}[method]), P(y0), C.c_double(Cd), C.c_double(A), C.c_double(m), C.c_double(epoch_jd),
                                      C.c_double(gmst0), C.c_double(t_span[0]), C.c_double(t_span[1]), C.c_double(rtol),
                                      C.c_double(atol), C.c_double(t_eval0), C.c_double(t_eval_dt), C.c_int64(n_eval), P(samples),
                                      C.byref(ns), C.byref(te), ye, C.byref(na), C.byref(nr), C.byref(nf), P(step_t),
                                      C.c_int64(max_steps))
    n = ns.value
    return dict(t=t_eval0 + t_eval_dt * np.arange(n), y=samples[:n, :6].T.copy(), altitude=samples[:n, 6].copy(),
                speed=samples[:n, 7].copy(), t_event=te.value, y_final=np.array(ye), status=st, n_accepted=na.value,
                n_rejected=nr.value, nfev=nf.value, step_t=step_t[: na.value + 1].copy())
