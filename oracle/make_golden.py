"""TEST INFRASTRUCTURE -- generate tests/golden/* from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; see oracle/ref_shim.py).
Everything written here is an OUTPUT OF THE REFERENCE'S OWN CODE:

  leaves.json          known-answer values of every hot-path leaf at the inputs of the
                       reference's commented-out smoke snippets
  leaves_random.npz    the same leaves on seeded random inputs
  rhs_random.npz       SpacecraftModel.equations_of_motion (unmodified, :437-487) on random states
  c1_single.npz        config C1 (single capsule, dt = 1 s): O2 = scipy RK45 pinned to a fixed
                       step with the unmodified Python RHS, O3 = njit DP5 over the reference's
                       njit leaves, and O1 = run_simulation as shipped (adaptive; loose check)
  c2_mc.npz            48 Monte-Carlo reentry dispersions (dt = 0.5 s, stride 16), O3 (+O2 on 2)
  c3_sso.npz           8 sun-synchronous 550 km satellites, dt = 10 s, 1 day, stride 360, O3
  c4_geo_heo.npz       4 GEO + 4 HEO, dt = 60 s, 6 days, stride 1440, O3
  c3_sso_7d.npz        64 sun-synchronous 550 km satellites, dt = 10 s, 7 days, stride 360, O3
  c4_geo_heo_30d.npz   32 GEO + 32 HEO, dt = 60 s, 30 days, stride 1440, O3
  edge.npz             edge cases: start below the event altitude, steep entry, per-trajectory
                       Cd/A/m, event in the first step, no event within the horizon
  ground_track.npz     N2: per-sample geodetic coordinates, cumulative haversine down-range and
                       100 km crossings of four O3 trajectories, from the reference's own functions
  thermal.npz          N1: thermal entries of the unmodified equations_of_motion dict + direct
                       spacecraft_temperature calls
  adaptive.npz         N4: run_simulation as shipped (scipy adaptive RK45) on four launch geometries
  adaptive_rk23.npz    the same with sim_type='RK23' (scipy's Bogacki-Shampine 3(2) pair)
  adaptive_dop853.npz  the same with sim_type='DOP853', plus the reference's sensitivity to a 1e-14 change of y0
  recorded_trajectory.json   first samples of the trajectory embedded in temperature_model.py:104-105

    python oracle/make_golden.py        # ~2-3 min (numba JIT of the reference leaves dominates)
"""
from __future__ import annotations

import json
import os
import re
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
EPOCH_JD = 2460310.5  # Time('2024-01-01 00:00:00').jd, spacecraft_model.py:393
JD_STAR = EPOCH_JD + 1 / 86400


def fl(x):
    return [float(v) for v in np.atleast_1d(np.asarray(x, dtype=np.float64)).ravel()]


def leaves_known_answers(sm, cc, const):
    mdl = R.make_model(1.3, 14.0, 5000.0, EPOCH_JD, 0.0)
    y0a = mdl.get_initial_state(7540, 45, -75, 500e3, 90, -2, gmst=0)
    y0b = mdl.get_initial_state(7800, 0, 0, 120e3, 90, -1, gmst=0)
    eom0 = mdl.equations_of_motion(0.0, y0b)
    y1 = np.array([const.EARTH_R + 60e3, 0, 0, 0, 7000.0, 100.0])
    eom1 = mdl.equations_of_motion(100.0, y1)
    sun0 = sm.sun_position_vector(2451545.0)
    moon0 = sm.moon_position_vector(2451545.0)
    sat = np.array([8000000.0, 0.0, 0.0])
    rho8 = sm.atmosphere_model(8010.012, 0, JD_STAR)[0]
    ka = {
        "moon_position_vector(2451545.0)": fl(moon0),
        "sun_position_vector(2451545.0)": fl(sun0),
        "moon_position_vector(2460310.5)": fl(sm.moon_position_vector(EPOCH_JD)),
        "sun_position_vector(2460310.5)": fl(sm.sun_position_vector(EPOCH_JD)),
        "third_body_acceleration([8e6,0,0],sun(2451545.0),SUN_K)": fl(sm.third_body_acceleration(sat, sun0, const.SUN_K)),
        "third_body_acceleration([8e6,0,0],moon(2451545.0),MOON_K)": fl(sm.third_body_acceleration(sat, moon0, const.MOON_K)),
        "J2_perturbation_numba([7e6,1e6,2e6])": fl(sm.J2_perturbation_numba(np.array([7e6, 1e6, 2e6]), const.EARTH_MU, const.EARTH_J2, const.EARTH_R)),
        "atmosphere_model(102010.012,0,jd*)": fl(sm.atmosphere_model(102010.012, 0, JD_STAR)),
        "solar_activity_factor(jd*,102010.012)": fl(sm.solar_activity_factor(JD_STAR, 102010.012)),
        "atmosphere_model(8010.012,0,jd*)": fl(sm.atmosphere_model(8010.012, 0, JD_STAR)),
        "solar_activity_factor(jd*,8010.012)": fl(sm.solar_activity_factor(JD_STAR, 8010.012)),
        "atmosphere_model(50000,45,jd*)": fl(sm.atmosphere_model(50000, 45, JD_STAR)),
        "atmosphere_model(15000,-30,jd*)": fl(sm.atmosphere_model(15000, -30, JD_STAR)),
        "atmosphere_model(85000,10,jd*)": fl(sm.atmosphere_model(85000, 10, JD_STAR)),
        "atmosphere_model(89999,0,jd*)": fl(sm.atmosphere_model(89999, 0, JD_STAR)),
        "atmosphere_model(90000,0,jd*)": fl(sm.atmosphere_model(90000, 0, JD_STAR)),
        "atmosphere_model(600000,0,jd*)": fl(sm.atmosphere_model(600000, 0, JD_STAR)),
        "atmosphere_model(0,0,jd*)": fl(sm.atmosphere_model(0, 0, JD_STAR)),
        "atmosphere_model(-5,0,jd*)": fl(sm.atmosphere_model(-5, 0, JD_STAR)),
        "atmospheric_drag(2.2,0.1,rho(8010.012),[7500,0,0],500)": fl(sm.atmospheric_drag(2.2, 0.1, rho8, np.array([7500.0, 0, 0]), 500)),
        "eci_to_ecef([7e6,1e6,2e6],1.0)": fl(cc.eci_to_ecef(np.array([7e6, 1e6, 2e6]), 1.0)),
        "ecef_to_eci([7e6,1e6,2e6],1.0)": fl(cc.ecef_to_eci(np.array([7e6, 1e6, 2e6]), 1.0)),
        "geodetic_to_spheroid(45,-75,500e3)": fl(cc.geodetic_to_spheroid(45, -75, 500e3)),
        "ecef_to_geodetic(1260744.91,-4705164.05,4840901.80)": fl(cc.ecef_to_geodetic(1260744.91, -4705164.05, 4840901.80)),
        "enu_to_ecef([100,200,300],45,-75)": fl(cc.enu_to_ecef(np.array([100.0, 200.0, 300.0]), 45, -75)),
        "get_initial_state(7540,45,-75,500e3,90,-2,0)": fl(y0a),
        "get_initial_state(7800,0,0,120e3,90,-1,0)": fl(y0b),
        "get_initial_state(7800,20,10,120e3,45,-1.5,1.25)": fl(mdl.get_initial_state(7800, 20, 10, 120e3, 45, -1.5, gmst=1.25)),
        "eom(0,y0b).acceleration": fl(eom0["acceleration"]),
        "eom(0,y0b).drag_acceleration": fl(eom0["drag_acceleration"]),
        "eom(0,y0b).altitude": fl(eom0["altitude"]),
        "eom(100,[R+60e3,0,0,0,7000,100]).acceleration": fl(eom1["acceleration"]),
        "eom(100,[R+60e3,0,0,0,7000,100]).drag_acceleration": fl(eom1["drag_acceleration"]),
    }
    return ka


def leaves_random(sm, cc, const, n=256, seed=7):
    rng = np.random.default_rng(seed)
    jd = 2451545.0 + rng.uniform(0, 12000, n)
    r = rng.normal(size=(n, 3))
    r *= (rng.uniform(6.4e6, 5e7, n) / np.linalg.norm(r, axis=1))[:, None]
    v = rng.normal(size=(n, 3)) * 3000
    gmst = rng.uniform(0, 3000, n)
    alt = np.where(rng.uniform(size=n) < 0.6, rng.uniform(1, 120e3, n), rng.uniform(1, 1e6, n))
    alt[:8] = [11000.0, 20000.0, 32000.0, 47000.0, 51000.0, 71000.0, 84852.0, 90000.0]  # breakpoints exactly
    lat = rng.uniform(-90, 90, n)
    lon = rng.uniform(-180, 180, n)
    rho = 10 ** rng.uniform(-12, 0, n)
    d = dict(jd=jd, r=r, v=v, gmst=gmst, alt=alt, lat=lat, lon=lon, rho=rho)
    d["moon"] = np.array([sm.moon_position_vector(x) for x in jd])
    d["sun"] = np.array([sm.sun_position_vector(x) for x in jd])
    d["j2"] = np.array([sm.J2_perturbation_numba(x, const.EARTH_MU, const.EARTH_J2, const.EARTH_R) for x in r])
    d["third_moon"] = np.array([sm.third_body_acceleration(x, m, const.MOON_K) for x, m in zip(r, d["moon"])])
    d["third_sun"] = np.array([sm.third_body_acceleration(x, s, const.SUN_K) for x, s in zip(r, d["sun"])])
    d["eci_to_ecef"] = np.array([cc.eci_to_ecef(x, g) for x, g in zip(r, gmst)])
    d["ecef_to_eci"] = np.array([cc.ecef_to_eci(x, g) for x, g in zip(r, gmst)])
    d["ecef_to_geodetic"] = np.array([cc.ecef_to_geodetic(*x) for x in r])
    d["atmosphere"] = np.array([sm.atmosphere_model(a, la, j) for a, la, j in zip(alt, lat, jd)])
    d["solar_factor"] = np.array([sm.solar_activity_factor(j, a) for a, j in zip(alt, jd)])
    d["drag"] = np.array([sm.atmospheric_drag(1.3, 14.0, p, x, 5000.0) for p, x in zip(rho, v)])
    d["geodetic_to_spheroid"] = np.array([cc.geodetic_to_spheroid(la, lo, a) for la, lo, a in zip(lat, lon, alt)])
    d["enu_to_ecef"] = np.array([cc.enu_to_ecef(x, la, lo) for x, la, lo in zip(v, lat, lon)])
    return d


def rhs_random(n=192, seed=11):
    rng = np.random.default_rng(seed)
    Cd, A, m = 1.3, 14.0, 5000.0
    gmst0 = 0.4
    mdl = R.make_model(Cd, A, m, EPOCH_JD, gmst0)
    alt = np.concatenate([rng.uniform(1500, 120e3, n // 2), rng.uniform(120e3, 1e6, n // 4), rng.uniform(1e6, 4e7, n - n // 2 - n // 4)])
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    r = u * (6378137.0 + alt)[:, None]
    v = rng.normal(size=(n, 3)) * 4000
    t = rng.uniform(0, 86400 * 30, n)
    y = np.concatenate([r, v], axis=1)
    keys = ["acceleration", "gravitational_acceleration", "J2_acceleration", "moon_acceleration", "sun_acceleration", "drag_acceleration", "altitude"]
    out = {k: [] for k in keys}
    for ti, yi in zip(t, y):
        d = mdl.equations_of_motion(ti, yi)
        for k in keys:
            out[k].append(np.asarray(d[k], dtype=np.float64))
    res = {k: np.array(v_) for k, v_ in out.items()}
    res.update(t=t, y=y, Cd=Cd, A=A, m=m, epoch_jd=EPOCH_JD, gmst0=gmst0)
    return res


def c1_single(sm):
    Cd, A, m = 1.3, 14.0, 5000.0
    mdl = R.make_model(Cd, A, m, EPOCH_JD, 0.0)
    y0 = mdl.get_initial_state(7800, 0, 0, 120e3, 90, -1, gmst=0)
    n_steps = 800
    o2 = R.o2_scipy_pinned(mdl, y0, 0.0, 1.0, n_steps)
    o3 = R.o3_run(y0[None], Cd, A, m, EPOCH_JD, 0.0, 0.0, 1.0, n_steps)
    # O1: the reference as shipped (adaptive RK45), loose cross-check only
    sol = mdl.run_simulation((0, n_steps), y0, np.arange(0, n_steps, 1.0))
    return dict(Cd=Cd, A=A, m=m, epoch_jd=EPOCH_JD, gmst0=0.0, t0=0.0, dt=1.0, n_steps=n_steps, y0=y0,
                o2_t=o2["t"], o2_y=o2["y"], o2_t_event=o2["t_event"], o2_y_event=o2["y_event"],
                o3_ys=o3["ys"][0], o3_impact_step=o3["impact_step"][0], o3_t_event=o3["t_event"][0],
                o3_y_event=o3["y_event"][0],
                o1_t=sol.t, o1_y=sol.y, o1_t_event=sol.t_events[0][0],
                o1_altitude=sol.additional_data["altitude"], o1_velocity=sol.additional_data["velocity"])


def mc_inputs(n, seed):
    """C2 dispersion: nominal (7800, 20, 0, 120e3, 90, -1), 1-sigma (10, .5, .5, 500, .5, .1)."""
    rng = np.random.default_rng(seed)
    z = rng.normal(size=(n, 6))
    nominal = np.array([7800.0, 20.0, 0.0, 120e3, 90.0, -1.0])
    sigma = np.array([10.0, 0.5, 0.5, 500.0, 0.5, 0.1])
    return nominal + sigma * z


def c2_mc():
    Cd, A, m = 1.3, 14.0, 5000.0
    mdl = R.make_model(Cd, A, m, EPOCH_JD, 0.0)
    params = mc_inputs(48, 1234)
    y0 = np.array([mdl.get_initial_state(*p, gmst=0.0) for p in params])
    n_steps, stride, dt = 2048, 16, 0.5
    o3 = R.o3_run(y0, Cd, A, m, EPOCH_JD, 0.0, 0.0, dt, n_steps, stride)
    o2a = R.o2_scipy_pinned(mdl, y0[0], 0.0, dt, n_steps)
    o2b = R.o2_scipy_pinned(mdl, y0[1], 0.0, dt, n_steps)
    return dict(Cd=Cd, A=A, m=m, epoch_jd=EPOCH_JD, gmst0=0.0, t0=0.0, dt=dt, n_steps=n_steps, stride=stride,
                params=params, y0=y0, ys=o3["ys"], impact_step=o3["impact_step"], t_event=o3["t_event"],
                y_event=o3["y_event"],
                o2_0_y=o2a["y"][:, ::stride], o2_0_t_event=o2a["t_event"], o2_1_y=o2b["y"][:, ::stride],
                o2_1_t_event=o2b["t_event"])


def c3_sso(const, n=8, days=1.0):
    rng = np.random.default_rng(2345)
    Cd = np.full(n, 2.2); A = rng.uniform(0.5, 2, n); m = rng.uniform(50, 500, n)
    gmst0 = 1.0
    mdl = R.make_model(2.2, 1.0, 100.0, EPOCH_JD, gmst0)
    vcirc = np.sqrt(const.EARTH_MU / (const.EARTH_R + 550e3))
    lat = rng.uniform(-80, 80, n); lon = rng.uniform(-180, 180, n)
    params = np.stack([vcirc + rng.normal(0, 1, n), lat, lon, 550e3 + rng.normal(0, 5e3, n), np.full(n, 352.4), np.zeros(n)], axis=1)
    y0 = np.array([mdl.get_initial_state(*p, gmst=gmst0) for p in params])
    n_steps, stride, dt = int(round(8640 * days)), 360, 10.0
    o3 = R.o3_run(y0, Cd, A, m, EPOCH_JD, gmst0, 0.0, dt, n_steps, stride)
    return dict(Cd=Cd, A=A, m=m, epoch_jd=EPOCH_JD, gmst0=gmst0, t0=0.0, dt=dt, n_steps=n_steps, stride=stride,
                params=params, y0=y0, ys=o3["ys"], impact_step=o3["impact_step"])


def kepler_to_state(mu, a, e, inc, raan, argp, nu):
    p = a * (1 - e * e)
    r = p / (1 + e * np.cos(nu))
    rp = np.array([r * np.cos(nu), r * np.sin(nu), 0.0])
    vp = np.sqrt(mu / p) * np.array([-np.sin(nu), e + np.cos(nu), 0.0])
    cO, sO, cw, sw, ci, si = np.cos(raan), np.sin(raan), np.cos(argp), np.sin(argp), np.cos(inc), np.sin(inc)
    Rm = np.array([[cO * cw - sO * sw * ci, -cO * sw - sO * cw * ci, sO * si],
                   [sO * cw + cO * sw * ci, -sO * sw + cO * cw * ci, -cO * si],
                   [sw * si, cw * si, ci]])
    return np.concatenate([Rm @ rp, Rm @ vp])


def c4_geo_heo(const, half=4, days=6.0):
    rng = np.random.default_rng(3456)
    mu = const.EARTH_MU
    y0 = []
    for _ in range(half):  # GEO
        y0.append(kepler_to_state(mu, 42164e3, 0.0, np.radians(abs(rng.normal(0, 0.1))), rng.uniform(0, 2 * np.pi), 0.0, rng.uniform(0, 2 * np.pi)))
    rp, ra = const.EARTH_R + 600e3, const.EARTH_R + 39750e3
    for _ in range(half):  # Molniya-like HEO
        y0.append(kepler_to_state(mu, (rp + ra) / 2, (ra - rp) / (ra + rp), np.radians(63.4), rng.uniform(0, 2 * np.pi), np.radians(270.0), rng.uniform(0, 2 * np.pi)))
    y0 = np.array(y0)
    Cd, A, m = 2.2, 10.0, 2000.0
    n_steps, stride, dt = int(round(1440 * days)), 1440, 60.0
    o3 = R.o3_run(y0, Cd, A, m, EPOCH_JD, 0.3, 0.0, dt, n_steps, stride)
    return dict(Cd=Cd, A=A, m=m, epoch_jd=EPOCH_JD, gmst0=0.3, t0=0.0, dt=dt, n_steps=n_steps, stride=stride,
                y0=y0, ys=o3["ys"], impact_step=o3["impact_step"])


def edge_cases(const):
    mdl = R.make_model(1.0, 1.0, 1.0, EPOCH_JD, 0.0)
    gs = mdl.get_initial_state
    y0 = np.array([
        gs(100.0, 10, 20, 500.0, 90, -5),          # 0: starts BELOW the event altitude: never triggers (direction -1)
        gs(7000.0, -30, 40, 100e3, 120, -30),      # 1: steep entry
        gs(7800.0, 5, 5, 120e3, 80, -1),           # 2: light, draggy object
        gs(7800.0, 5, 5, 120e3, 80, -1),           # 3: heavy, slick object (same state, other Cd/A/m)
        gs(300.0, 0, 0, 1100.0, 90, -60),          # 4: event in the very first steps
        gs(7700.0, 60, -100, 400e3, 30, 0),        # 5: no event within the horizon
        gs(7800.0, 89.9, 0, 120e3, 0, -2),         # 6: near-polar pass
        gs(7800.0, 0, 179.9, 120e3, 270, -1.2),    # 7: retrograde across the date line
    ])
    Cd = np.array([2.2, 1.3, 2.2, 1.0, 1.3, 2.2, 1.3, 1.3])
    A = np.array([1.0, 14.0, 30.0, 1.0, 14.0, 4.0, 14.0, 14.0])
    m = np.array([100.0, 5000.0, 200.0, 8000.0, 5000.0, 800.0, 5000.0, 5000.0])
    n_steps, stride, dt, t0 = 1500, 10, 1.0, 250.0  # non-zero start time
    o3 = R.o3_run(y0, Cd, A, m, EPOCH_JD, 0.7, t0, dt, n_steps, stride)
    return dict(Cd=Cd, A=A, m=m, epoch_jd=EPOCH_JD, gmst0=0.7, t0=t0, dt=dt, n_steps=n_steps, stride=stride,
                y0=y0, ys=o3["ys"], impact_step=o3["impact_step"], t_event=o3["t_event"], y_event=o3["y_event"])


def ground_track(cc, const):
    """N2: the app's post-pass (app.py:225-243 ECI->ECEF->geodetic, :302-308 cumulative haversine,
    spacecraft_visualization.py:562-577 crossing finder restated inline -- that module imports plotly)
    evaluated with the reference's own functions on O3 trajectories."""
    c1 = np.load(os.path.join(GOLD, "c1_single.npz"))
    c2 = np.load(os.path.join(GOLD, "c2_mc.npz"))
    cases = [(c1["o3_ys"][:int(c1["o3_impact_step"])], 0.0, 1.0, 0.37)]
    for j in (0, 5, 17):
        ys = c2["ys"][j]
        ys = ys[~np.isnan(ys[:, 0])]
        cases.append((ys, 0.0, float(c2["dt"]) * int(c2["stride"]), 2.1 + j))
    out = {}
    for k, (ys, t0, dts, gmst0) in enumerate(cases):
        t = t0 + dts * np.arange(len(ys))
        geo = np.array([cc.ecef_to_geodetic(*cc.eci_to_ecef(np.ascontiguousarray(r), gmst0 + const.EARTH_OMEGA * tt))
                        for r, tt in zip(ys[:, :3], t)])
        dr = [0.0]
        for i in range(1, len(geo)):
            dr.append(dr[-1] + cc.haversine_distance(geo[i - 1, 0], geo[i - 1, 1], geo[i, 0], geo[i, 1]))
        alt = np.array([np.sqrt(r[0] ** 2 + r[1] ** 2 + r[2] ** 2) - const.EARTH_R for r in ys[:, :3]])
        cross = [i for i in range(1, len(alt)) if (alt[i] >= 100000 and alt[i - 1] < 100000) or
                 (alt[i] <= 100000 and alt[i - 1] > 100000)]
        out.update({f"ys_{k}": ys, f"t_{k}": t, f"gmst0_{k}": gmst0, f"geo_{k}": geo, f"downrange_{k}": np.array(dr),
                    f"altitude_{k}": alt, f"cross_{k}": np.array(cross, dtype=np.int64)})
    out["n_cases"] = len(cases)
    return out


def thermal(sm, const):
    """N1: thermal keys of the UNMODIFIED equations_of_motion dict (:468, :481-486) on reentry states, for two
    (material, dt, iter_fact) settings, plus direct spacecraft_temperature calls on random inputs."""
    rng = np.random.default_rng(21)
    out = {}
    keys = ["spacecraft_temperature", "spacecraft_heat_flux", "spacecraft_heat_flux_conduction",
            "spacecraft_heat_flux_radiation", "spacecraft_heat_flux_total", "spacecraft_temperature_change"]
    n = 96
    alt = rng.uniform(20e3, 130e3, n)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    y = np.concatenate([u * (const.EARTH_R + alt)[:, None], rng.normal(size=(n, 3)) * 3500], axis=1)
    t = rng.uniform(0, 5000, n)
    for tag, (mat, dt, itf, Cd, A, m) in {"a": ([0.167, 1260, 0.9, 0.7], 10, 2, 1.3, 14.0, 5000.0),
                                           "b": ([237, 903, 0.1, 0.1], 30, 4, 2.2, 20.0, 500.0)}.items():
        mdl = sm.SpacecraftModel(Cd=Cd, A=A, m=m, epoch=R.Epoch(EPOCH_JD), gmst0=0.2, material=mat, dt=dt, iter_fact=itf)
        rows = [mdl.equations_of_motion(ti, yi) for ti, yi in zip(t, y)]
        for k in keys:
            out[f"{tag}_{k}"] = np.array([float(r[k]) for r in rows])
        out[f"{tag}_cfg"] = np.array(mat + [dt, itf, Cd, A, m, mdl.height], dtype=np.float64)
    out.update(t=t, y=y, epoch_jd=EPOCH_JD, gmst0=0.2)
    # direct calls
    args = np.stack([rng.uniform(100, 8000, 64), rng.uniform(180, 300, 64), 10 ** rng.uniform(-6, 2, 64)], axis=1)
    out["direct_args"] = args
    out["direct"] = np.array([sm.spacecraft_temperature(np.array([a[0], 0.0, 0.0]), a[1], np.array([0.0, a[2], 0.0]), 2.7, 10,
                                                        0.167, 1260, 0.9, 0.7, 2, 5000.0) for a in args], dtype=np.float64)
    return out


def adaptive(sm, const, sim_type="RK45"):
    """N4: SpacecraftModel.run_simulation AS SHIPPED (scipy adaptive RK45, rtol 1e-8, atol 1e-10, spacecraft_model.py:516)
    on four launch geometries; plus the accepted-step times of the same solve_ivp call without t_eval.
    sim_type='RK23': the same through the constructor's sim_type (app.py:51 offers it; spacecraft_model.py:516 passes it on)."""
    from scipy.integrate import solve_ivp

    cases = [  # (Cd, A, m, gmst0, (v, lat, lon, alt, az, gamma), tf, dt_out)
        (1.3, 14.0, 5000.0, 0.0, (7800, 0, 0, 120e3, 90, -1), 800, 1.0),            # C1
        (2.2, 20.0, 500.0, 1.7, (7540, 45, -75, 500e3, 90, -2), 3700, 10.0),         # the app's defaults (app.py:37-54)
        (2.2, 4.0, 800.0, 0.3, (7670, 51.6, 10, 400e3, 45, 0), 7200, 30.0),          # no event within the horizon
        (1.3, 14.0, 5000.0, 2.9, (7000, -30, 40, 100e3, 120, -30), 200, 0.5),        # steep entry
    ]
    out = {"n_cases": len(cases)}
    for k, (Cd, A, m, gmst0, geo, tf, dt) in enumerate(cases):
        mdl = sm.SpacecraftModel(Cd=Cd, A=A, m=m, epoch=R.Epoch(EPOCH_JD), gmst0=gmst0, sim_type=sim_type, dt=10, iter_fact=2)
        y0 = mdl.get_initial_state(*geo, gmst=gmst0)
        t_eval = np.arange(0, tf, dt)
        sol = mdl.run_simulation((0, tf), y0, t_eval)

        def rhs(t, y):
            dy = mdl.equations_of_motion(t, y)
            return np.concatenate((dy["velocity"], dy["acceleration"]))

        def ev(t, y):
            return sm.euclidean_norm(y[0:3]) - const.EARTH_R - 1000.0
        ev.terminal = True
        ev.direction = -1
        s2 = solve_ivp(rhs, (0, tf), y0, method=sim_type, rtol=1e-8, atol=1e-10, events=[ev])
        te = sol.t_events[0]
        out.update({f"cfg_{k}": np.array([Cd, A, m, gmst0, tf, dt]), f"y0_{k}": y0, f"t_{k}": sol.t, f"y_{k}": sol.y,
                    f"t_event_{k}": (te[0] if len(te) else np.nan),
                    f"y_event_{k}": (sol.y_events[0][0] if len(te) else np.full(6, np.nan)),
                    f"nfev_{k}": sol.nfev, f"status_{k}": sol.status, f"steps_t_{k}": s2.t,
                    f"altitude_{k}": sol.additional_data["altitude"]})
        if sim_type == "DOP853":
            # The reference against ITSELF with the initial state moved by one part in 1e14 (a few ulp): over its first
            # microsecond steps the 8th-order error estimate is pure rounding noise, so the step sequence -- and, across the
            # atmosphere model's layer boundaries, the result -- moves by far more than rounding.  Recorded as the
            # yardstick for the parity bar of this method (tests/test_adaptive.py).
            sp = mdl.run_simulation((0, tf), y0 * (1 + 1e-14), t_eval)
            n = min(sol.y.shape[1], sp.y.shape[1])
            out[f"self_pos_{k}"] = float(np.max(np.linalg.norm(sp.y[:3, :n] - sol.y[:3, :n], axis=0) / np.linalg.norm(sol.y[:3, :n], axis=0)))
            out[f"self_vel_{k}"] = float(np.max(np.linalg.norm(sp.y[3:, :n] - sol.y[3:, :n], axis=0) / np.linalg.norm(sol.y[3:, :n], axis=0)))
            out[f"self_te_{k}"] = float(abs(sp.t_events[0][0] - te[0])) if len(te) and len(sp.t_events[0]) else 0.0
            out[f"self_nfev_{k}"] = int(sp.nfev)
    out["epoch_jd"] = EPOCH_JD
    out["sim_type"] = sim_type
    return out


def recorded_trajectory():
    """First samples of the trajectory recorded inside temperature_model.py:104-105 (10 s spacing;
    app defaults v=7540, 45N, 75W, 500 km, az 90, gamma -2).  Sample 0 pins get_initial_state
    exactly; the vacuum phase pins gravity + J2 to < 1 m."""
    src = open(os.path.join(R.REFERENCE_DIR, "temperature_model.py")).read()
    out = {}
    for name in ("v_norms", "altitudes"):
        mt = re.search(rf"^{name} = \[(.*?)\]", src, re.M | re.S)
        out[name] = [float(x) for x in mt.group(1).split(",")[:120]]
    out["epoch_jd"] = float(re.search(r"^epoch_jd = ([0-9.]+)", src, re.M).group(1))
    out["sample_spacing_s"] = 10.0
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    sm, cc, const = R.load()
    if "--only-adaptive" in sys.argv:
        np.savez_compressed(os.path.join(GOLD, "adaptive.npz"), **adaptive(sm, const))
        return
    if "--only-dop853" in sys.argv:
        np.savez_compressed(os.path.join(GOLD, "adaptive_dop853.npz"), **adaptive(sm, const, "DOP853"))
        return
    if "--only-rk23" in sys.argv:
        np.savez_compressed(os.path.join(GOLD, "adaptive_rk23.npz"), **adaptive(sm, const, "RK23"))
        return
    if "--only-thermal" in sys.argv:
        np.savez_compressed(os.path.join(GOLD, "thermal.npz"), **thermal(sm, const))
        return
    if "--only-long" in sys.argv:
        np.savez_compressed(os.path.join(GOLD, "c3_sso_7d.npz"), **c3_sso(const, n=64, days=7.0))
        np.savez_compressed(os.path.join(GOLD, "c4_geo_heo_30d.npz"), **c4_geo_heo(const, half=32, days=30.0))
        return
    if "--only-ground-track" in sys.argv:
        np.savez_compressed(os.path.join(GOLD, "ground_track.npz"), **ground_track(cc, const))
        return
    import scipy, numba  # noqa
    meta = dict(generated_by="oracle/make_golden.py", numpy=np.__version__, scipy=scipy.__version__,
                numba=numba.__version__, epoch_jd=EPOCH_JD, jd_star=JD_STAR)
    with open(os.path.join(GOLD, "leaves.json"), "w") as f:
        json.dump(dict(meta=meta, values=leaves_known_answers(sm, cc, const)), f, indent=1)
    np.savez_compressed(os.path.join(GOLD, "leaves_random.npz"), **leaves_random(sm, cc, const))
    np.savez_compressed(os.path.join(GOLD, "rhs_random.npz"), **rhs_random())
    print("leaves done")
    np.savez_compressed(os.path.join(GOLD, "c1_single.npz"), **c1_single(sm))
    print("c1 done")
    np.savez_compressed(os.path.join(GOLD, "c2_mc.npz"), **c2_mc())
    print("c2 done")
    np.savez_compressed(os.path.join(GOLD, "c3_sso.npz"), **c3_sso(const))
    np.savez_compressed(os.path.join(GOLD, "c4_geo_heo.npz"), **c4_geo_heo(const))
    np.savez_compressed(os.path.join(GOLD, "c3_sso_7d.npz"), **c3_sso(const, n=64, days=7.0))
    np.savez_compressed(os.path.join(GOLD, "c4_geo_heo_30d.npz"), **c4_geo_heo(const, half=32, days=30.0))
    print("c3/c4 done")
    np.savez_compressed(os.path.join(GOLD, "edge.npz"), **edge_cases(const))
    np.savez_compressed(os.path.join(GOLD, "ground_track.npz"), **ground_track(cc, const))
    np.savez_compressed(os.path.join(GOLD, "thermal.npz"), **thermal(sm, const))
    np.savez_compressed(os.path.join(GOLD, "adaptive.npz"), **adaptive(sm, const))
    np.savez_compressed(os.path.join(GOLD, "adaptive_rk23.npz"), **adaptive(sm, const, "RK23"))
    with open(os.path.join(GOLD, "recorded_trajectory.json"), "w") as f:
        json.dump(recorded_trajectory(), f)
    print("wrote", sorted(os.listdir(GOLD)))


if __name__ == "__main__":
    main()
