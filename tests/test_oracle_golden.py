"""Pins the CPU oracle (oracle/reentry_oracle.c) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py; spacecraft_model.py / coordinate_converter.py).

Bars: scalar leaves bit-exact (same libm, same operation order); the three matrix-product
leaves (eci_to_ecef, ecef_to_eci, enu_to_ecef -- numba sends `@`/np.dot to OpenBLAS, whose
FMA use is CPU-dependent) within 2 ulp of the vector norm; whole trajectories <= 1e-12
relative; impact step identical; impact time within 1e-9 s.
"""
import json
import os

import numpy as np
import pytest

EPS = np.finfo(float).eps


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def vec_close(a, b, ulps=2):
    a = np.asarray(a, float); b = np.asarray(b, float)
    scale = np.linalg.norm(b, axis=-1, keepdims=True)
    return np.all(np.abs(a - b) <= ulps * EPS * scale)


def test_known_answers(oracle, golden_dir):
    O = oracle
    g = json.load(open(os.path.join(golden_dir, "leaves.json")))
    v, jd = g["values"], g["meta"]["jd_star"]
    MU, J2, R = 398600441800000.0, 0.00108263, 6378137.0
    exact = {
        "moon_position_vector(2451545.0)": O.moon_position_vector(2451545.0),
        "sun_position_vector(2451545.0)": O.sun_position_vector(2451545.0),
        "moon_position_vector(2460310.5)": O.moon_position_vector(2460310.5),
        "sun_position_vector(2460310.5)": O.sun_position_vector(2460310.5),
        "third_body_acceleration([8e6,0,0],sun(2451545.0),SUN_K)": O.third_body_acceleration([8e6, 0, 0], O.sun_position_vector(2451545.0), 1.32712440042e20),
        "third_body_acceleration([8e6,0,0],moon(2451545.0),MOON_K)": O.third_body_acceleration([8e6, 0, 0], O.moon_position_vector(2451545.0), 4.9048695e12),
        "J2_perturbation_numba([7e6,1e6,2e6])": O.j2_perturbation([7e6, 1e6, 2e6], MU, J2, R),
        "atmosphere_model(102010.012,0,jd*)": O.atmosphere_model(102010.012, 0, jd),
        "solar_activity_factor(jd*,102010.012)": [O.solar_activity_factor(jd, 102010.012)],
        "atmosphere_model(8010.012,0,jd*)": O.atmosphere_model(8010.012, 0, jd),
        "solar_activity_factor(jd*,8010.012)": [O.solar_activity_factor(jd, 8010.012)],
        "atmosphere_model(50000,45,jd*)": O.atmosphere_model(50000, 45, jd),
        "atmosphere_model(15000,-30,jd*)": O.atmosphere_model(15000, -30, jd),
        "atmosphere_model(85000,10,jd*)": O.atmosphere_model(85000, 10, jd),
        "atmosphere_model(89999,0,jd*)": O.atmosphere_model(89999, 0, jd),
        "atmosphere_model(90000,0,jd*)": O.atmosphere_model(90000, 0, jd),
        "atmosphere_model(600000,0,jd*)": O.atmosphere_model(600000, 0, jd),
        "atmosphere_model(0,0,jd*)": O.atmosphere_model(0, 0, jd),
        "atmosphere_model(-5,0,jd*)": O.atmosphere_model(-5, 0, jd),
        "atmospheric_drag(2.2,0.1,rho(8010.012),[7500,0,0],500)": O.atmospheric_drag(2.2, 0.1, O.atmosphere_model(8010.012, 0, jd)[0], [7500, 0, 0], 500),
        "geodetic_to_spheroid(45,-75,500e3)": O.geodetic_to_spheroid(45, -75, 500e3),
        "ecef_to_geodetic(1260744.91,-4705164.05,4840901.80)": O.ecef_to_geodetic(1260744.91, -4705164.05, 4840901.80),
    }
    for k, got in exact.items():
        assert list(map(float, got)) == v[k], k
    blas = {
        "eci_to_ecef([7e6,1e6,2e6],1.0)": O.eci_to_ecef([7e6, 1e6, 2e6], 1.0),
        "ecef_to_eci([7e6,1e6,2e6],1.0)": O.ecef_to_eci([7e6, 1e6, 2e6], 1.0),
        "enu_to_ecef([100,200,300],45,-75)": O.enu_to_ecef([100, 200, 300], 45, -75),
    }
    for k, got in blas.items():
        assert vec_close(got, v[k]), k
    for k, args in {
        "get_initial_state(7540,45,-75,500e3,90,-2,0)": (7540, 45, -75, 500e3, 90, -2, 0.0),
        "get_initial_state(7800,0,0,120e3,90,-1,0)": (7800, 0, 0, 120e3, 90, -1, 0.0),
        "get_initial_state(7800,20,10,120e3,45,-1.5,1.25)": (7800, 20, 10, 120e3, 45, -1.5, 1.25),
    }.items():
        y = O.get_initial_state(*args)
        assert vec_close(y[:3], v[k][:3]) and vec_close(y[3:], v[k][3:]), k
    # the survey's headline known answers, spelled out
    assert v["moon_position_vector(2451545.0)"] == [-359325029.8704828, -160751475.53313145, 34707652.30058943]
    assert v["get_initial_state(7800,0,0,120e3,90,-1,0)"][0] == 6498137.0
    y0b = O.get_initial_state(7800, 0, 0, 120e3, 90, -1, 0.0)
    dy, parts = O.equations_of_motion(0.0, y0b, 2460310.5, 0.0, 1.3, 14.0, 5000.0)
    assert vec_close(parts["acceleration"], v["eom(0,y0b).acceleration"], ulps=8)
    assert vec_close(parts["drag_acceleration"], v["eom(0,y0b).drag_acceleration"], ulps=8)
    assert parts["altitude"] == v["eom(0,y0b).altitude"][0] == 120000.0
    y1 = [6378137.0 + 60e3, 0, 0, 0, 7000.0, 100.0]
    dy, parts = O.equations_of_motion(100.0, y1, 2460310.5, 0.0, 1.3, 14.0, 5000.0)
    assert vec_close(parts["acceleration"], v["eom(100,[R+60e3,0,0,0,7000,100]).acceleration"], ulps=8)
    assert vec_close(parts["drag_acceleration"], v["eom(100,[R+60e3,0,0,0,7000,100]).drag_acceleration"], ulps=8)


def test_random_leaves(oracle, golden_dir):
    O = oracle
    g = load(golden_dir, "leaves_random.npz")
    MU, J2, R = 398600441800000.0, 0.00108263, 6378137.0
    n = len(g["jd"])
    for i in range(n):
        jd, r, v, gm = g["jd"][i], g["r"][i], g["v"][i], g["gmst"][i]
        alt, lat, lon, rho = g["alt"][i], g["lat"][i], g["lon"][i], g["rho"][i]
        assert np.array_equal(O.moon_position_vector(jd), g["moon"][i])
        assert np.array_equal(O.sun_position_vector(jd), g["sun"][i])
        assert np.array_equal(O.j2_perturbation(r, MU, J2, R), g["j2"][i])
        assert np.array_equal(O.third_body_acceleration(r, g["moon"][i], 4.9048695e12), g["third_moon"][i])
        assert np.array_equal(O.third_body_acceleration(r, g["sun"][i], 1.32712440042e20), g["third_sun"][i])
        assert vec_close(O.eci_to_ecef(r, gm), g["eci_to_ecef"][i])
        assert vec_close(O.ecef_to_eci(r, gm), g["ecef_to_eci"][i])
        assert np.array_equal(O.ecef_to_geodetic(*r), g["ecef_to_geodetic"][i])
        assert np.array_equal(O.atmosphere_model(alt, lat, jd), g["atmosphere"][i])
        assert O.solar_activity_factor(jd, alt) == g["solar_factor"][i]
        assert np.array_equal(O.atmospheric_drag(1.3, 14.0, rho, v, 5000.0), g["drag"][i])
        assert np.array_equal(O.geodetic_to_spheroid(lat, lon, alt), g["geodetic_to_spheroid"][i])
        assert vec_close(O.enu_to_ecef(v, lat, lon), g["enu_to_ecef"][i])


def test_rhs_random(oracle, golden_dir):
    """Unmodified SpacecraftModel.equations_of_motion on random states, term by term."""
    O = oracle
    g = load(golden_dir, "rhs_random.npz")
    names = {"acceleration": "acceleration", "gravitational_acceleration": "gravitational_acceleration",
             "J2_acceleration": "J2_acceleration", "moon_acceleration": "moon_acceleration",
             "sun_acceleration": "sun_acceleration", "drag_acceleration": "drag_acceleration"}
    for i in range(len(g["t"])):
        dy, p = O.equations_of_motion(g["t"][i], g["y"][i], float(g["epoch_jd"]), float(g["gmst0"]),
                                      float(g["Cd"]), float(g["A"]), float(g["m"]))
        assert p["altitude"] == g["altitude"][i]
        for k in names:
            ref = g[k][i]
            assert np.all(np.abs(p[k] - ref) <= 1e-13 * np.linalg.norm(ref) + 1e-300), (k, i)
        assert np.array_equal(dy[:3], g["y"][i][3:])


def relerr(a, b):
    return np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1))


def test_c1_single_capsule(oracle, golden_dir):
    """C1: O2 (scipy pinned, unmodified Python RHS) and O3 (njit leaves): impact step 695,
    t_event 694.13091187."""
    O = oracle
    g = load(golden_dir, "c1_single.npz")
    r = O.propagate(g["y0"][None], float(g["Cd"]), float(g["A"]), float(g["m"]), float(g["epoch_jd"]),
                    float(g["gmst0"]), float(g["t0"]), float(g["dt"]), int(g["n_steps"]))
    assert r["impact_step"][0] == int(g["o3_impact_step"]) == 695
    assert r["status"][0] == 1 and r["n_samples"][0] == 695
    assert abs(r["impact_time"][0] - float(g["o3_t_event"])) < 1e-9
    assert abs(r["impact_time"][0] - float(g["o2_t_event"])) < 1e-9
    assert abs(r["impact_time"][0] - 694.13091187) < 1e-8
    s = r["samples"][0, :695]
    assert relerr(s[:, :3], g["o3_ys"][:695, :3]) < 1e-12 and relerr(s[:, 3:6], g["o3_ys"][:695, 3:]) < 1e-12
    o2 = g["o2_y"].T
    assert o2.shape[0] == 695  # scipy emits the grid samples strictly before the event
    assert relerr(s[:, :3], o2[:, :3]) < 1e-12 and relerr(s[:, 3:6], o2[:, 3:]) < 1e-12
    assert np.all(np.isnan(r["samples"][0, 695:]))
    assert np.allclose(r["final_state"][0], g["o3_y_event"], rtol=1e-11, atol=1e-9)
    assert np.allclose(r["final_state"][0], g["o2_y_event"], rtol=1e-11, atol=1e-9)
    # altitude / speed columns are the reference's additional_data (spacecraft_model.py:458, app.py:264)
    assert np.allclose(s[:, 6], np.linalg.norm(s[:, :3], axis=1) - 6378137.0, rtol=0, atol=1e-8)
    assert np.allclose(s[:, 7], np.linalg.norm(s[:, 3:6], axis=1), rtol=1e-15)
    # O1 (the reference as shipped, adaptive): loose agreement only (~19 m)
    n1 = min(len(g["o1_t"]), 695)
    assert np.max(np.linalg.norm(g["o1_y"].T[:n1, :3] - s[:n1, :3], axis=1)) < 100.0
    assert abs(float(g["o1_t_event"]) - r["impact_time"][0]) < 0.05


@pytest.mark.parametrize("name", ["c2_mc.npz", "edge.npz"])
def test_reentry_ensembles(oracle, golden_dir, name):
    O = oracle
    g = load(golden_dir, name)
    r = O.propagate(g["y0"], g["Cd"], g["A"], g["m"], float(g["epoch_jd"]), float(g["gmst0"]), float(g["t0"]),
                    float(g["dt"]), int(g["n_steps"]), int(g["stride"]))
    assert np.array_equal(r["impact_step"], g["impact_step"].astype(np.int32))
    hit = g["impact_step"] > 0
    assert np.allclose(r["impact_time"][hit], g["t_event"][hit], rtol=0, atol=1e-9)
    assert np.all(np.isnan(r["impact_time"][~hit]))
    assert np.allclose(r["final_state"][hit], g["y_event"][hit], rtol=1e-10, atol=1e-8)
    ys = g["ys"]
    assert np.array_equal(np.isnan(ys[:, :, 0]), np.isnan(r["samples"][:, :, 0]))
    ok = ~np.isnan(ys[:, :, 0])
    assert relerr(r["samples"][ok][:, :3], ys[ok][:, :3]) < 1e-12
    assert relerr(r["samples"][ok][:, 3:6], ys[ok][:, 3:]) < 1e-12
    if name == "c2_mc.npz":
        for j in (0, 1):
            o2 = g[f"o2_{j}_y"].T
            n = o2.shape[0]
            assert relerr(r["samples"][j, :n, :3], o2[:, :3]) < 1e-12
            assert abs(r["impact_time"][j] - float(g[f"o2_{j}_t_event"])) < 1e-9
    else:
        assert r["impact_step"][0] == -1 and r["status"][0] == 0      # starts below the event altitude
        assert r["impact_step"][5] == -1 and r["status"][5] == 0      # no event within the horizon
        assert r["impact_step"][2] != r["impact_step"][3]            # Cd/A/m are per trajectory


@pytest.mark.parametrize("name,tol", [("c3_sso.npz", 1e-11), ("c4_geo_heo.npz", 1e-11),
                                      ("c3_sso_7d.npz", 1e-10), ("c4_geo_heo_30d.npz", 1e-10)])
def test_orbit_ensembles(oracle, golden_dir, name, tol):
    O = oracle
    g = load(golden_dir, name)
    r = O.propagate(g["y0"], g["Cd"], g["A"], g["m"], float(g["epoch_jd"]), float(g["gmst0"]), float(g["t0"]),
                    float(g["dt"]), int(g["n_steps"]), int(g["stride"]))
    assert np.all(r["impact_step"] == -1) and np.all(g["impact_step"] == -1)
    assert relerr(r["samples"][:, :, :3], g["ys"][:, :, :3]) < tol
    assert relerr(r["samples"][:, :, 3:6], g["ys"][:, :, 3:]) < tol


def test_recorded_trajectory(oracle, golden_dir):
    """The only recorded data in the reference repo (temperature_model.py:104-105): sample 0 pins
    get_initial_state exactly; the vacuum phase pins gravity + J2 to < 1 m."""
    O = oracle
    g = json.load(open(os.path.join(golden_dir, "recorded_trajectory.json")))
    y0 = O.get_initial_state(7540, 45, -75, 500e3, 90, -2, 0.0)
    assert np.linalg.norm(y0[:3]) - 6378137.0 == g["altitudes"][0] == 489349.92941798735
    assert abs(np.linalg.norm(y0[3:]) - g["v_norms"][0]) < 1e-11
    r = O.propagate(y0[None], 2.2, 20.0, 500.0, g["epoch_jd"], 0.0, 0.0, 1.0, 1000, 10)
    n = 100
    assert np.max(np.abs(r["samples"][0, :n, 6] - np.array(g["altitudes"][:n]))) < 1.0
    assert np.max(np.abs(r["samples"][0, :n, 7] - np.array(g["v_norms"][:n]))) < 1e-2


def test_tableau_matches_scipy(oracle):
    """The fixed-step contract inherits scipy's RK45 tableau (rk.py:541-565)."""
    from scipy.integrate._ivp.rk import RK45

    c, a, b, p = oracle.tableau()
    assert np.array_equal(c, RK45.C) and np.array_equal(b, RK45.B)
    assert np.array_equal(a, np.asarray(RK45.A)[:6, :5]) and np.array_equal(p, RK45.P)
