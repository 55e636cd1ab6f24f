"""N1: the thermal entries of the equations_of_motion dict
(spacecraft_model.py:223-255, :468, :481-486), against the unmodified reference (tests/golden/thermal.npz)."""
import os

import numpy as np
import pytest

KEYS = ["spacecraft_temperature", "spacecraft_heat_flux", "spacecraft_heat_flux_conduction",
        "spacecraft_heat_flux_radiation", "spacecraft_heat_flux_total", "spacecraft_temperature_change"]
# dict key -> index in spacecraft_temperature's return tuple (Qc, Qr, Q_net, Q, T_s, dT) after the
# reference's permuted unpacking q_gen, q_c, q_r, q_net, T_s, dT (spacecraft_model.py:468, :481-486)
RET_INDEX = {"spacecraft_temperature": 4, "spacecraft_heat_flux": 3, "spacecraft_heat_flux_conduction": 1,
             "spacecraft_heat_flux_radiation": 2, "spacecraft_heat_flux_total": 0, "spacecraft_temperature_change": 5}


def close(got, ref, rtol, atol):
    """elementwise closeness; the reference's explicit Euler loop overflows to +-inf / NaN on hot samples,
    which must be reproduced as such"""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    fin = np.isfinite(ref)
    same_special = np.array_equal(np.isnan(ref), np.isnan(got)) and np.array_equal(ref[np.isinf(ref)], got[np.isinf(ref)])
    return same_special and np.all(np.abs(got[fin] - ref[fin]) <= rtol * np.abs(ref[fin]) + atol)


def test_oracle_thermal_vs_reference(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "thermal.npz"))
    got = np.array([oracle.spacecraft_temperature(a[0], a[1], a[2], 2.7, 10, 0.167, 1260, 0.9, 0.7, 2, 5000.0)
                    for a in g["direct_args"]])
    assert np.array_equal(got, g["direct"])
    for tag in ("a", "b"):
        k_, cp, em, ab, dt, itf, Cd, A, m, height = g[f"{tag}_cfg"]
        for i in range(len(g["t"])):
            _, p = oracle.equations_of_motion(g["t"][i], g["y"][i], float(g["epoch_jd"]), float(g["gmst0"]), Cd, A, m)
            ret = oracle.spacecraft_temperature(np.linalg.norm(p["v_rel"]), p["T"], np.linalg.norm(p["drag_acceleration"]),
                                                height, dt, k_, cp, em, ab, itf, m)
            for key in KEYS:
                ref = g[f"{tag}_{key}"][i]
                assert close(ret[RET_INDEX[key]], ref, 1e-12, 1e-300), (tag, key, i)
    assert np.isnan(oracle.spacecraft_temperature(100, 200, 1, 2.7, 1, 1, 1, 1, 1, 2, 500)[0])   # int(1/2) == 0 iterations


@pytest.mark.gpu
def test_device_thermal_vs_reference(golden_dir):
    import torch

    from reentry_old_b200 import SpacecraftModel

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    g = np.load(os.path.join(golden_dir, "thermal.npz"))
    for tag in ("a", "b"):
        k_, cp, em, ab, dt, itf, Cd, A, m, height = g[f"{tag}_cfg"]
        mdl = SpacecraftModel(Cd=Cd, A=A, m=m, epoch=float(g["epoch_jd"]), gmst0=float(g["gmst0"]),
                              material=[k_, cp, em, ab], dt=dt, iter_fact=itf)
        assert abs(mdl.height - height) < 1e-12
        d = mdl.equations_of_motion_batch(g["t"], g["y"])
        for key in KEYS:
            got, ref = d[key].cpu().numpy(), g[f"{tag}_{key}"]
            assert close(got, ref, 1e-9, 1e-12), (tag, key)
        one = mdl.equations_of_motion(float(g["t"][3]), g["y"][3])
        assert set(KEYS) <= set(one) and close(one["spacecraft_temperature"], g[f"{tag}_spacecraft_temperature"][3], 1e-9, 1e-12)
