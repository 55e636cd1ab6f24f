"""N2: ground-track post-pass -- per-sample ECI->ECEF->geodetic (app.py:225-243),
cumulative haversine down-range (app.py:302-308, coordinate_converter.py:125-136) and 100 km
crossings (spacecraft_visualization.py:562-577).  Golden values come from the reference's own
functions (oracle/make_golden.py:ground_track)."""
import os

import numpy as np
import pytest


def cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "ground_track.npz"))
    for k in range(int(g["n_cases"])):
        yield {n: g[f"{n}_{k}"] for n in ("ys", "t", "gmst0", "geo", "downrange", "altitude", "cross")}


def test_oracle_ground_track_vs_reference(oracle, golden_dir):
    for c in cases(golden_dir):
        geo, dr, ci = oracle.ground_track(c["ys"][:, :3], c["t"], c["altitude"], float(c["gmst0"]))
        assert np.max(np.abs(geo[:, 0] - c["geo"][:, 0])) < 1e-13              # the reference's rotation goes through BLAS
        assert np.max(np.abs(geo[:, 1] - c["geo"][:, 1])) < 1e-13              # longitude: BLAS rotation rounding
        assert np.max(np.abs(geo[:, 2] - c["geo"][:, 2])) < 1e-8
        assert np.max(np.abs(dr - c["downrange"])) < 1e-7
        assert np.array_equal(ci, c["cross"])
    assert oracle.haversine_distance(0, 0, 0, 90) == pytest.approx(np.pi / 2 * 6378137.0, rel=1e-15)


@pytest.mark.gpu
def test_device_ground_track_vs_reference(oracle, golden_dir):
    import torch

    from reentry_old_b200 import SpacecraftModel
    from reentry_old_b200.model import EnsembleResult

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    for c in cases(golden_dir):
        ys, n_out = c["ys"], len(c["ys"])
        # two trajectories: the golden one and a truncated copy (tests n_samples handling)
        samples = torch.full((n_out, 8, 2), float("nan"), dtype=torch.float64, device="cuda")
        samples[:, :6, 0] = torch.as_tensor(ys, device="cuda")
        samples[:, 6, 0] = torch.as_tensor(c["altitude"], device="cuda")
        half = n_out // 2
        samples[:half, :, 1] = samples[:half, :, 0]
        dt = float(c["t"][1] - c["t"][0])
        res = EnsembleResult(t=c["t"], samples=samples, final_state=None, impact_step=None, impact_time=None, status=None,
                             n_samples=torch.tensor([n_out, half], dtype=torch.int32, device="cuda"), dt=dt, out_stride=1,
                             n_steps=n_out - 1)
        m = SpacecraftModel(gmst0=float(c["gmst0"]))
        gt = m.ground_track(res)
        geo = gt["geo"][:, :, 0].cpu().numpy()
        assert np.max(np.abs(geo[:, 0] - c["geo"][:, 0])) < 1e-12              # degrees
        dlon = (geo[:, 1] - c["geo"][:, 1] + 180.0) % 360.0 - 180.0
        assert np.max(np.abs(dlon)) < 1e-12
        assert np.max(np.abs(geo[:, 2] - c["geo"][:, 2])) < 1e-6               # metres
        dr = gt["downrange"][:, 0].cpu().numpy()
        assert np.max(np.abs(dr - c["downrange"])) < 1e-5 and dr[0] == 0.0
        nc = gt["n_crossings"].cpu().numpy()
        assert nc[0] == len(c["cross"])
        if len(c["cross"]):
            i0 = int(c["cross"][0])
            assert gt["crossing_time"][0].item() == c["t"][i0]
            assert abs(gt["crossing_downrange"][0].item() - c["downrange"][i0]) < 1e-5
        else:
            assert np.isnan(gt["crossing_time"][0].item())
        # truncated copy: identical prefix, NaN after its last valid sample
        assert torch.equal(gt["geo"][:half, :, 1], gt["geo"][:half, :, 0]) and bool(torch.isnan(gt["geo"][half:, :, 1]).all())


@pytest.mark.gpu
def test_impact_points(oracle):
    import torch

    from reentry_old_b200 import SpacecraftModel
    from reentry_old_b200.configs import reentry_dispersion_params

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    m = SpacecraftModel(Cd=1.3, A=14.0, m=5000.0, gmst0=0.8)
    p = reentry_dispersion_params(256, seed=3)
    y0 = m.get_initial_states(*p.T, gmst=0.8)
    res = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=64)
    ip = m.impact_points(res).cpu().numpy()
    fs, it = res.final_state.t().cpu().numpy(), res.impact_time.cpu().numpy()
    for i in range(0, 256, 17):
        g = oracle.ecef_to_geodetic(*oracle.eci_to_ecef(fs[i, :3], 0.8 + 7.292114146686322e-05 * it[i]))
        assert abs(ip[0, i] - g[0]) < 1e-12 and abs(ip[1, i] - g[1]) < 1e-12 and abs(ip[2, i] - g[2]) < 1e-6
    # the event is at 1000 m SPHERICAL altitude (spacecraft_model.py:499); above the ellipsoid at ~17 deg that is ~2.8 km
    assert np.all((ip[2] > 1000.0) & (ip[2] < 4000.0))
