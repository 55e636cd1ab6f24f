"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(ctypes -> libreentry_b200.so); the CPU oracle and the golden vectors of the unmodified
reference are the checkers.

Tolerances (BASELINE.json north_star): per-step position <= 1e-9 relative AND < 1 m absolute,
velocity <= 1e-9 relative, identical impact step except where the crossing is within 1e-6 m of
the step boundary; impact time within 1e-7 s.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POS_REL, VEL_REL, POS_ABS, T_ABS = 1e-9, 1e-9, 1.0, 1e-7


@pytest.fixture(scope="module")
def torch():
    import torch as t

    if not t.cuda.is_available():
        pytest.skip("no GPU")
    return t


@pytest.fixture(scope="module")
def Model(torch):
    from reentry_old_b200 import SpacecraftModel

    return SpacecraftModel


def compare_states(got, ref, what=""):
    """got/ref [..., 6] with identical NaN pattern."""
    ok = ~np.isnan(ref[..., 0])
    assert np.array_equal(ok, ~np.isnan(got[..., 0])), f"{what}: sample validity differs"
    g, r = got[ok], ref[ok]
    dpos = np.linalg.norm(g[:, :3] - r[:, :3], axis=1)
    rel = dpos / np.linalg.norm(r[:, :3], axis=1)
    relv = np.linalg.norm(g[:, 3:6] - r[:, 3:6], axis=1) / np.linalg.norm(r[:, 3:6], axis=1)
    assert rel.max() <= POS_REL, (what, "pos rel", rel.max())
    near = np.linalg.norm(r[:, :3], axis=1) < 1e9   # the < 1 m bar is for trajectories near the Earth (an edge-case
    assert dpos[near].max() < POS_ABS, (what, "pos abs", dpos[near].max())  # lane is flung to 1e13 m through r -> 0)
    assert relv.max() <= VEL_REL, (what, "vel rel", relv.max())
    return rel.max(), relv.max()


def run_golden(Model, g, **kw):
    m = Model(Cd=float(np.atleast_1d(g["Cd"])[0]), A=float(np.atleast_1d(g["A"])[0]), m=float(np.atleast_1d(g["m"])[0]),
              epoch=float(g["epoch_jd"]), gmst0=float(g["gmst0"]))
    per = {k: (g[key] if np.ndim(g[key]) else None) for k, key in (("Cd", "Cd"), ("A", "A"), ("m", "m"))}
    stride = int(g["stride"]) if "stride" in g else 1
    return m.run_ensemble(g["y0"].reshape(-1, 6), t0=float(g["t0"]), dt=float(g["dt"]), n_steps=int(g["n_steps"]),
                          out_stride=stride, **per, **kw)


def samples_np(res):
    return res.samples.permute(2, 0, 1).cpu().numpy()  # [N, n_out, 8]


# ------------------------------------------------------------------------------- leaves
def test_initial_states_vs_oracle(Model, oracle):
    rng = np.random.default_rng(3)
    n = 2000
    p = np.stack([rng.uniform(100, 9000, n), rng.uniform(-89.9, 89.9, n), rng.uniform(-180, 180, n),
                  rng.uniform(0, 1e6, n), rng.uniform(0, 360, n), rng.uniform(-90, 90, n)], axis=1)
    m = Model()
    for gmst in (0.0, 1.25, 2000.5):
        got = m.get_initial_states(*p.T, gmst=gmst).cpu().numpy()
        ref = oracle.get_initial_states(*p.T, gmst)
        assert np.max(np.linalg.norm(got[:, :3] - ref[:, :3], axis=1) / np.linalg.norm(ref[:, :3], axis=1)) < 1e-14
        assert np.max(np.linalg.norm(got[:, 3:] - ref[:, 3:], axis=1) / np.linalg.norm(ref[:, 3:], axis=1)) < 1e-13
    y = m.get_initial_state(7800, 0, 0, 120e3, 90, -1, gmst=0)   # scalar call surface of the reference
    assert isinstance(y, np.ndarray) and y.shape == (6,) and y[0] == 6498137.0
    assert np.allclose(y, [6498137.0, 0.0, 0.0, -136.12877021081138, 7798.812022219852, 4.77539509008172e-13],
                       rtol=1e-14, atol=1e-12)


def test_equations_of_motion_vs_golden(Model, golden_dir):
    g = np.load(os.path.join(golden_dir, "rhs_random.npz"))
    m = Model(Cd=float(g["Cd"]), A=float(g["A"]), m=float(g["m"]), epoch=float(g["epoch_jd"]), gmst0=float(g["gmst0"]))
    d = m.equations_of_motion_batch(g["t"], g["y"])
    for key in ("acceleration", "gravitational_acceleration", "J2_acceleration", "moon_acceleration",
                "sun_acceleration", "drag_acceleration"):
        got, ref = d[key].cpu().numpy(), g[key]
        err = np.abs(got - ref).max(axis=1)
        # 1e-12 of the term; the third-body terms are differences of ~6e-3 m/s^2 pieces (1e-17 floor)
        tol = 5e-12 if key in ("drag_acceleration", "acceleration") else 1e-13   # exp() of |exponents| ~ 1e2 in the density tail
        assert np.all(err <= tol * np.linalg.norm(ref, axis=1) + 1e-17), (key, err.max())
    assert np.allclose(d["altitude"].cpu().numpy(), g["altitude"], rtol=0, atol=2e-8)  # 1 ulp of |r| (fma in the norm)
    one = m.equations_of_motion(float(g["t"][5]), g["y"][5])        # dict surface of the reference (:472-480)
    assert np.allclose(one["acceleration"], g["acceleration"][5], rtol=5e-12, atol=1e-17)
    assert abs(one["altitude"] - g["altitude"][5]) < 2e-8


# ------------------------------------------------------------------------- trajectories
def test_c1_run_simulation_drop_in(Model, golden_dir):
    """The reference's own call sequence (app.py:110,140,171,190) on C1."""
    g = np.load(os.path.join(golden_dir, "c1_single.npz"))
    m = Model(Cd=1.3, A=14.0, m=5000.0, epoch=float(g["epoch_jd"]), gmst0=0.0)
    y0 = m.get_initial_state(7800, 0, 0, 120e3, 90, -1, gmst=0.0)
    calls = []
    sol = m.run_simulation((0, 800), y0, np.arange(0, 800, 1.0), progress_callback=lambda p, e: calls.append(p), integrator="fixed")
    assert calls and calls[-1] == 1.0
    assert sol.y.shape == (6, 695) and sol.t.shape == (695,) and sol.t[-1] == 694.0
    assert sol.status == 1 and len(sol.t_events[0]) == 1
    assert abs(sol.t_events[0][0] - float(g["o3_t_event"])) < T_ABS
    assert abs(sol.t_events[0][0] - float(g["o2_t_event"])) < T_ABS
    compare_states(sol.y.T, g["o3_ys"][:695], "C1 vs O3")
    compare_states(sol.y.T, g["o2_y"].T, "C1 vs O2 (scipy pinned, unmodified RHS)")
    assert np.allclose(sol.y_events[0][0], g["o3_y_event"], rtol=1e-9, atol=1e-6)
    ad = sol.additional_data
    assert ad["altitude"].shape == (695,) and ad["velocity"].shape == (695, 3) and ad["acceleration"].shape == (695, 3)
    assert np.allclose(ad["altitude"], np.linalg.norm(sol.y[:3], axis=0) - 6378137.0, rtol=0, atol=1e-8)
    assert np.array_equal(ad["velocity"], sol.y[3:6].T)
    # loose agreement with the reference as shipped (adaptive): ~19 m
    n1 = min(len(g["o1_t"]), 695)
    assert np.max(np.linalg.norm(g["o1_y"].T[:n1, :3] - sol.y.T[:n1, :3], axis=1)) < 100.0
    assert np.max(np.abs(g["o1_altitude"][:n1] - ad["altitude"][:n1])) < 100.0


@pytest.mark.parametrize("name", ["c2_mc.npz", "edge.npz"])
def test_golden_reentry(Model, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    res = run_golden(Model, g)
    s = samples_np(res)
    compare_states(s[:, :, :6], g["ys"], name)
    assert np.array_equal(res.impact_step.cpu().numpy(), g["impact_step"].astype(np.int32))
    hit = g["impact_step"] > 0
    it = res.impact_time.cpu().numpy()
    assert np.allclose(it[hit], g["t_event"][hit], rtol=0, atol=T_ABS) and np.all(np.isnan(it[~hit]))
    fs = res.final_state.t().cpu().numpy()
    assert np.allclose(fs[hit], g["y_event"][hit], rtol=1e-9, atol=1e-5)
    st = res.status.cpu().numpy()
    assert np.array_equal(st == 1, hit)
    ns = res.n_samples.cpu().numpy()
    assert np.array_equal(ns, (~np.isnan(g["ys"][:, :, 0])).sum(axis=1))
    # altitude / speed columns
    ok = ~np.isnan(s[:, :, 0])
    rn = np.linalg.norm(s[ok][:, :3], axis=1)
    assert np.all(np.abs(s[ok][:, 6] - (rn - 6378137.0)) <= 4e-16 * rn)
    assert np.allclose(s[ok][:, 7], np.linalg.norm(s[ok][:, 3:6], axis=1), rtol=1e-15)


@pytest.mark.parametrize("name", ["c3_sso.npz", "c4_geo_heo.npz", "c3_sso_7d.npz", "c4_geo_heo_30d.npz"])
def test_golden_orbits(Model, golden_dir, name):
    """Compared horizon, strict bars: 1 day x 8 and 7 days x 64 satellites (C3), 6 days x 8 and 30 days x 64 (C4);
    every golden is the output of the unmodified reference's functions (oracle/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, name))
    res = run_golden(Model, g, steps_per_launch=1024)
    rel, relv = compare_states(samples_np(res)[:, :, :6], g["ys"], name)
    print(f"{name}: max rel position {rel:.2e}, velocity {relv:.2e}")
    assert np.all(res.impact_step.cpu().numpy() == -1) and np.all(res.status.cpu().numpy() == 0)


@pytest.mark.parametrize("name", ["c3_sso.npz", "c4_geo_heo.npz", "c2_mc.npz"])
def test_skip_negligible_drag_is_exact(Model, torch, golden_dir, name):
    """RB_FLAG_SKIP_NEGLIGIBLE_DRAG: bit-identical states above 400 km (the skipped term is < 1e-40 m/s^2),
    and no effect at all on a reentry that stays below the threshold."""
    g = np.load(os.path.join(golden_dir, name))
    a = run_golden(Model, g, steps_per_launch=512)
    b = run_golden(Model, g, steps_per_launch=512, skip_negligible_drag=True)
    assert torch.equal(torch.nan_to_num(a.samples, nan=-1.0), torch.nan_to_num(b.samples, nan=-1.0))
    assert torch.equal(a.final_state, b.final_state) and torch.equal(a.impact_step, b.impact_step)
    compare_states(samples_np(b)[:, :, :6], g["ys"], name + " (skip)")


def test_mc_vs_oracle_512(Model, oracle):
    """C2-shaped ensemble against the oracle on the same seeded inputs."""
    from reentry_old_b200.configs import reentry_dispersion_params

    n, n_steps, dt, stride = 512, 2048, 0.5, 16
    p = reentry_dispersion_params(n, seed=4321)
    m = Model(Cd=1.3, A=14.0, m=5000.0, epoch=2460310.5, gmst0=0.0)
    y0 = m.get_initial_states(*p.T, gmst=0.0)
    res = m.run_ensemble(y0, dt=dt, n_steps=n_steps, out_stride=stride)
    ref = oracle.propagate(y0.cpu().numpy(), 1.3, 14.0, 5000.0, 2460310.5, 0.0, 0.0, dt, n_steps, stride)
    compare_states(samples_np(res)[:, :, :6], ref["samples"][:, :, :6], "MC-512")
    assert np.array_equal(res.impact_step.cpu().numpy(), ref["impact_step"])
    assert np.allclose(res.impact_time.cpu().numpy(), ref["impact_time"], rtol=0, atol=T_ABS)
    assert np.array_equal(res.n_samples.cpu().numpy(), ref["n_samples"].astype(np.int32))
    assert np.array_equal(res.status.cpu().numpy(), ref["status"])
    assert np.allclose(res.final_state.t().cpu().numpy(), ref["final_state"], rtol=1e-9, atol=1e-5)


# -------------------------------------------------------------- launch-structure invariance
def test_results_do_not_depend_on_launch_structure(Model, torch):
    """Bit-identical results for any steps_per_launch / compaction / early-exit setting
    (trajectories are independent; FSAL state and tiles carry no launch dependence)."""
    from reentry_old_b200.configs import reentry_dispersion_params

    n, n_steps = 3000, 1800
    p = reentry_dispersion_params(n, seed=77)
    p[:, 5] -= 0.5  # steeper: impacts spread over many launches
    m = Model(Cd=1.3, A=14.0, m=5000.0)
    y0 = m.get_initial_states(*p.T)
    base = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=7)
    for kw in (dict(steps_per_launch=16), dict(steps_per_launch=50), dict(steps_per_launch=n_steps),
               dict(compact=False), dict(early_exit=False), dict(compact=False, early_exit=False, steps_per_launch=100)):
        r = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=7, **kw)
        for a, b in ((r.samples, base.samples), (r.final_state, base.final_state), (r.impact_time, base.impact_time)):
            assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)), kw
        assert torch.equal(r.impact_step, base.impact_step) and torch.equal(r.status, base.status), kw
        assert torch.equal(r.n_samples, base.n_samples), kw


def test_two_partition_overlap_is_invisible(Model, torch):
    """Ensembles >= 600k trajectories run as two partitions on two streams; results are bit-identical
    to the single-stream run and the caller's stream ordering holds (results complete after its sync)."""
    from reentry_old_b200.configs import reentry_dispersion_params

    n = 640_001   # odd: unequal halves
    m = Model(Cd=1.3, A=14.0, m=5000.0)
    y0 = m.get_initial_states(*reentry_dispersion_params(n, seed=21).T)
    a = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=128, overlap=True)
    b = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=128, overlap=False)
    torch.cuda.synchronize()
    assert a.stats["propagate_launches"] > b.stats["propagate_launches"]
    assert torch.equal(torch.nan_to_num(a.samples, nan=-1.0), torch.nan_to_num(b.samples, nan=-1.0))
    for k in ("final_state", "impact_step", "status", "n_samples"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert torch.equal(torch.nan_to_num(a.impact_time), torch.nan_to_num(b.impact_time))
    assert int((a.status == 1).sum()) == n
    c = m.run_ensemble(y0, dt=0.5, n_steps=2048, out_stride=128, overlap=True, early_exit=False, compact=False)
    assert torch.equal(c.impact_step, a.impact_step) and torch.equal(c.final_state, a.final_state)


def test_checkpoint_resume_equals_one_shot(Model, torch):
    """RB_FLAG_RESUME: the output buffers are the checkpoint; chunked propagation == one call, bit for bit."""
    from reentry_old_b200.configs import reentry_dispersion_params

    n, n_steps, stride = 4000, 2048, 16
    p = reentry_dispersion_params(n, seed=41)
    p[::5, 5] -= 0.8                                   # impacts spread over several chunks
    m = Model(Cd=1.3, A=14.0, m=5000.0)
    y0 = m.get_initial_states(*p.T)
    one = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride)
    for chunk in (16, 336, 1024, 4096):
        seen = []
        for res in m.run_ensemble_chunked(y0, 0.5, n_steps, chunk, out_stride=stride):
            seen.append(int((res.status == 1).sum()))
        assert seen == sorted(seen) and seen[-1] == int((one.status == 1).sum())
        assert res.n_steps == n_steps
        assert torch.equal(torch.nan_to_num(res.samples, nan=-1.0), torch.nan_to_num(one.samples, nan=-1.0)), chunk
        for k in ("final_state", "impact_step", "status", "n_samples"):
            assert torch.equal(getattr(res, k), getattr(one, k)), (chunk, k)
        assert torch.equal(torch.nan_to_num(res.impact_time), torch.nan_to_num(one.impact_time)), chunk


def test_soa_input_and_per_trajectory_bodies(Model, torch, oracle):
    rng = np.random.default_rng(8)
    n = 300
    from reentry_old_b200.configs import reentry_dispersion_params

    p = reentry_dispersion_params(n, seed=5)
    m = Model(Cd=9.9, A=9.9, m=9.9)  # scalars must be ignored when arrays are given
    y0 = m.get_initial_states(*p.T)
    Cd, A, mass = rng.uniform(1, 2.2, n), rng.uniform(1, 30, n), rng.uniform(200, 8000, n)
    y0_soa_view = y0.t().contiguous().t()  # [N,6] view with strides (1, N)
    assert y0_soa_view.stride() == (1, n)
    r = m.run_ensemble(y0_soa_view, dt=0.5, n_steps=2600, out_stride=50, Cd=Cd, A=A, m=mass)
    ref = oracle.propagate(y0.cpu().numpy(), Cd, A, mass, m.epoch, m.gmst0, 0.0, 0.5, 2600, 50)
    compare_states(samples_np(r)[:, :, :6], ref["samples"][:, :, :6], "per-trajectory Cd/A/m")
    assert np.array_equal(r.impact_step.cpu().numpy(), ref["impact_step"])


# ----------------------------------------------------------------- size-independent properties
def test_large_ensemble_properties(Model, torch, oracle):
    """200k trajectories: properties that need no oracle run at full size + a strided oracle subsample."""
    from reentry_old_b200.configs import reentry_dispersion_params

    n, n_steps, stride = 200_000, 2048, 16
    p = reentry_dispersion_params(n, seed=1234)
    m = Model(Cd=1.3, A=14.0, m=5000.0)
    y0 = m.get_initial_states(*p.T)
    r = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride)
    assert int((r.status == 1).sum()) == n
    step = r.impact_step.to(torch.int64)
    assert int(step.min()) > 900 and int(step.max()) < n_steps
    t_imp = r.impact_time
    assert bool(((t_imp > (step - 1) * 0.5) & (t_imp <= step * 0.5)).all())           # event inside its step
    assert torch.equal(r.n_samples.to(torch.int64), (step - 1) // stride + 1)          # samples strictly before it
    valid = ~torch.isnan(r.samples[:, 0, :])                                           # [n_out, N]
    assert torch.equal(valid.sum(0).to(torch.int32), r.n_samples)                      # NaN exactly after the event
    assert bool((valid[:-1] >= valid[1:]).all())                                       # validity is a prefix
    alt_final = r.final_state[:3].norm(dim=0) - 6378137.0
    assert float((alt_final - 1000.0).abs().max()) < 1e-6                              # dense-output root sits at 1000 m
    alt = torch.where(valid, r.samples[:, 6, :], torch.full_like(r.samples[:, 6, :], 1e9))
    assert float(alt.min()) > 1000.0                                                   # no stored sample below the event
    sub = np.arange(0, n, n // 64)[:64]
    ref = oracle.propagate(y0[sub].cpu().numpy(), 1.3, 14.0, 5000.0, m.epoch, m.gmst0, 0.0, 0.5, n_steps, stride)
    compare_states(samples_np(r)[sub][:, :, :6], ref["samples"][:, :, :6], "200k subsample")
    assert np.array_equal(r.impact_step[sub].cpu().numpy(), ref["impact_step"])
    assert r.stats["propagate_launches"] < n_steps // 32  # early exit stopped the launch sequence


# ----------------------------------------------------------------------- host-buffer (e2e) path
def test_propagate_host_matches_device_path(Model, torch):
    import ctypes as C

    from reentry_old_b200 import _lib as L
    from reentry_old_b200.configs import reentry_dispersion_params
    from reentry_old_b200.model import DeviceContext

    n, n_steps, stride = 5000, 2048, 16
    p = reentry_dispersion_params(n, seed=11)
    m = Model(Cd=1.3, A=14.0, m=5000.0)
    y0 = m.get_initial_states(*p.T)
    r = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride)
    ctx = DeviceContext.get()
    n_out = n_steps // stride + 1
    y0h = y0.cpu().numpy()
    samples = np.full((n_out, 8, n), np.nan)
    fs = np.empty((n, 6)); istep = np.empty(n, np.int32); itime = np.empty(n); st = np.empty(n, np.uint8)
    ns = np.empty(n, np.int32); used = C.c_int64(0)
    run = L.RbRun(first_step=0, n_steps=n_steps, out_stride=stride, steps_per_launch=32, event_altitude=1000.0,
                  flags=L.RB_FLAG_COMPACT)
    rc = ctx.lib.rb_propagate_host(ctx.handle, n, y0h.ctypes.data, None, None, None, 1.3, 14.0, 5000.0, m.epoch, m.gmst0,
                                   0.0, 0.5, C.byref(run), samples.ctypes.data, n_out, fs.ctypes.data, istep.ctypes.data,
                                   itime.ctypes.data, st.ctypes.data, ns.ctypes.data, C.byref(used))
    assert rc == 0, ctx.lib.rb_ctx_last_error(ctx.handle)
    assert np.array_equal(istep, r.impact_step.cpu().numpy()) and np.array_equal(st, r.status.cpu().numpy())
    assert np.array_equal(itime, r.impact_time.cpu().numpy()) and np.array_equal(ns, r.n_samples.cpu().numpy())
    assert np.array_equal(fs, r.final_state.t().cpu().numpy())
    k = used.value
    assert ns.max() <= k <= n_out
    assert np.array_equal(np.nan_to_num(samples[:k], nan=-1.0), np.nan_to_num(r.samples[:k].cpu().numpy(), nan=-1.0))


def test_propagate_host_zero_copy_samples(Model, torch):
    """Pinned sample buffers: valid entries equal the device path bit for bit whichever way the rows travel;
    entries after each trajectory's event are NaN or keep the caller's fill value."""
    import ctypes as C

    from reentry_old_b200 import _lib as L
    from reentry_old_b200.configs import reentry_dispersion_params
    from reentry_old_b200.model import DeviceContext

    n, n_steps, stride = 20_000, 2048, 16
    m = Model(Cd=1.3, A=14.0, m=5000.0)
    y0 = m.get_initial_states(*reentry_dispersion_params(n, seed=12).T)
    r = m.run_ensemble(y0, dt=0.5, n_steps=n_steps, out_stride=stride)
    ctx = DeviceContext.get()
    n_out = n_steps // stride + 1
    y0h = y0.cpu().pin_memory()
    fs = torch.empty((n, 6), dtype=torch.float64).pin_memory(); ist = torch.empty(n, dtype=torch.int32).pin_memory()
    it = torch.empty(n, dtype=torch.float64).pin_memory(); st = torch.empty(n, dtype=torch.uint8).pin_memory()
    ns = torch.empty(n, dtype=torch.int32).pin_memory(); used = C.c_int64(0)
    # pinned buffer, default: copy-engine rows while the partition is dense, then direct stores;
    # RB_FLAG_DIRECT_SAMPLES: direct stores only; RB_FLAG_STAGED_SAMPLES: device staging + NaN fill
    for flags in (L.RB_FLAG_COMPACT, L.RB_FLAG_COMPACT | L.RB_FLAG_DIRECT_SAMPLES, L.RB_FLAG_COMPACT | L.RB_FLAG_STAGED_SAMPLES):
        samples = torch.full((n_out, 8, n), -7.0, dtype=torch.float64).pin_memory()
        run = L.RbRun(first_step=0, n_steps=n_steps, out_stride=stride, steps_per_launch=32, event_altitude=1000.0, flags=flags)
        rc = ctx.lib.rb_propagate_host(ctx.handle, n, y0h.data_ptr(), None, None, None, 1.3, 14.0, 5000.0, m.epoch, m.gmst0, 0.0,
                                       0.5, C.byref(run), samples.data_ptr(), n_out, fs.data_ptr(), ist.data_ptr(), it.data_ptr(),
                                       st.data_ptr(), ns.data_ptr(), C.byref(used))
        assert rc == 0, ctx.lib.rb_ctx_last_error(ctx.handle)
        k = used.value
        ref = r.samples[:k].cpu()
        valid = ~torch.isnan(ref)
        assert torch.equal(samples[:k][valid], ref[valid]), flags
        rest = samples[:k][~valid]
        if flags & L.RB_FLAG_STAGED_SAMPLES:
            assert bool(torch.isnan(rest).all())
        elif flags & L.RB_FLAG_DIRECT_SAMPLES:
            assert bool((rest == -7.0).all())                     # untouched: the caller's fill value
        else:
            assert bool((torch.isnan(rest) | (rest == -7.0)).all())
        assert bool((samples[k:] == -7.0).all())
        assert torch.equal(ns, r.n_samples.cpu()) and torch.equal(ist, r.impact_step.cpu())


def test_error_codes(Model, torch):
    import ctypes as C

    from reentry_old_b200 import _lib as L
    from reentry_old_b200.model import DeviceContext

    lib = L.load()
    h = L.vp()
    assert lib.rb_ctx_create(C.byref(h), 9999) == -1           # invalid device
    assert lib.rb_ctx_create(C.byref(h), 0) == 0
    rin, run, out = L.RbIn(n=4), L.RbRun(n_steps=10, out_stride=1), L.RbOut()
    assert lib.rb_propagate(h, C.byref(rin), C.byref(run), C.byref(out), None) == -1   # null buffers
    y0 = torch.zeros(4, 6, dtype=torch.float64, device="cuda")
    fs = torch.zeros(6, 4, dtype=torch.float64, device="cuda")
    st = torch.zeros(4, dtype=torch.uint8, device="cuda")
    ist = torch.zeros(4, dtype=torch.int32, device="cuda")
    rin = L.RbIn(n=4, y0=y0.data_ptr(), y0_traj_stride=6, y0_comp_stride=1, cd_scalar=1, area_scalar=1, mass_scalar=1)
    out = L.RbOut(final_state=fs.data_ptr(), status=st.data_ptr(), impact_step=ist.data_ptr())
    assert lib.rb_propagate(h, C.byref(rin), C.byref(run), C.byref(out), None) == -3   # no time table
    tab = np.empty((lib.rb_time_table_entries(5), 16))
    assert lib.rb_build_time_table(2460310.5, 0.0, 0.0, 1.0, 5, tab.ctypes.data, 1) == 0
    assert lib.rb_ctx_set_time_table(h, tab.ctypes.data, 5, 0.0, 1.0, None) == 0
    assert lib.rb_propagate(h, C.byref(rin), C.byref(run), C.byref(out), None) == -4   # 10 steps > table of 5
    run.out_stride = 0
    assert lib.rb_propagate(h, C.byref(rin), C.byref(run), C.byref(out), None) == -1
    assert lib.rb_ctx_destroy(h) == 0
    m = Model(sim_type="DOP853")
    with pytest.raises(NotImplementedError):
        m.run_ensemble(np.zeros((1, 6)), dt=1.0, n_steps=1)
    with pytest.raises(ValueError):
        Model().run_simulation((0, 10), np.ones(6), np.array([0.0, 1.0, 3.0]), integrator="fixed")          # non-uniform t_eval
    # zero trajectories / zero steps are valid and empty
    r = Model().run_ensemble(np.zeros((0, 6)), dt=1.0, n_steps=5)
    assert r.samples.shape == (6, 8, 0)
    y = Model().get_initial_state(7800, 0, 0, 120e3, 90, -1)
    r = Model().run_ensemble(y[None], dt=1.0, n_steps=0)
    assert r.samples.shape == (1, 8, 1) and int(r.status[0]) == 0 and np.allclose(r.samples[0, :6, 0].cpu().numpy(), y)


def test_nonfinite_lane_is_retired(Model, torch):
    m = Model()
    y = m.get_initial_state(7800, 0, 0, 300e3, 90, 0)
    y0 = np.stack([y, [0.0, 0.0, 0.0, 1.0, 0.0, 0.0], y])     # r = 0 -> division by zero -> NaN
    r = m.run_ensemble(y0, dt=1.0, n_steps=64, out_stride=8)
    assert r.status.cpu().tolist() == [0, 2, 0]
    assert torch.equal(torch.nan_to_num(r.samples[:, :, 0]), torch.nan_to_num(r.samples[:, :, 2]))
    assert int(r.impact_step[1]) == 1 and int(r.n_samples[1]) == 1


def test_math_primitives(torch):
    """The guard-free FP64 primitives of csrc/rb_math.cuh against numpy (float64, correctly rounded libm)."""
    from reentry_old_b200 import _lib as L
    from reentry_old_b200.model import DeviceContext

    ctx = DeviceContext.get()
    rng = np.random.default_rng(0)
    n = 200_000

    def run(op, x, y=None):
        xd = torch.as_tensor(x, device="cuda")
        yd = torch.as_tensor(y, device="cuda") if y is not None else None
        out = torch.empty_like(xd)
        ctx.check(ctx.lib.rb_math_probe(ctx.handle, op, n, xd.data_ptr(), yd.data_ptr() if yd is not None else None,
                                        out.data_ptr(), None))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    def ulps(got, ref):
        return np.max(np.abs(got - ref) / np.spacing(np.abs(ref)))

    x = 10 ** rng.uniform(-30, 30, n)
    assert ulps(run(0, x), 1.0 / x) <= 1.0
    assert ulps(run(1, x), 1.0 / np.sqrt(x)) <= 2.0   # vs a twice-rounded reference
    x = np.concatenate([rng.uniform(-700, 5, n - 4), [0.0, -1e-300, -707.9, 1.0]])
    assert ulps(run(2, x), np.exp(x)) <= 1.5
    assert run(2, np.full(n, -1e5)).max() < 1e-300 and np.isnan(run(2, np.full(n, np.nan))).all()
    x = rng.uniform(0.7, 1.3, n)
    assert np.max(np.abs(run(3, x) - np.log(x))) <= 1.2e-16
    # pow as the density uses it: (T/T_i)^(-gM/(R* L_i)) inside each gradient layer (constants.py:124-127)
    layer = rng.integers(0, 5, n)
    lo = np.array([0.7518, 1.0, 1.0, 0.7930, 0.8709])[layer]; hi = np.array([1.0, 1.0554, 1.1837, 1.0, 1.0])[layer]
    k = np.array([5.2558, -34.163, -12.201, 12.201, 17.081])[layer]
    x = rng.uniform(lo, hi)
    assert np.max(np.abs(run(5, x, k) / np.power(x, k) - 1.0)) <= 2e-15
    z = x - 1.0                                                        # T/T_i - 1 = (L/T_i) dh, as density() forms it
    assert np.max(np.abs(run(6, z) - np.log1p(z))) <= 1.2e-16
    assert np.max(np.abs(run(7, z, k) / np.power(1.0 + z, k) - 1.0)) <= 1e-14   # k <= 34: exponent error k ulp(log)
    # the budgeted variants of the hot RHS (FAST paths of rb_math.cuh / rb_physics.cuh), each against its stated bound
    x = rng.uniform(-3.3, 0.9, n)                                      # exponents of the density's two exponentials
    assert np.max(np.abs(run(8, x) / np.exp(x) - 1.0)) <= 3e-15        # r^6/720 dropped: 2.2e-15
    assert np.max(np.abs(run(9, z) / np.log1p(np.where(z == 0, 1e-300, z)) - 1.0)) <= 2.5e-15   # u^8/17 dropped: 1.6e-15
    x = 10 ** rng.uniform(-4, 16, n)                                   # |v_rel|^2, T, 1 + e^s, |d|^2
    # one Newton step on the MUFU seed (relative error e/2 <= 9e-7): 3/8 e^2 for sqrt / rsqrt, e^2 for rcp, 15/8 e^2 for x^-1.5
    errs = dict(sqrt_q=np.max(np.abs(run(10, x) / np.sqrt(x) - 1.0)), rsqrt_q=np.max(np.abs(run(12, x) * np.sqrt(x) - 1.0)),
                rcp_q=np.max(np.abs(run(13, x) * x - 1.0)))
    assert ulps(run(15, x), np.sqrt(x)) <= 1.5                         # full-precision sqrt (|v_rel| of the drag term)
    kk = rng.uniform(1e12, 2e20, n)
    errs["k3"] = np.max(np.abs(run(11, x, kk) / (kk * x ** -1.5) - 1.0))
    print("second-order forms, max relative error:", errs)
    assert errs["sqrt_q"] <= 2e-12 and errs["rsqrt_q"] <= 2e-12 and errs["rcp_q"] <= 5e-12 and errs["k3"] <= 1e-11, errs
    th = np.concatenate([rng.uniform(0, np.pi / 2, n - 5), [0.0, np.pi / 2, np.pi / 4, 1e-9, np.pi / 2 - 1e-9]])
    s, c = np.sin(th), np.cos(th)
    nrm = np.hypot(s, c); s, c = s / nrm, c / nrm
    assert np.max(np.abs(run(4, s, c) - np.arctan2(s, c))) <= 4.5e-16
    assert np.max(np.abs(run(14, s, c) - np.arctan2(s, c))) <= 1.5e-13   # hot-RHS variant: u^3/7 dropped (6e-14), rcp by one Newton step


def test_fp64_peak_probe(torch):
    from reentry_old_b200.model import DeviceContext

    tf = DeviceContext.get().fp64_peak_tflops(100.0)
    assert 5.0 < tf < 80.0, tf
