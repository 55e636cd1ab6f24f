"""Host-side mirror of the reference's spacecraft-model call surface for the propagation path.

`SpacecraftModel` keeps the names, argument meaning and result layout of
`/root/reference/spacecraft_model.py:392-519` (constructor :393, get_initial_state :412,
equations_of_motion :437, run_simulation :490) and adds the ensemble entry point
`run_ensemble`.  Every number is produced by the sm_100a kernels behind the C ABI
(include/reentry_b200.h); torch is used only for device memory and streams.  There is no
CPU fallback: without a CUDA device or without the built library the calls raise.

Fixed-step contract: the reference integrates with scipy's adaptive
RK45; this path integrates with the same Dormand-Prince tableau at a FIXED step (the spacing
of `t_eval`, or `max_step`), which is the reference's own stepper with step-size control
switched off (`sim_type` must be 'RK45' there).  The adaptive mode -- run_simulation's default, the reference as
shipped -- also offers 'RK23' and 'DOP853'; the implicit Radau / BDF / LSODA raise NotImplementedError.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib as L

DEFAULT_EPOCH_JD = 2460310.5  # Time('2024-01-01 00:00:00').jd, the reference's default (spacecraft_model.py:393)
EARTH_R = 6378137.0


class Epoch:
    """Anything with `.jd` works as `epoch` (the reference passes an astropy Time)."""

    def __init__(self, jd: float = DEFAULT_EPOCH_JD):
        self.jd = float(jd)


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise L.ReentryB200Error("no CUDA device: the B200 propagation path has no CPU fallback")
    return torch


class DeviceContext:
    """One rb_ctx per (process, device)."""

    _cache: dict = {}

    def __init__(self, device: int):
        self.lib = L.load()
        self.device = int(device)
        h = L.vp()
        L.check(self.lib.rb_ctx_create(C.byref(h), self.device), None, "rb_ctx_create")
        self.handle = h
        self.table_key = None

    @classmethod
    def get(cls, device: Optional[int] = None) -> "DeviceContext":
        torch = _torch()
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device(device).index if not isinstance(device, int) else device
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    def check(self, rc, what=""):
        L.check(rc, self.handle, what)

    def close(self):
        if self.handle:
            self.lib.rb_ctx_destroy(self.handle)
            self.handle = None
            self._cache.pop(self.device, None)

    # -- time table -------------------------------------------------------------------
    def stage_time_table(self, epoch_jd, gmst0, t0, dt, n_steps, stream_ptr, n_threads=0):
        # the C context is the source of truth: its generation counter changes whenever anything (this method, a
        # *_host call, another binding) restages the table, and the cached key is only trusted while it has not
        key = (float(epoch_jd), float(gmst0), float(t0), float(dt))
        gen = self.lib.rb_ctx_table_generation(self.handle)
        if self.table_key is not None and self.table_key[:4] == key and self.table_key[4] >= n_steps and self.table_key[5] == gen:
            return
        torch = _torch()
        n_entries = self.lib.rb_time_table_entries(n_steps)
        host = torch.empty(n_entries * L.RB_TABLE_ENTRY_DOUBLES, dtype=torch.float64).pin_memory()
        self.check(self.lib.rb_build_time_table(epoch_jd, gmst0, t0, dt, n_steps, host.data_ptr(), n_threads),
                   "rb_build_time_table")
        self.check(self.lib.rb_ctx_set_time_table(self.handle, host.data_ptr(), n_steps, t0, dt, stream_ptr),
                   "rb_ctx_set_time_table")
        torch.cuda.current_stream(self.device).synchronize()  # pinned staging buffer is released below
        self.table_key = key + (int(n_steps), self.lib.rb_ctx_table_generation(self.handle))

    def stats(self):
        s = L.RbStats()
        self.check(self.lib.rb_ctx_stats(self.handle, C.byref(s)))
        return dict(kernel_launches=s.kernel_launches, propagate_launches=s.propagate_launches,
                    steps_enqueued=s.steps_enqueued, last_polled_live=s.last_polled_live,
                    propagate_ms=s.propagate_ms, sample_bytes_sent=s.sample_bytes_sent)

    def fp64_peak_tflops(self, ms=200.0):
        torch = _torch()
        out = L.f64(0)
        self.check(self.lib.rb_fp64_peak_probe(self.handle, float(ms), C.byref(out),
                                               torch.cuda.current_stream(self.device).cuda_stream))
        return out.value


class PackedSamples:
    """Samples as PACKED SLABS (include/reentry_b200.h, rb_sample_slab): the rows each launch wrote, only as wide as
    the ensemble that was still alive.  `arena` is a flat float64 tensor (device, or pinned host memory for the host
    path), `slabs` a list of dicts.  Column j of a slab is the j-th trajectory (ascending index) of its partition with
    impact_step == -1 or > first_step: the column map is derived from the summaries, no index array is stored."""

    def __init__(self, arena, slabs, impact_step, status, n_samples, n, n_out):
        self.arena, self.slabs, self.n, self.n_out = arena, slabs, int(n), int(n_out)
        self.impact_step, self.status, self.n_samples = impact_step, status, n_samples

    @staticmethod
    def slabs_from_ctypes(arr, count):
        return [dict(offset=a.offset, first_row=a.first_row, n_rows=a.n_rows, first_traj=a.first_traj, n_traj=a.n_traj,
                     width=a.width, first_step=a.first_step, n_live=a.n_live) for a in arr[:count]]

    @property
    def used(self):     # doubles of the arena in use
        return max((s["offset"] + s["n_rows"] * L.RB_SAMPLE_FIELDS * s["width"] for s in self.slabs), default=0)

    def _resolve_live(self):
        # rb_propagate is asynchronous: its host descriptors carry n_live = -1 and the counts sit in a device array
        sl = getattr(self, "slab_live", None)
        if sl is not None and any(s["n_live"] < 0 for s in self.slabs):
            live = sl.cpu().tolist()
            for i, s in enumerate(self.slabs):
                if s["n_live"] < 0:
                    s["n_live"] = int(live[i])

    def columns(self, slab):
        """Trajectory indices of the slab's columns (tensor on the device of impact_step)."""
        import torch

        self._resolve_live()
        a, b = slab["first_traj"], slab["first_traj"] + slab["n_traj"]
        if slab["n_live"] == slab["n_traj"]:   # nothing retired yet -- or a run without compaction, whose list never shrinks
            return torch.arange(a, b, device=self.impact_step.device)
        keep = (self.status[a:b] == L.TRAJ_ALIVE) | (self.impact_step[a:b] > slab["first_step"])
        return torch.nonzero(keep, as_tuple=False).flatten() + a

    def view(self, slab):
        """[n_rows, 8, width] view of the slab in the arena."""
        sz = slab["n_rows"] * L.RB_SAMPLE_FIELDS * slab["width"]
        return self.arena[slab["offset"]: slab["offset"] + sz].view(slab["n_rows"], L.RB_SAMPLE_FIELDS, slab["width"])

    def row(self, k):
        """Sample k of the ensemble: (trajectory indices [m], values [8, m]) of the trajectories alive at the launch
        that wrote the row (entries of a trajectory that retired inside that launch before step k*out_stride are NaN)."""
        import torch

        ids, vals = [], []
        for s in self.slabs:
            if s["first_row"] <= k < s["first_row"] + s["n_rows"]:
                c = self.columns(s)
                ids.append(c)
                vals.append(self.view(s)[k - s["first_row"], :, : c.numel()].to(c.device))
        if not ids:
            raise IndexError(k)
        return torch.cat(ids), torch.cat(vals, dim=1)

    def to_dense(self, n_out=None):
        """[n_out, 8, N] with NaN wherever there is no valid sample (the layout of non-packed runs)."""
        import torch

        n_out = self.n_out if n_out is None else int(n_out)
        dense = torch.full((n_out, L.RB_SAMPLE_FIELDS, self.n), float("nan"), dtype=torch.float64, device=self.arena.device)
        for s in self.slabs:
            c = self.columns(s).to(self.arena.device)
            dense[s["first_row"]: s["first_row"] + s["n_rows"], :, c] = self.view(s)[:, :, : c.numel()]
        return dense

    def trajectory(self, i):
        """[n_samples_i, 8] samples of trajectory i."""
        import torch

        out = []
        for s in self.slabs:
            if s["first_traj"] <= i < s["first_traj"] + s["n_traj"]:
                c = self.columns(s)
                j = torch.searchsorted(c, torch.tensor([i], device=c.device))
                if j.item() < c.numel() and c[j].item() == i:
                    out.append((s["first_row"], self.view(s)[:, :, j.item()]))
        out.sort(key=lambda t: t[0])
        rows = torch.cat([v for _, v in out], dim=0) if out else torch.empty((0, L.RB_SAMPLE_FIELDS), dtype=torch.float64)
        return rows[: int(self.n_samples[i])]


@dataclass
class EnsembleResult:
    """Device-resident result of `run_ensemble` (torch tensors on the GPU unless noted)."""
    t: np.ndarray                      # [n_out] sample times (host)
    samples: Optional["object"]        # [n_out, 8, N] f64: x,y,z,vx,vy,vz,altitude,speed (NaN after the event)
    final_state: "object"              # [6, N]
    impact_step: "object"              # [N] int32, -1 = none
    impact_time: "object"              # [N] f64, NaN = none
    status: "object"                   # [N] uint8: 0 alive, 1 impacted, 2 non-finite
    n_samples: "object"                # [N] int32
    dt: float = 0.0
    out_stride: int = 1
    n_steps: int = 0
    stats: dict = field(default_factory=dict)
    packed: Optional[PackedSamples] = None   # run_ensemble(..., samples="packed"): `samples` is then None

    @property
    def y(self):          # [N, 6, n_out] view (sol.y of trajectory i is y[i])
        return self.samples[:, :6, :].permute(2, 1, 0)

    @property
    def altitude(self):   # [N, n_out]
        return self.samples[:, 6, :].transpose(0, 1)

    @property
    def speed(self):      # [N, n_out]
        return self.samples[:, 7, :].transpose(0, 1)

    @property
    def t_events(self):   # impact times of the trajectories that hit, host
        it = self.impact_time.cpu().numpy()
        return it[~np.isnan(it)]


class OdeResultLike(dict):
    """scipy OdeResult stand-in (attribute access), plus `.additional_data` as the reference adds (:518)."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__


class SpacecraftModel:
    def __init__(self, Cd=2.2, A=20.0, m=500.0, epoch=None, gmst0=0.0, sim_type="RK45",
                 material=(233, 1, 1, 0.1), dt=10, iter_fact=2, device=None, integrator="adaptive"):
        # same attributes as spacecraft_model.py:394-410
        self.Cd = Cd
        self.A = A
        self.height = float(np.sqrt(self.A / np.pi) * 1.315)
        self.m = m
        if epoch is None:
            epoch = Epoch()
        self.epoch = float(epoch.jd) if hasattr(epoch, "jd") else float(epoch)
        self.start_time = time.time()
        self.A_over_m = self.A / self.m
        self.gmst0 = float(gmst0)  # radians in use (app.py:109)
        self.thickness = 0.1
        self.sim_type = sim_type
        self.mat = material
        self.thermal_conductivity, self.specific_heat_capacity, self.emissivity, self.ablation_efficiency = material[:4]
        self.dt = dt
        self.iter_fact = iter_fact
        self.device = device
        # integrator of run_simulation: 'adaptive' (default) = scipy's error-controlled RK45 exactly as the reference calls
        # it (spacecraft_model.py:516) -- what a drop-in user expects; 'fixed' = the fixed-step Dormand-Prince contract of
        # the ensemble hot path (step = t_eval spacing or max_step).  run_ensemble is always fixed-step.
        if integrator not in ("fixed", "adaptive"):
            raise ValueError("integrator must be 'adaptive' (scipy's error-controlled RK45 exactly as the reference calls it) "
                             "or 'fixed' (fixed-step Dormand-Prince, the ensemble hot path)")
        self.integrator = integrator

    # ------------------------------------------------------------------------- plumbing
    def _ctx(self) -> DeviceContext:
        return DeviceContext.get(self.device)

    def _dev(self, x, n=None):
        """numpy / scalar / tensor -> contiguous float64 CUDA tensor (broadcast to n if given)."""
        torch = _torch()
        ctx = self._ctx()
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x, dtype=np.float64))
        x = x.to(device=f"cuda:{ctx.device}", dtype=torch.float64)
        if n is not None:
            x = x.reshape(-1) if x.ndim else x.reshape(1)
            if x.numel() == 1 and n != 1:
                x = x.expand(n)
            if x.numel() != n:
                raise ValueError(f"expected {n} values, got {x.numel()}")
        return x.contiguous()

    # --------------------------------------------------------------- get_initial_state
    def get_initial_state(self, v, lat, lon, alt, azimuth, gamma, gmst=0.0):
        """spacecraft_model.py:412-435.  Scalars -> float64[6] (numpy, like the reference);
        arrays of length N -> CUDA tensor [N, 6]."""
        scalar = all(np.ndim(a) == 0 for a in (v, lat, lon, alt, azimuth, gamma))
        y0 = self.get_initial_states(v, lat, lon, alt, azimuth, gamma, gmst)
        return y0[0].cpu().numpy() if scalar else y0

    def get_initial_states(self, v, lat, lon, alt, azimuth, gamma, gmst=0.0):
        torch = _torch()
        ctx = self._ctx()
        n = max(int(np.size(a)) if not hasattr(a, "numel") else int(a.numel()) for a in (v, lat, lon, alt, azimuth, gamma))
        args = [self._dev(a, n) for a in (v, lat, lon, alt, azimuth, gamma)]
        y0 = torch.empty((n, 6), dtype=torch.float64, device=args[0].device)
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        ctx.check(ctx.lib.rb_initial_states(ctx.handle, n, *[a.data_ptr() for a in args], float(gmst),
                                            y0.data_ptr(), 6, 1, stream), "rb_initial_states")
        return y0

    def get_dispersed_initial_states(self, n, seed, nominal, sigma, first_index=0, gmst=0.0, return_params=False):
        """Monte-Carlo initial states generated on the device (counter-based Philox4x32-10; the
        state of trajectory first_index + i is independent of sharding).  nominal / sigma are the six
        launch-geometry values (v, lat, lon, alt, azimuth, gamma) and their 1-sigma dispersions."""
        torch = _torch()
        ctx = self._ctx()
        dev = f"cuda:{ctx.device}"
        y0 = torch.empty((n, 6), dtype=torch.float64, device=dev)
        params = torch.empty((6, n), dtype=torch.float64, device=dev) if return_params else None
        nom = (C.c_double * 6)(*[float(x) for x in nominal])
        sig = (C.c_double * 6)(*[float(x) for x in sigma])
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        ctx.check(ctx.lib.rb_dispersed_initial_states(ctx.handle, int(n), int(first_index), int(seed), C.addressof(nom),
                                                      C.addressof(sig), float(gmst),
                                                      None if params is None else params.data_ptr(), y0.data_ptr(), 6, 1,
                                                      stream), "rb_dispersed_initial_states")
        return (y0, params.t()) if return_params else y0

    # ------------------------------------------------------------- equations_of_motion
    def equations_of_motion_batch(self, t, y, Cd=None, A=None, m=None):
        """Device evaluation of the RHS dict for N (t, y) pairs.  t [N]; y [N,6] -> dict of CUDA tensors."""
        torch = _torch()
        ctx = self._ctx()
        y = self._dev(y)
        y = y.reshape(-1, 6)
        n = y.shape[0]
        t = self._dev(t, n)
        ysoa = y.t().contiguous()
        out = torch.empty((L.RB_DIAG_FIELDS + L.RB_THERMAL_FIELDS, n), dtype=torch.float64, device=y.device)
        arr = [None if a is None else self._dev(a, n) for a in (Cd, A, m)]
        ptr = [0 if a is None else a.data_ptr() for a in arr]
        th = L.RbThermal(capsule_length=self.height, dt=float(self.dt), iter_fact=float(self.iter_fact),
                         thermal_conductivity=float(self.thermal_conductivity),
                         specific_heat_capacity=float(self.specific_heat_capacity), emissivity=float(self.emissivity),
                         ablation_efficiency=float(self.ablation_efficiency))
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        ctx.check(ctx.lib.rb_equations_of_motion(ctx.handle, n, t.data_ptr(), ysoa.data_ptr(), self.epoch, self.gmst0,
                                                 ptr[0], ptr[1], ptr[2], float(self.Cd), float(self.A), float(self.m),
                                                 C.byref(th), out.data_ptr(), stream), "rb_equations_of_motion")
        o = out
        # thermal rows are spacecraft_temperature's RETURN order (Qc, Qr, Q_net, Q, T_s, dT); the reference
        # unpacks them as q_gen, q_c, q_r, q_net, T_s, dT (spacecraft_model.py:468) and labels the dict from
        # those names (:481-486) -- the permutation is reproduced so that the keys hold what the app gets
        return {
            "velocity": y[:, 3:6],
            "acceleration": o[0:3].t(), "gravitational_acceleration": o[3:6].t(), "J2_acceleration": o[6:9].t(),
            "moon_acceleration": o[9:12].t(), "sun_acceleration": o[12:15].t(), "drag_acceleration": o[15:18].t(),
            "altitude": o[18],
            "spacecraft_temperature": o[27], "spacecraft_heat_flux": o[26], "spacecraft_heat_flux_conduction": o[24],
            "spacecraft_heat_flux_radiation": o[25], "spacecraft_heat_flux_total": o[23],
            "spacecraft_temperature_change": o[28],
            "latitude": o[19], "density": o[20], "temperature": o[21], "relative_speed": o[22],
        }

    def equations_of_motion(self, t, y):
        """spacecraft_model.py:437-487 for one state: dict of numpy values (thermal entries are
        outside this path)."""
        d = self.equations_of_motion_batch(np.array([float(t)]), np.asarray(y, dtype=np.float64).reshape(1, 6))
        out = {}
        for k, v_ in d.items():
            a = v_.cpu().numpy()
            out[k] = a[0].copy() if a.ndim == 2 else float(a[0])
        return out

    # ------------------------------------------------------------------- run_ensemble
    def run_ensemble(self, y0, t0=0.0, dt=1.0, n_steps=0, out_stride=1, Cd=None, A=None, m=None,
                     store_samples=True, steps_per_launch=0, compact=True, early_exit=True, refine=True,
                     event_altitude=1000.0, table_threads=0, profile=False, skip_negligible_drag=False, overlap=True,
                     samples="dense", packed_capacity=None) -> EnsembleResult:
        """Propagate N independent initial states y0[N,6] for up to n_steps fixed steps of dt.

        Per-trajectory Cd / A / m may be given as arrays [N]; otherwise the model's scalars apply.
        Samples are stored every `out_stride` steps: samples="dense" as [n_out, 8, N] (SoA, coalesced, NaN after a
        trajectory's event); samples="packed" as slabs only as wide as the live ensemble (`result.packed`,
        PackedSamples) -- what a run "until impact" with a generous step cap should use: nothing is allocated,
        filled or written for the part of the cap the ensemble never flies.  packed_capacity (doubles of the arena)
        defaults to the dense size, the worst case; pass what the run needs when the cap is generous.
        """
        torch = _torch()
        ctx = self._ctx()
        if self.sim_type != "RK45":
            raise NotImplementedError("only the RK45 (Dormand-Prince) tableau is implemented on this path")
        if not (dt > 0) or n_steps < 0 or out_stride < 1:
            raise ValueError("need dt > 0, n_steps >= 0, out_stride >= 1")
        if not (isinstance(y0, torch.Tensor) and y0.is_cuda and y0.dtype == torch.float64 and y0.ndim == 2
                and y0.device.index == ctx.device):
            y0 = self._dev(y0)   # strided CUDA tensors (e.g. SoA views) are passed as they are
        if y0.ndim == 1:
            y0 = y0.reshape(1, 6)
        if y0.ndim != 2 or y0.shape[1] != 6:
            raise ValueError("y0 must be [N, 6]")
        n = y0.shape[0]
        dev = y0.device
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        ctx.stage_time_table(self.epoch, self.gmst0, float(t0), float(dt), int(n_steps), stream, table_threads)
        n_out = n_steps // out_stride + 1
        if samples not in ("dense", "packed"):
            raise ValueError("samples must be 'dense' or 'packed'")
        packed = store_samples and samples == "packed"
        samples = torch.full((n_out, L.RB_SAMPLE_FIELDS, n), float("nan"), dtype=torch.float64, device=dev) \
            if store_samples and not packed else None
        arena = slabs = slab_live = None
        n_slabs = C.c_int64(0)
        if packed:
            spl = int(steps_per_launch) if steps_per_launch > 0 else 32
            slab_cap = 2 * (n_steps // spl + 2) + 1
            cap = int(packed_capacity) if packed_capacity else n_out * L.RB_SAMPLE_FIELDS * ((n + 15) // 16 * 16 + 16)
            arena = torch.empty(cap, dtype=torch.float64, device=dev)
            slabs = (L.RbSampleSlab * slab_cap)()
            slab_live = torch.zeros(slab_cap, dtype=torch.int32, device=dev)
        final_state = torch.empty((6, n), dtype=torch.float64, device=dev)
        impact_step = torch.empty(n, dtype=torch.int32, device=dev)
        impact_time = torch.empty(n, dtype=torch.float64, device=dev)
        status = torch.empty(n, dtype=torch.uint8, device=dev)
        n_samples = torch.empty(n, dtype=torch.int32, device=dev)
        arr = [None if a is None else self._dev(a, n) for a in (Cd, A, m)]
        rin = L.RbIn(n=n, y0=y0.data_ptr(), y0_traj_stride=y0.stride(0), y0_comp_stride=y0.stride(1),
                     cd=None if arr[0] is None else arr[0].data_ptr(),
                     area=None if arr[1] is None else arr[1].data_ptr(),
                     mass=None if arr[2] is None else arr[2].data_ptr(),
                     cd_scalar=float(self.Cd), area_scalar=float(self.A), mass_scalar=float(self.m))
        flags = (L.RB_FLAG_COMPACT if compact else 0) | (L.RB_FLAG_EARLY_EXIT if early_exit else 0) | \
                (0 if refine else L.RB_FLAG_NO_REFINE) | (L.RB_FLAG_PROFILE if profile else 0) | \
                (L.RB_FLAG_SKIP_NEGLIGIBLE_DRAG if skip_negligible_drag else 0) | (0 if overlap else L.RB_FLAG_NO_OVERLAP) | \
                (L.RB_FLAG_PACKED_SAMPLES if packed else 0)
        run = L.RbRun(first_step=0, n_steps=int(n_steps), out_stride=int(out_stride),
                      steps_per_launch=int(steps_per_launch), event_altitude=float(event_altitude), flags=flags)
        out = L.RbOut(samples=(arena.data_ptr() if packed else None if samples is None else samples.data_ptr()),
                      sample_stride=L.RB_SAMPLE_FIELDS * n, field_stride=n, n_out_capacity=n_out,
                      final_state=final_state.data_ptr(), impact_step=impact_step.data_ptr(),
                      impact_time=impact_time.data_ptr(), status=status.data_ptr(), n_samples=n_samples.data_ptr(),
                      packed_capacity=(arena.numel() if packed else 0), slabs=slabs if packed else None,
                      slab_capacity=(len(slabs) if packed else 0), n_slabs=C.pointer(n_slabs) if packed else None,
                      slab_live=slab_live.data_ptr() if packed else None)
        if n > 0:
            ctx.check(ctx.lib.rb_propagate(ctx.handle, C.byref(rin), C.byref(run), C.byref(out), stream), "rb_propagate")
        t = float(t0) + float(dt) * out_stride * np.arange(n_out)
        pk = None
        if packed:
            pk = PackedSamples(arena, PackedSamples.slabs_from_ctypes(slabs, n_slabs.value), impact_step, status, n_samples, n, n_out)
            pk.slab_live = slab_live   # device [n_slabs]: live count of every slab (n_live = -1 in the host descriptors)
        return EnsembleResult(t=t, samples=samples, final_state=final_state, impact_step=impact_step,
                              impact_time=impact_time, status=status, n_samples=n_samples, dt=float(dt),
                              out_stride=int(out_stride), n_steps=int(n_steps), stats=ctx.stats(), packed=pk)

    # ------------------------------------------------------------- host-buffer path
    def host_buffers(self, n, n_steps, out_stride, packed_capacity, steps_per_launch=0):
        """Pinned host buffers for run_ensemble_host (allocate once, reuse for every batch of the same shape)."""
        torch = _torch()
        spl = int(steps_per_launch) if steps_per_launch > 0 else 32
        b = dict(n=int(n), n_steps=int(n_steps), out_stride=int(out_stride), steps_per_launch=int(steps_per_launch))
        b["y0"] = torch.empty((n, 6), dtype=torch.float64).pin_memory()
        b["packed"] = torch.empty(int(packed_capacity), dtype=torch.float64).pin_memory() if packed_capacity else None
        b["slabs"] = (L.RbSampleSlab * (2 * (n_steps // spl + 2) + 1))()
        b["final_state"] = torch.empty((n, 6), dtype=torch.float64).pin_memory()
        b["impact_step"] = torch.empty(n, dtype=torch.int32).pin_memory()
        b["impact_time"] = torch.empty(n, dtype=torch.float64).pin_memory()
        b["status"] = torch.empty(n, dtype=torch.uint8).pin_memory()
        b["n_samples"] = torch.empty(n, dtype=torch.int32).pin_memory()
        return b

    def run_ensemble_host(self, buffers, t0=0.0, dt=1.0, event_altitude=1000.0) -> EnsembleResult:
        """End to end with HOST memory: buffers["y0"] [N,6] in, everything else out (host tensors in `buffers`).  One
        call = upload of the initial states, time table built chunk-wise on the host while the GPU runs, propagation
        until impact (cap n_steps), every packed sample slab sent by ONE copy-engine transfer as soon as its launch
        finished, summaries; returns when all of it is in host memory.  Samples come back packed (result.packed, host
        arena); buffers["packed"] = None propagates without samples."""
        ctx = self._ctx()
        if self.sim_type != "RK45":
            raise NotImplementedError("only the RK45 (Dormand-Prince) tableau is implemented on this path")
        b = buffers
        n, n_steps, out_stride = b["n"], b["n_steps"], b["out_stride"]
        run = L.RbRun(first_step=0, n_steps=n_steps, out_stride=out_stride, steps_per_launch=b["steps_per_launch"],
                      event_altitude=float(event_altitude), flags=0)
        n_slabs, used = C.c_int64(0), C.c_int64(0)
        pk = b["packed"]
        rc = ctx.lib.rb_propagate_host_packed(
            ctx.handle, n, b["y0"].data_ptr(), None, None, None, float(self.Cd), float(self.A), float(self.m), self.epoch,
            self.gmst0, float(t0), float(dt), C.byref(run), pk.data_ptr() if pk is not None else None,
            pk.numel() if pk is not None else 0, b["slabs"], len(b["slabs"]), C.byref(n_slabs), C.byref(used),
            b["final_state"].data_ptr(), b["impact_step"].data_ptr(), b["impact_time"].data_ptr(), b["status"].data_ptr(),
            b["n_samples"].data_ptr())
        ctx.check(rc, "rb_propagate_host_packed")
        n_out = n_steps // out_stride + 1
        packed = None
        if pk is not None:
            packed = PackedSamples(pk, PackedSamples.slabs_from_ctypes(b["slabs"], n_slabs.value), b["impact_step"], b["status"],
                                   b["n_samples"], n, n_out)
        return EnsembleResult(t=float(t0) + float(dt) * out_stride * np.arange(n_out), samples=None,
                              final_state=b["final_state"].t(), impact_step=b["impact_step"], impact_time=b["impact_time"],
                              status=b["status"], n_samples=b["n_samples"], dt=float(dt), out_stride=out_stride,
                              n_steps=n_steps, stats=ctx.stats(), packed=packed)

    # ------------------------------------------------------------- ground track (N2)
    def ground_track(self, res: EnsembleResult, threshold=100000.0, t0=0.0):
        """The app's post-pass for every trajectory of an ensemble (app.py:225-243, :302-309): dict of
        CUDA tensors geo [n_out, 3, N] (lat deg, lon deg, h), downrange [n_out, N], crossing_time [N],
        crossing_downrange [N], n_crossings [N] (NaN after the event / when there is no crossing)."""
        torch = _torch()
        ctx = self._ctx()
        n_out, _, n = res.samples.shape
        dev = res.samples.device
        geo = torch.full((n_out, 3, n), float("nan"), dtype=torch.float64, device=dev)
        dr = torch.full((n_out, n), float("nan"), dtype=torch.float64, device=dev)
        ct = torch.empty(n, dtype=torch.float64, device=dev)
        cd = torch.empty(n, dtype=torch.float64, device=dev)
        nc = torch.empty(n, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        ctx.check(ctx.lib.rb_ground_track(ctx.handle, n, n_out, res.samples.data_ptr(), L.RB_SAMPLE_FIELDS * n, n,
                                          res.n_samples.data_ptr(), float(t0), res.dt * res.out_stride, self.gmst0,
                                          float(threshold), geo.data_ptr(), dr.data_ptr(), ct.data_ptr(), cd.data_ptr(),
                                          nc.data_ptr(), stream), "rb_ground_track")
        return dict(geo=geo, downrange=dr, crossing_time=ct, crossing_downrange=cd, n_crossings=nc)

    def impact_points(self, res: EnsembleResult, t0=0.0):
        """Geodetic impact (or end) points [3, N] = lat deg, lon deg, h of an ensemble."""
        torch = _torch()
        ctx = self._ctx()
        n = res.status.shape[0]
        out = torch.empty((3, n), dtype=torch.float64, device=res.status.device)
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        ctx.check(ctx.lib.rb_impact_points(ctx.handle, n, res.final_state.data_ptr(), res.impact_time.data_ptr(),
                                           res.status.data_ptr(), float(t0) + res.dt * res.n_steps, self.gmst0,
                                           out.data_ptr(), stream), "rb_impact_points")
        return out

    # ------------------------------------------------------------ checkpoint / resume
    def run_ensemble_chunked(self, y0, dt, n_steps, chunk_steps, t0=0.0, out_stride=1, Cd=None, A=None, m=None,
                             store_samples=True, event_altitude=1000.0):
        """Generator: propagate in chunks of `chunk_steps` (a multiple of out_stride) and yield the SAME
        EnsembleResult after every chunk -- its tensors are the checkpoint (state, status, impact step / time,
        samples filled so far), so a long run can be streamed, inspected or stopped and continued.  The
        concatenation of the chunks equals run_ensemble(n_steps) bit for bit."""
        torch = _torch()
        ctx = self._ctx()
        if chunk_steps < 1 or chunk_steps % out_stride:
            raise ValueError("chunk_steps must be a positive multiple of out_stride")
        y0 = self._dev(y0) if not (isinstance(y0, torch.Tensor) and y0.is_cuda and y0.dtype == torch.float64) else y0
        y0 = y0.reshape(-1, 6)
        n, dev = y0.shape[0], y0.device
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        ctx.stage_time_table(self.epoch, self.gmst0, float(t0), float(dt), int(n_steps), stream)
        n_out = n_steps // out_stride + 1
        samples = torch.full((n_out, L.RB_SAMPLE_FIELDS, n), float("nan"), dtype=torch.float64, device=dev) \
            if store_samples else None
        res = EnsembleResult(t=float(t0) + float(dt) * out_stride * np.arange(n_out), samples=samples,
                             final_state=torch.empty((6, n), dtype=torch.float64, device=dev),
                             impact_step=torch.empty(n, dtype=torch.int32, device=dev),
                             impact_time=torch.empty(n, dtype=torch.float64, device=dev),
                             status=torch.empty(n, dtype=torch.uint8, device=dev),
                             n_samples=torch.empty(n, dtype=torch.int32, device=dev), dt=float(dt), out_stride=int(out_stride),
                             n_steps=0)
        arr = [None if a is None else self._dev(a, n) for a in (Cd, A, m)]
        rin = L.RbIn(n=n, y0=y0.data_ptr(), y0_traj_stride=y0.stride(0), y0_comp_stride=y0.stride(1),
                     cd=None if arr[0] is None else arr[0].data_ptr(), area=None if arr[1] is None else arr[1].data_ptr(),
                     mass=None if arr[2] is None else arr[2].data_ptr(),
                     cd_scalar=float(self.Cd), area_scalar=float(self.A), mass_scalar=float(self.m))
        out = L.RbOut(samples=None if samples is None else samples.data_ptr(), sample_stride=L.RB_SAMPLE_FIELDS * n,
                      field_stride=n, n_out_capacity=n_out, final_state=res.final_state.data_ptr(),
                      impact_step=res.impact_step.data_ptr(), impact_time=res.impact_time.data_ptr(),
                      status=res.status.data_ptr(), n_samples=res.n_samples.data_ptr())
        done = 0
        while done < n_steps:
            steps = min(chunk_steps, n_steps - done)
            flags = L.RB_FLAG_COMPACT | L.RB_FLAG_EARLY_EXIT | (L.RB_FLAG_RESUME if done else 0)
            run = L.RbRun(first_step=done, n_steps=steps, out_stride=int(out_stride), steps_per_launch=0,
                          event_altitude=float(event_altitude), flags=flags)
            if n > 0:
                ctx.check(ctx.lib.rb_propagate(ctx.handle, C.byref(rin), C.byref(run), C.byref(out), stream), "rb_propagate")
            done += steps
            res.n_steps = done
            res.stats = ctx.stats()
            yield res

    # ------------------------------------------------------------ adaptive mode (N4)
    def run_ensemble_adaptive(self, y0, t_span, t_eval_dt, n_eval=None, rtol=1e-8, atol=1e-10, Cd=None, A=None, m=None,
                              store_samples=True, max_steps=1_000_000, event_altitude=1000.0, ephemeris_dt=0.0,
                              step_log_capacity=0) -> EnsembleResult:
        """The reference's integrator as shipped (solve_ivp RK45, rtol 1e-8, atol 1e-10, spacecraft_model.py:516)
        for N independent initial states, each on its own adaptive time axis.  Samples are the dense-output
        values at t_span[0] + k * t_eval_dt (k < n_eval; default: np.arange(t0, tf, dt), app.py:171) that
        are not after the terminal altitude event.  ephemeris_dt > 0 takes Sun / Moon / solar factor from a grid of that
        spacing (device-built, 4-point Lagrange; 16 s changes positions by < 1e-10 m) instead of evaluating the
        reference's ephemeris functions at every stage time of every trajectory.  step_log_capacity > 0 also returns
        `res.step_log` [capacity, 7, n]: start time and state of every accepted step (what solve_ivp reports as the roots of
        run_simulation's always-zero progress_event, spacecraft_model.py:505-511)."""
        torch = _torch()
        ctx = self._ctx()
        if self.sim_type not in ("RK45", "RK23", "DOP853"):
            raise NotImplementedError("sim_type must be 'RK45', 'RK23' or 'DOP853' (scipy's implicit Radau / BDF / LSODA are not built)")
        t0, tf = float(t_span[0]), float(t_span[1])
        if n_eval is None:
            n_eval = len(np.arange(t0, tf, t_eval_dt))
        if not (isinstance(y0, torch.Tensor) and y0.is_cuda and y0.dtype == torch.float64 and y0.ndim == 2):
            y0 = self._dev(y0)
        if y0.ndim == 1:
            y0 = y0.reshape(1, 6)
        n = y0.shape[0]
        dev = y0.device
        samples = torch.full((n_eval, L.RB_SAMPLE_FIELDS, n), float("nan"), dtype=torch.float64, device=dev) \
            if store_samples else None
        final_state = torch.empty((6, n), dtype=torch.float64, device=dev)
        impact_time = torch.empty(n, dtype=torch.float64, device=dev)
        status = torch.empty(n, dtype=torch.uint8, device=dev)
        n_samples = torch.empty(n, dtype=torch.int32, device=dev)
        n_acc = torch.empty(n, dtype=torch.int32, device=dev)
        n_rej = torch.empty(n, dtype=torch.int32, device=dev)
        step_log = torch.full((int(step_log_capacity), 7, n), float("nan"), dtype=torch.float64, device=dev) \
            if step_log_capacity > 0 else None
        arr = [None if a is None else self._dev(a, n) for a in (Cd, A, m)]
        rin = L.RbIn(n=n, y0=y0.data_ptr(), y0_traj_stride=y0.stride(0), y0_comp_stride=y0.stride(1),
                     cd=None if arr[0] is None else arr[0].data_ptr(), area=None if arr[1] is None else arr[1].data_ptr(),
                     mass=None if arr[2] is None else arr[2].data_ptr(),
                     cd_scalar=float(self.Cd), area_scalar=float(self.A), mass_scalar=float(self.m))
        run = L.RbAdaptiveRun(t0=t0, t_bound=tf, rtol=float(rtol), atol=float(atol), t_eval0=t0, t_eval_dt=float(t_eval_dt),
                              n_eval=int(n_eval), max_steps=int(max_steps), event_altitude=float(event_altitude),
                              epoch_jd=self.epoch, gmst0=self.gmst0, ephemeris_dt=float(ephemeris_dt),
                              method={"RK45": L.RB_METHOD_RK45, "RK23": L.RB_METHOD_RK23, "DOP853": L.RB_METHOD_DOP853}[self.sim_type])
        out = L.RbAdaptiveOut(samples=None if samples is None else samples.data_ptr(), sample_stride=L.RB_SAMPLE_FIELDS * n,
                              field_stride=n, final_state=final_state.data_ptr(), impact_time=impact_time.data_ptr(),
                              status=status.data_ptr(), n_samples=n_samples.data_ptr(), n_accepted=n_acc.data_ptr(),
                              n_rejected=n_rej.data_ptr(), step_log=None if step_log is None else step_log.data_ptr(),
                              step_capacity=int(step_log_capacity) if step_log is not None else 0)
        stream = torch.cuda.current_stream(ctx.device).cuda_stream
        if n > 0:
            ctx.check(ctx.lib.rb_propagate_adaptive(ctx.handle, C.byref(rin), C.byref(run), C.byref(out), stream),
                      "rb_propagate_adaptive")
        res = EnsembleResult(t=t0 + float(t_eval_dt) * np.arange(n_eval), samples=samples, final_state=final_state,
                             impact_step=n_acc, impact_time=impact_time, status=status, n_samples=n_samples,
                             dt=float(t_eval_dt), out_stride=1, n_steps=int(n_eval) - 1,
                             stats=dict(kernel_launches=1, propagate_launches=1, steps_enqueued=0, last_polled_live=-1,
                                        propagate_ms=0.0))
        res.n_accepted, res.n_rejected = n_acc, n_rej
        res.step_log = step_log
        return res

    # ------------------------------------------------------------------ run_simulation
    def run_simulation(self, t_span, y0, t_eval, progress_callback=None, max_step=None, integrator=None):
        """spacecraft_model.py:490-519 for ONE trajectory.  With the model's default integrator ('adaptive') this is the
        reference's own adaptive RK45 run; integrator='fixed' (argument or constructor) selects the fixed-step contract:

        `t_eval` must be a uniform grid starting at t_span[0] (the app builds
        `np.arange(ts, tf, dt)`, app.py:171); the integration step is its spacing, or `max_step`
        if given (the spacing must then be an integer multiple of it).  Returns an
        OdeResult-like object: t [n], y [6, n] (grid samples not after the event, scipy t_eval
        semantics), t_events = [array([impact time]) or empty, empty], y_events, status,
        additional_data (dict of [n] / [n, 3] arrays, keys of :472-480).
        """
        t_eval = np.asarray(t_eval, dtype=np.float64)
        t0, tf = float(t_span[0]), float(t_span[1])
        if t_eval.ndim != 1 or t_eval.size < 1:
            raise ValueError("t_eval must be a non-empty 1-D grid")
        if (integrator or self.integrator) == "adaptive":
            return self._run_simulation_adaptive(t0, tf, y0, t_eval, progress_callback)
        if t_eval.size > 1:
            spacing = float(t_eval[1] - t_eval[0])
            if not np.allclose(np.diff(t_eval), spacing, rtol=1e-9, atol=1e-12) or spacing <= 0:
                raise ValueError("this path integrates at a fixed step: t_eval must be uniformly spaced")
        else:
            spacing = tf - t0
        if abs(t_eval[0] - t0) > 1e-9 * max(1.0, abs(t0)):
            raise ValueError("t_eval must start at t_span[0]")
        h = float(max_step) if max_step else spacing
        stride = int(round(spacing / h))
        if stride < 1 or abs(stride * h - spacing) > 1e-9 * spacing:
            raise ValueError("t_eval spacing must be an integer multiple of max_step")
        n_steps = int(np.floor((tf - t0) / h + 1e-9))
        res = self.run_ensemble(np.asarray(y0, dtype=np.float64).reshape(1, 6), t0=t0, dt=h, n_steps=n_steps,
                                out_stride=stride, steps_per_launch=0)
        n_valid = min(int(res.n_samples[0].item()), t_eval.size)
        samples = res.samples[:n_valid, :, 0].cpu().numpy()       # [n, 8]
        status = int(res.status[0].item())
        te = float(res.impact_time[0].item())
        sol = OdeResultLike()
        sol.t = t_eval[:n_valid].copy()
        sol.y = np.ascontiguousarray(samples[:, :6].T)
        hit = status == L.TRAJ_IMPACTED
        sol.t_events = [np.array([te]) if hit else np.array([]), np.array([])]
        sol.y_events = [res.final_state[:, 0].cpu().numpy()[None, :] if hit else np.empty((0, 6)), np.empty((0, 6))]
        sol.status = 1 if hit else (0 if status == L.TRAJ_ALIVE else -1)
        sol.success = status != L.TRAJ_NONFINITE
        sol.message = ("A termination event occurred." if hit else
                       "The solver successfully reached the end of the integration interval." if sol.success
                       else "Non-finite state encountered.")
        steps_done = (int(res.impact_step[0].item()) if status != L.TRAJ_ALIVE else n_steps)
        sol.nfev = 6 * steps_done + 1
        sol.njev = 0
        sol.nlu = 0
        sol.sol = None
        # additional_data: re-evaluate the RHS at every output sample, as the reference does (:517-518)
        d = self.equations_of_motion_batch(sol.t, samples[:, :6]) if n_valid else {}
        sol.additional_data = {k: v_.cpu().numpy().copy() for k, v_ in d.items()}
        if progress_callback is not None:
            progress_callback(1.0, time.time() - self.start_time)
        return sol

    def _run_simulation_adaptive(self, t0, tf, y0, t_eval, progress_callback):
        """run_simulation with the reference's own integrator (adaptive RK45, rtol 1e-8, atol 1e-10)."""
        if t_eval.size > 1:
            spacing = float(t_eval[1] - t_eval[0])
            if not np.allclose(np.diff(t_eval), spacing, rtol=1e-9, atol=1e-12) or spacing <= 0:
                raise ValueError("t_eval must be uniformly spaced (np.arange(ts, tf, dt), app.py:171)")
        else:
            spacing = max(tf - t0, 1.0)
        if abs(t_eval[0] - t0) > 1e-9 * max(1.0, abs(t0)):
            raise ValueError("t_eval must start at t_span[0]")
        res = self.run_ensemble_adaptive(np.asarray(y0, dtype=np.float64).reshape(1, 6), (t0, tf), spacing, n_eval=t_eval.size,
                                         step_log_capacity=1 << 17)
        n_valid = int(res.n_samples[0].item())
        samples = res.samples[:n_valid, :, 0].cpu().numpy()
        status = int(res.status[0].item())
        hit = status == L.TRAJ_IMPACTED
        sol = OdeResultLike()
        sol.t = t_eval[:n_valid].copy()
        sol.y = np.ascontiguousarray(samples[:, :6].T)
        # progress_event (spacecraft_model.py:505-511) is identically zero: solve_ivp reports the START of every accepted step
        # as its root (ivp.py:677-699, brentq returns the bracket's left end when f is zero there)
        n_log = min(int(res.n_accepted[0].item()), res.step_log.shape[0])
        log = res.step_log[:n_log, :, 0].cpu().numpy()
        sol.t_events = [np.array([float(res.impact_time[0].item())]) if hit else np.array([]), log[:, 0].copy()]
        sol.y_events = [res.final_state[:, 0].cpu().numpy()[None, :] if hit else np.empty((0, 6)), log[:, 1:].copy()]
        if progress_callback is not None and tf > t0:
            for ts in log[:, 0]:   # the reference reports progress from inside progress_event, once per step
                progress_callback((ts - t0) / (tf - t0), time.time() - self.start_time)
        sol.status = 1 if hit else (0 if status == L.TRAJ_ALIVE else -1)
        sol.success = status in (L.TRAJ_ALIVE, L.TRAJ_IMPACTED)
        sol.message = ("A termination event occurred." if hit else
                       "The solver successfully reached the end of the integration interval." if sol.success
                       else "Required step size is less than spacing between numbers.")
        n_acc, n_rej = int(res.n_accepted[0].item()), int(res.n_rejected[0].item())
        if self.sim_type == "DOP853":   # 12 evaluations per attempt + 3 for the dense output of every accepted step
            sol.nfev = 2 + 12 * (n_acc + n_rej) + 3 * n_acc
        else:
            sol.nfev = 2 + (3 if self.sim_type == "RK23" else 6) * (n_acc + n_rej)
        sol.njev = 0
        sol.nlu = 0
        sol.sol = None
        d = self.equations_of_motion_batch(sol.t, samples[:, :6]) if n_valid else {}
        sol.additional_data = {k: v_.cpu().numpy().copy() for k, v_ in d.items()}
        if progress_callback is not None:
            progress_callback(1.0, time.time() - self.start_time)
        return sol
