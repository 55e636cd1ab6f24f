/* reentry_b200.h -- C ABI of the B200-native ensemble trajectory propagator.
 *
 * Drop-in boundary for the ONE hot path of JMMonte/Reentry-OLD: the trajectory
 * integration behind SpacecraftModel.run_simulation (spacecraft_model.py:490-519) and the
 * acceleration evaluation it drives (spacecraft_model.py:437-470), run for ensembles of
 * independent initial states.  The reference has no FFI of its own (it is pure Python +
 * numba); each entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only: no torch / numpy / C++ types cross this boundary.
 *
 * Conventions
 *   - every function returns int32 status: 0 = RB_OK, negative = rb_error (rb_strerror()).
 *   - device pointers are plain CUDA device addresses on the context's device; the CALLER
 *     allocates and frees every buffer it passes; the library owns only its context
 *     (time table, live-index workspace, counters).
 *   - work is enqueued on the cudaStream_t passed as `void* stream` (NULL = legacy default
 *     stream).  Nothing synchronises unless stated (rb_ctx_destroy, RB_FLAG_EARLY_EXIT
 *     polling, the *_host convenience calls).
 *   - a context is bound to one device, is not re-entrant, one per rank/GPU.
 *   - per-trajectory numerical failures are reported in status[i], never as a process abort.
 *   - all floating point is IEEE float64; time is seconds from the model epoch.
 *
 * Fixed-step contract: Dormand-Prince 5(4) with scipy's RK45 tableau
 * (scipy/integrate/_ivp/rk.py:541-565), FSAL, t_{n+1} = t_n + dt, stage times t_n + C_s*dt;
 * event g = |r| - 6378137 - event_altitude, active in the step where g_old >= 0 and
 * g_new <= 0 (scipy ivp.py:151, direction -1); impact time = root of g on the step's
 * quartic dense output (rk.py:554-565, 723-737).
 */
#ifndef REENTRY_B200_H
#define REENTRY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_ABI_VERSION 3
#define RB_TABLE_ENTRY_DOUBLES 16 /* one time-table entry = 128 B */
#define RB_STAGES_PER_STEP 5      /* distinct new stage times per DP5 step (c = 1/5, 3/10, 4/5, 8/9, 1) */
#define RB_SAMPLE_FIELDS 8        /* x, y, z, vx, vy, vz, altitude, speed */

typedef enum rb_error {
    RB_OK = 0,
    RB_ERR_INVALID_ARG = -1,
    RB_ERR_CUDA = -2,       /* rb_ctx_last_error() has the CUDA text */
    RB_ERR_NO_TIME_TABLE = -3,
    RB_ERR_TABLE_RANGE = -4, /* run asks for steps the staged time table does not cover */
    RB_ERR_NOMEM = -5,
    RB_ERR_NO_DEVICE = -6,
    RB_ERR_SAMPLE_CAPACITY = -7 /* packed sample arena or slab list too small for the run */
} rb_error;

/* status[i] values */
typedef enum rb_traj_status {
    RB_TRAJ_ALIVE = 0,     /* reached n_steps without an event */
    RB_TRAJ_IMPACTED = 1,  /* terminal altitude event fired (spacecraft_model.py:496-504) */
    RB_TRAJ_NONFINITE = 2, /* state became NaN/Inf; lane retired */
    RB_TRAJ_STEP_FAILED = 3 /* adaptive mode only: step size below 10 ulp of t, or max_steps reached */
} rb_traj_status;

/* run flags */
#define RB_FLAG_COMPACT 1u      /* order-preserving compaction of live trajectories between launches */
#define RB_FLAG_EARLY_EXIT 2u   /* host polls the live count (2 launches behind) and stops when it is 0 */
#define RB_FLAG_NO_REFINE 4u    /* skip the dense-output impact-time refinement (impact_time = NaN) */
#define RB_FLAG_SKIP_NEGLIGIBLE_DRAG 16u /* above RB_NEGLIGIBLE_DRAG_ALTITUDE (400 km) the model's density
                                   (<= 4e-49 kg/m^3, spacecraft_model.py:84-89) makes rho*bc*v^2 < 1e-40 m/s^2 for any
                                   bc < 1e8 m^2/kg -- below one ulp of every other term -- so latitude, density and
                                   drag are not evaluated there.  Off by default; orbit ensembles run ~2.5x faster */
#define RB_NEGLIGIBLE_DRAG_ALTITUDE 400000.0
#define RB_FLAG_NO_OVERLAP 32u   /* do not split large ensembles (>= RB_OVERLAP_MIN_TRAJECTORIES) into two
                                   partitions running on two streams (the second is library-owned; both are
                                   ordered after prior work on `stream` and joined back before the call's last
                                   kernel, so the caller still sees ONE stream) */
#define RB_OVERLAP_MIN_TRAJECTORIES 600000
#define RB_FLAG_RESUME 256u       /* checkpoint / resume: continue a run whose previous chunk ended at table step
                                   `first_step` (a multiple of out_stride).  in->y0 is ignored; out->final_state,
                                   status, impact_step, impact_time (and samples, indexed by the absolute step) are the
                                   buffers the previous chunk wrote.  Chunked runs equal the one-shot run bit for bit */
#define RB_FLAG_DIRECT_SAMPLES 128u /* rb_propagate_host, pinned buffer: the step kernel stores EVERY sample row straight
                                      into host memory (default: dense rows travel by copy engine from a device staging
                                      buffer, only the sparse rows of the retiring phase are stored directly) */
#define RB_FLAG_STAGED_SAMPLES 64u /* rb_propagate_host: never write samples directly into pinned host memory */
#define RB_FLAG_PACKED_SAMPLES 512u /* samples are written as PACKED SLABS (see rb_sample_slab) instead of dense
                                      [n_out][8][n] rows: each launch's rows are only as wide as the live ensemble at
                                      that time, so a run "until impact" writes (and, on the host path, sends over
                                      PCIe) the valid samples and nothing else.  Needs RB_FLAG_COMPACT | RB_FLAG_EARLY_EXIT
                                      to narrow the rows; not combinable with RB_FLAG_RESUME */
#define RB_FLAG_PROFILE 8u      /* bracket every step-kernel launch with CUDA events; rb_propagate then
                                   synchronises the stream at the end and rb_stats.propagate_ms is valid */

typedef struct rb_ctx rb_ctx;

int32_t rb_abi_version(void);
const char *rb_strerror(int32_t code);
const char *rb_ctx_last_error(const rb_ctx *ctx);

/* Replaces SpacecraftModel.__init__ (spacecraft_model.py:393-410) for the path: binds a
 * device; the physical constants are compiled in from include/reentry_constants.h. */
int32_t rb_ctx_create(rb_ctx **out, int32_t device);
int32_t rb_ctx_destroy(rb_ctx *ctx); /* synchronises the device */

/* ---- time table: everything in the RHS that depends on time only -------------------
 * Entry e (16 doubles): sun[3], moon[3] (sun/moon_position_vector, spacecraft_model.py:
 * 270-344), k*s/|s|^3 for Sun and Moon (indirect term of third_body_acceleration, :362),
 * cos/sin(gmst0 + EARTH_OMEGA*t) (:443, coordinate_converter.py:32-33), the time part of
 * solar_activity_factor (:151-159), t.  Entry 0 is t0; step n owns entries 1+5n .. 5+5n
 * (stage times t_n + C_s*dt, s = 1..5; the FSAL evaluation at t_n + dt reuses entry 5+5n). */
int64_t rb_time_table_entries(int64_t n_steps); /* 5*n_steps + 1 */

/* Host builder (libm, optional OpenMP threads; n_threads <= 0 = all).  host_table must
 * hold rb_time_table_entries(n_steps) * 16 doubles. */
int32_t rb_build_time_table(double epoch_jd, double gmst0, double t0, double dt, int64_t n_steps,
                            double *host_table, int32_t n_threads);

/* Upload a host table (copied; the context owns the device copy and derives from it, on the device, the 96-byte
 * entries the step kernel streams: Sun, Moon, -(sum of the two indirect terms), solar factor, t, solar factor - 1). */
int32_t rb_ctx_set_time_table(rb_ctx *ctx, const double *host_table, int64_t n_steps, double t0,
                              double dt, void *stream);

/* ---- initial states: batched SpacecraftModel.get_initial_state (spacecraft_model.py:412-435)
 * Inputs are device arrays [n] (v m/s, lat deg, lon deg, alt m, azimuth deg, gamma deg);
 * y0 is written SoA: y0[c * y0_comp_stride + i * y0_traj_stride]. */
int32_t rb_initial_states(rb_ctx *ctx, int64_t n, const double *v, const double *lat, const double *lon,
                          const double *alt, const double *azimuth, const double *gamma, double gmst,
                          double *y0, int64_t y0_traj_stride, int64_t y0_comp_stride, void *stream);

/* ---- Monte-Carlo dispersions generated on the device (new, the reference
 * propagates one trajectory): Philox4x32-10 keyed by `seed`, counter = (first_index + i, draw),
 * Box-Muller normals z[6]; (v, lat, lon, alt, azimuth, gamma)[c] = nominal6[c] + sigma6[c]*z[c], then
 * get_initial_state.  The result for trajectory first_index + i does not depend on how the ensemble
 * is sharded.  nominal6 / sigma6 are HOST arrays; params (device SoA [6][n]) may be NULL. */
int32_t rb_dispersed_initial_states(rb_ctx *ctx, int64_t n, int64_t first_index, uint64_t seed,
                                    const double *nominal6, const double *sigma6, double gmst, double *params,
                                    double *y0, int64_t y0_traj_stride, int64_t y0_comp_stride, void *stream);

/* ---- propagation ------------------------------------------------------------------- */
typedef struct rb_in {
    int64_t n;               /* trajectories */
    const double *y0;        /* device; element (i, c) at y0[i*y0_traj_stride + c*y0_comp_stride] */
    int64_t y0_traj_stride;  /* AoS [n][6]: 6, SoA [6][n]: 1 */
    int64_t y0_comp_stride;  /* AoS: 1, SoA: n */
    const double *cd;        /* device [n] or NULL -> cd_scalar  (SpacecraftModel.Cd) */
    const double *area;      /* device [n] or NULL -> area_scalar (SpacecraftModel.A) */
    const double *mass;      /* device [n] or NULL -> mass_scalar (SpacecraftModel.m) */
    double cd_scalar, area_scalar, mass_scalar;
} rb_in;

typedef struct rb_run {
    int64_t first_step;       /* index of the first step within the staged time table (normally 0) */
    int64_t n_steps;          /* steps to take (cap for "until impact") */
    int64_t out_stride;       /* store every out_stride-th step (>= 1); sample k = step k*out_stride */
    int64_t steps_per_launch; /* compaction period in steps; 0 = library default (32) */
    double event_altitude;    /* 1000.0 in the reference (spacecraft_model.py:501) */
    uint32_t flags;           /* RB_FLAG_* */
    uint32_t reserved;
} rb_run;

/* One packed slab = the sample rows ONE launch of the step kernel wrote for ONE partition of the ensemble
 * (trajectories [first_traj, first_traj + n_traj)): rows first_row .. first_row + n_rows - 1, each row
 * [RB_SAMPLE_FIELDS][width] doubles at arena[offset + (k - first_row)*8*width + f*width + j].  Column j is the
 * j-th trajectory, in ascending index order, of the partition that was still alive when the launch started, i.e.
 * whose impact_step is -1 or > first_step (order-preserving compaction keeps the live list sorted) -- the column
 * map is a function of the per-trajectory summaries, no index array travels.  (n_live == n_traj: column j is
 * trajectory first_traj + j; that is also the case of a run without RB_FLAG_COMPACT, whose list never shrinks.)
 * Columns >= n_live are padding (the host's view of the live count lags the device by one launch); entries of a
 * trajectory at rows >= n_samples[i] are NaN. */
typedef struct rb_sample_slab {
    int64_t offset;      /* doubles from the start of the arena */
    int64_t first_row;   /* sample index of the first row (sample k is step k*out_stride) */
    int64_t n_rows;
    int64_t first_traj, n_traj;   /* the partition */
    int64_t width;       /* doubles per field row (>= n_live, multiple of 16) */
    int64_t first_step;  /* step index (run frame) at which the launch started */
    int64_t n_live;      /* valid columns; -1 where the call could not know it (rb_propagate: read slab_live) */
} rb_sample_slab;

typedef struct rb_out {
    /* per-sample output, any layout expressible by two strides (elements), or NULL:
     * field f of sample k of trajectory i at samples[k*sample_stride + f*field_stride + i]. */
    double *samples;
    int64_t sample_stride;   /* e.g. 8*n for [n_out][8][n] */
    int64_t field_stride;    /* e.g. n */
    int64_t n_out_capacity;  /* samples the buffer can hold (>= n_steps/out_stride + 1) */
    /* per-trajectory summary; final_state is REQUIRED (it doubles as the in-place state),
     * SoA [6][n]: the dense-output state at the impact time, or the last step's state. */
    double *final_state;
    int32_t *impact_step;    /* [n], -1 = none */
    double *impact_time;     /* [n], NaN = none */
    uint8_t *status;         /* [n] rb_traj_status (REQUIRED) */
    int32_t *n_samples;      /* [n] valid samples of trajectory i (grid times not after the event), or NULL */
    /* RB_FLAG_PACKED_SAMPLES only (ignored otherwise): `samples` is then a device arena of packed_capacity doubles
     * (sample_stride / field_stride / n_out_capacity unused) */
    int64_t packed_capacity;
    rb_sample_slab *slabs;   /* HOST array [slab_capacity], filled while the launches are enqueued */
    int64_t slab_capacity;   /* two per launch are needed at most (two partitions) */
    int64_t *n_slabs;        /* HOST out: slabs written */
    int32_t *slab_live;      /* DEVICE [slab_capacity] or NULL: n_live of every slab, written by its launch */
} rb_out;

/* Replaces SpacecraftModel.run_simulation (spacecraft_model.py:490-519; solve_ivp call :516)
 * for ensembles.  Asynchronous on `stream` (see RB_FLAG_EARLY_EXIT). */
int32_t rb_propagate(rb_ctx *ctx, const rb_in *in, const rb_run *run, const rb_out *out, void *stream);

/* Launch statistics of the last rb_propagate on this context (host-side counters). */
typedef struct rb_stats {
    int64_t kernel_launches;     /* all kernels of the call */
    int64_t propagate_launches;  /* launches of the step kernel */
    int64_t steps_enqueued;      /* steps covered by those launches */
    int64_t last_polled_live;    /* last live count the host saw (-1 if never polled) */
    double propagate_ms;         /* device time of the step-kernel launch sequence, first launch to last
                                    (RB_FLAG_PROFILE only, else 0) */
    int64_t sample_bytes_sent;   /* rb_propagate_host_packed: bytes of sample slabs the copy engine moved to the host */
} rb_stats;
int32_t rb_ctx_stats(const rb_ctx *ctx, rb_stats *out);

/* ---- per-sample diagnostics: the dict of SpacecraftModel.equations_of_motion
 * (spacecraft_model.py:472-487) for n given (t, y) pairs; used for additional_data
 * (spacecraft_model.py:517-518).  t device [n]; y SoA [6][n]; out SoA [rows][n]:
 * a_total[3], a_grav[3], a_j2[3], a_moon[3], a_sun[3], a_drag[3], altitude, |latitude| deg, rho, T,
 * |v_rel| (RB_DIAG_FIELDS = 23 rows), then -- when `thermal` is not NULL -- the six values of
 * spacecraft_temperature (spacecraft_model.py:243-255, N1) in ITS return order Qc, Qr, Q_net, Q, T_s,
 * dT (RB_THERMAL_FIELDS = 6 more rows; the reference's caller at :468 stores them under permuted
 * names, which the Python mirror reproduces). */
#define RB_DIAG_FIELDS 23
#define RB_THERMAL_FIELDS 6
typedef struct rb_thermal {
    double capsule_length;          /* SpacecraftModel.height, spacecraft_model.py:396 */
    double dt, iter_fact;           /* iterations = int(dt / iter_fact), :246 */
    double thermal_conductivity, specific_heat_capacity, emissivity, ablation_efficiency; /* material, :405-408 */
} rb_thermal;
int32_t rb_equations_of_motion(rb_ctx *ctx, int64_t n, const double *t, const double *y, double epoch_jd,
                               double gmst0, const double *cd, const double *area, const double *mass,
                               double cd_scalar, double area_scalar, double mass_scalar,
                               const rb_thermal *thermal, double *out, void *stream);

/* ---- ground-track post-pass: app.py:225-243 (per-sample ECI -> ECEF ->
 * ecef_to_geodetic, coordinate_converter.py:81-96), app.py:302-308 (cumulative haversine down-range,
 * cc:125-136), spacecraft_visualization.py:562-577 (crossings of an altitude threshold, 100 km in the
 * app).  samples as written by rb_propagate; sample k is at t0 + k*sample_dt.  Outputs (device, any
 * may be NULL): geo [n_out][3][n] = lat deg, lon deg, h m; downrange [n_out][n] m; time / down-range
 * of the first crossing [n] (NaN = none); number of crossings [n]. */
int32_t rb_ground_track(rb_ctx *ctx, int64_t n, int64_t n_out, const double *samples, int64_t sample_stride,
                        int64_t field_stride, const int32_t *n_samples, double t0, double sample_dt, double gmst0,
                        double threshold, double *geo, double *downrange, double *cross_time,
                        double *cross_downrange, int32_t *n_cross, void *stream);

/* Geodetic impact (or end) point of every trajectory from rb_propagate's summaries: final_state SoA
 * [6][n] taken at impact_time[i] (impacted) or t_end; out SoA [3][n] = lat deg, lon deg, h m.  (The
 * app's own touchdown read-out, app.py:322-327, feeds un-rotated ECI coordinates to ecef_to_geodetic;
 * this is the rotated -- correct -- form, equal to the app's geodetic_coords[-1] convention.) */
int32_t rb_impact_points(rb_ctx *ctx, int64_t n, const double *final_state, const double *impact_time,
                         const uint8_t *status, double t_end, double gmst0, double *out, void *stream);

/* ---- adaptive mode: the reference's integrator AS SHIPPED ----------------
 * run_simulation calls scipy.integrate.solve_ivp(method='RK45', rtol=1e-8, atol=1e-10, t_eval,
 * events=[altitude_event]) (spacecraft_model.py:516).  One thread per trajectory, each on its own time
 * axis: scipy's initial-step heuristic, error-controlled Dormand-Prince steps, terminal altitude event
 * (Brent on the dense output) and samples at t_eval_k = t_eval0 + k*t_eval_dt (k < n_eval, <= the event
 * time) interpolated on each step's quartic dense output.  Parity with the reference is at the level of
 * the solver tolerance, not of rounding: the error norm is a cancellation of O(1e8), so accept/reject
 * decisions and step sizes depend on summation order (the reference's own run depends on the host BLAS). */
typedef struct rb_adaptive_run {
    double t0, t_bound;          /* t_span */
    double rtol, atol;           /* 1e-8, 1e-10 in the reference */
    double t_eval0, t_eval_dt;   /* np.arange(ts, tf, dt), app.py:171 */
    int64_t n_eval;
    int64_t max_steps;           /* accepted + rejected attempts per trajectory */
    double event_altitude;       /* 1000.0 */
    double epoch_jd, gmst0;
    double ephemeris_dt;         /* 0: Sun / Moon / solar factor evaluated at every stage time, as the reference does;
                                    > 0: evaluated on a grid of this spacing (seconds, on the device) and interpolated
                                    (4-point Lagrange; error 3e-11 m on the Moon's position at 16 s): faster ensembles */
    int32_t method;              /* RB_METHOD_RK45 (the reference's default sim_type), RB_METHOD_RK23 or RB_METHOD_DOP853: solve_ivp's
                                    `method`, which run_simulation takes from SpacecraftModel.sim_type (spacecraft_model.py:516; the
                                    app's solver list, app.py:51).  The implicit ones (Radau / BDF / LSODA) are not built */
    int32_t reserved;
} rb_adaptive_run;
#define RB_METHOD_RK45 0
#define RB_METHOD_RK23 1
#define RB_METHOD_DOP853 2

typedef struct rb_adaptive_out {
    double *samples;             /* [n_eval][8][n] via sample_stride / field_stride, or NULL */
    int64_t sample_stride, field_stride;
    double *final_state;         /* SoA [6][n] REQUIRED: dense-output state at the event, else state at t_bound */
    double *impact_time;         /* [n] NaN = none, or NULL */
    uint8_t *status;             /* [n] REQUIRED */
    int32_t *n_samples, *n_accepted, *n_rejected; /* [n] or NULL */
    double *step_log;            /* [step_capacity][7][n] or NULL: start time and start state (t_old, y_old[6]) of every ACCEPTED
                                    step -- the roots and states solve_ivp reports for run_simulation's progress_event
                                    (spacecraft_model.py:505-511: it returns 0 on every step, so each step is an "event" whose
                                    root is the step's start, ivp.py:677-699) = sol.t_events[1] / sol.y_events[1].  Steps beyond
                                    the capacity are not logged (n_accepted still counts them) */
    int64_t step_capacity;
} rb_adaptive_out;

int32_t rb_propagate_adaptive(rb_ctx *ctx, const rb_in *in, const rb_adaptive_run *run, const rb_adaptive_out *out,
                              void *stream);

/* ---- host-buffer convenience (the end-to-end call of a ctypes / cgo style binding) ----
 * Everything is HOST memory (pinned memory from rb_host_alloc makes the copies asynchronous):
 * y0 [n][6]; cd/area/mass [n] or NULL; samples [n_out][8][n] or NULL; final_state [n][6].
 * Builds the time table, uploads, propagates, downloads, synchronises.
 * samples: only rows < *n_samples_used are touched.  If the buffer is pinned (rb_host_alloc,
 * cudaHostAlloc, cudaHostRegister) rows travel while they are produced: dense rows by copy engine from a
 * device staging buffer, the sparse rows of the retiring phase as direct (zero-copy) stores of the step
 * kernel; entries of trajectory i at rows >= n_samples[i] are NaN or keep what the caller had there.  A
 * pageable buffer is filled through a device staging buffer after the fact and those entries are NaN. */
int32_t rb_propagate_host(rb_ctx *ctx, int64_t n, const double *y0, const double *cd, const double *area,
                          const double *mass, double cd_scalar, double area_scalar, double mass_scalar,
                          double epoch_jd, double gmst0, double t0, double dt, const rb_run *run,
                          double *samples, int64_t n_out_capacity, double *final_state,
                          int32_t *impact_step, double *impact_time, uint8_t *status, int32_t *n_samples,
                          int64_t *n_samples_used);
/* The same call with PACKED sample slabs (RB_FLAG_PACKED_SAMPLES is implied): `packed` is a host arena of
 * packed_capacity doubles (pinned: every slab is sent by ONE contiguous copy-engine transfer as soon as its launch
 * has finished, while later launches run; no SM ever waits on the link and only live columns cross it), `slabs` a
 * host array the call fills (n_live included).  The time table is built on the host in chunks while the GPU
 * already runs, and only as far as the ensemble actually flies.  RB_ERR_SAMPLE_CAPACITY if an arena or slab list
 * is too small (nothing useful is returned then). */
int32_t rb_propagate_host_packed(rb_ctx *ctx, int64_t n, const double *y0, const double *cd, const double *area,
                                 const double *mass, double cd_scalar, double area_scalar, double mass_scalar,
                                 double epoch_jd, double gmst0, double t0, double dt, const rb_run *run,
                                 double *packed, int64_t packed_capacity, rb_sample_slab *slabs,
                                 int64_t slab_capacity, int64_t *n_slabs, int64_t *packed_used,
                                 double *final_state, int32_t *impact_step, double *impact_time, uint8_t *status,
                                 int32_t *n_samples);
/* Host-side format conversion (no physics): scatter packed slabs into dense rows dense[k][f][i] (n_out rows;
 * everything not covered by a valid sample is NaN).  All pointers are HOST memory. */
int32_t rb_unpack_samples(const double *packed, const rb_sample_slab *slabs, int64_t n_slabs, int64_t n,
                          const int32_t *impact_step, const uint8_t *status, int64_t n_out, double *dense,
                          int32_t n_threads);
/* Generation counter of the context's staged time table (incremented by every rb_ctx_set_time_table and by the
 * *_host calls, which restage it): a caller that caches "the table for (epoch, gmst0, t0, dt) is staged" compares it. */
int64_t rb_ctx_table_generation(const rb_ctx *ctx);
void *rb_host_alloc(int64_t bytes); /* pinned host memory (cudaHostAlloc) or NULL */
void rb_host_free(void *p);

/* Self-test hook for the library's own FP64 primitives (csrc/rb_math.cuh), elementwise on device
 * arrays: op 0 rcp(x), 1 rsqrt(x), 2 exp(x), 3 log(x) for x in [0.7, 1.3], 4 atan2(x, y) for
 * x, y >= 0 on the unit circle, 5 pow(x, y) for x in [0.7, 1.3], 6 log1p(x) for |x| <= 0.3, 7 (1 + x)^y as the
 * density's pressure law forms it; the budgeted variants of the hot RHS: 8 exp(x) without its last term, 9 log1p(x)
 * without its last two, 10 sqrt(x), 12 rsqrt(x), 13 rcp(x) by one Newton step, 11 y / x^1.5 (third-body factor), 14 atan2 as 4 without its last term; 15 sqrt(x) at full precision.
 * y may be NULL for unary ops. */
int32_t rb_math_probe(rb_ctx *ctx, int32_t op, int64_t n, const double *x, const double *y, double *out,
                      void *stream);

/* FP64 FMA peak probe used for the roofline denominator (MEASURED_PEAKS.json has no FP64
 * figure): runs a register-resident DFMA kernel for about `ms` milliseconds on `stream`,
 * returns achieved TFLOP/s (DFMA = 2 flop). */
int32_t rb_fp64_peak_probe(rb_ctx *ctx, double ms, double *tflops_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* REENTRY_B200_H */
