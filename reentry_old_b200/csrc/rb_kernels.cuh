// rb_kernels.cuh -- sm_100a kernels of the ensemble propagator.
//
//   k_init            copy y0 -> in-place state (SoA), identity live list, sample 0
//   k_propagate       THE hot kernel: one FP64 thread per live trajectory, fixed-step Dormand-Prince
//                     (scipy RK45 tableau, FSAL, Nystrom stage form), time-table tiles staged into shared
//                     memory with cp.async.bulk (TMA bulk copy) + mbarrier, double buffered; the
//                     acceleration rows Ka_0..Ka_6 of the stage derivatives parked in shared memory so
//                     that 80 registers buy 24 resident warps; warp-vote retirement of impacted lanes;
//                     SoA streaming stores of the decimated samples (to HBM or straight to pinned host
//                     memory).  Template bool SKIP = RB_FLAG_SKIP_NEGLIGIBLE_DRAG instantiation.
//   k_count_live / k_scan_blocks / k_scatter_live
//                     order-preserving compaction of the live-trajectory index list
//   k_refine          impact-time refinement on the step's quartic dense output (Brent), n_samples
//   k_adaptive        N4: scipy's error-controlled RK45 as run_simulation calls it, one thread per trajectory
//   k_initial_states / k_dispersed_states
//                     batched get_initial_state; N3 counter-based (Philox4x32-10) dispersion generator
//   k_eom             N1 per-sample diagnostics (equations_of_motion dict incl. the thermal entries)
//   k_ground_track / k_impact_points
//                     N2 the app's post-pass (geodetic coordinates, haversine down-range, crossings)
//   k_math_probe, k_dfma_peak
//                     self-test of rb_math.cuh; FP64 FMA peak probe for the roofline denominator
#pragma once

#include <cuda_runtime.h>
#include <float.h>

#include <type_traits>

#include "rb_dop853_tables.h"
#include "reentry_b200.h"
#include "rb_physics.cuh"

namespace rb {

#ifndef RB_TILE_STEPS
#define RB_TILE_STEPS 8
#endif
#ifndef RB_TILE_BUFFERS
#define RB_TILE_BUFFERS 2
#endif
constexpr int TILE_BUFFERS = RB_TILE_BUFFERS;  // 1 = single buffer (TMA latency exposed once per tile), 2 = double
constexpr int TILE_STEPS = RB_TILE_STEPS;                                      // steps per time-table tile
constexpr int TILE_ENTRIES = RB_STAGES_PER_STEP * TILE_STEPS + 1;  // 41 entries
constexpr int TILE_DOUBLES = TILE_ENTRIES * HOT_DOUBLES;            // 492 doubles = 3936 B
constexpr int K_SLOTS = 6;                                          // Ka_0..Ka_5 (acceleration rows only; Ka_6 = next step's Ka_0 goes to slot 0)
constexpr int K_ROWS = 3;

__constant__ Tableau c_T = make_tableau();  // run-time operands from the constant bank
__constant__ double c_P[7][4] = {RB_DP_P_ROWS};

// Stage state of stage S (1..6; 6 = the step's end point, row 6 of A = B), tableau pre-scaled by h (HTableau):
//   v_s = v_n + sum_j (h a_sj) Ka_j ;  r_s = r_n + (c_s h) v_n + sum_j (h^2 abar_sj) Ka_j
// Stages 1..5 accumulate straight onto the state (every fma has a constant-bank operand); the end point sums the
// increment first and adds it once, as scipy's y_new = y + h (K^T b) does, so the carried state gets one rounding.
// Ka_j lives in shared memory at (32-bit shared address) Kt + (slot_j*3 + c)*BLOCK*8; slot_0 = k0 (FSAL toggle), slot_j = j.
// k_refine repeats exactly this arithmetic (stage_update below) when it recomputes the crossing step.
template <int S>
RB_HD void stage_update(const HTableau &ht, const double k[6][3], const double y[6], double ys[6]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double base = fma(ht.ch[S], y[3 + c], y[c]);
        if (S < 6) {
            double v = y[3 + c], r = base;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                v = fma(ht.ha[S][j], k[j][c], v);
                if (j <= S - 2) r = fma(ht.hha[S][j], k[j][c], r);   // abar[s][s-1] = 0
            }
            ys[3 + c] = v; ys[c] = r;
        } else {
            double dv = ht.ha[6][0] * k[0][c], dr = ht.hha[6][0] * k[0][c];
#pragma unroll
            for (int j = 2; j < 6; ++j) dv = fma(ht.ha[6][j], k[j][c], dv);      // B[1] = 0
#pragma unroll
            for (int j = 1; j < 5; ++j) dr = fma(ht.hha[6][j], k[j][c], dr);
            ys[3 + c] = y[3 + c] + dv; ys[c] = base + dr;
        }
    }
}

// The K rows are addressed through ONE 32-bit shared-memory address kept in a register (the compiler otherwise
// rebuilds the generic pointer -- thread id, shared window base -- twice per stage).
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double &a, double &b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double a, double b) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
// Stage state of stage S from the K rows in shared memory.  Ka_j is stored as an (x, y) PAIR at kp + j*BLOCK*16 (kp = 32-bit
// shared address of this thread's pair in slot 0: one 16-byte load) and its z at kz + j*BLOCK*8: two loads per row instead
// of three; all offsets are immediates of the LDS.  Ka_{S-1} has just been evaluated: it comes in registers (klast) -- the
// same value its slot holds, no load at all.
template <int S, int BLOCK>
__device__ __forceinline__ void stage_state(uint32_t kp, uint32_t kz, const double y[6], const HTableau &ht, double ys[6], const double klast[3]) {
    double k[6][3];
#pragma unroll
    for (int j = 0; j < S; ++j) {
        if (j == S - 1) { k[j][0] = klast[0]; k[j][1] = klast[1]; k[j][2] = klast[2]; }
        else {
            lds_f64x2(kp + (uint32_t)(j * BLOCK * 16), k[j][0], k[j][1]);
            k[j][2] = lds_f64(kz + (uint32_t)(j * BLOCK * 8));
        }
    }
    stage_update<S>(ht, k, y, ys);
}

struct StepParams {
    // time table
    const double *table;       // device, [entries][16]
    int64_t table_step0;       // absolute step index (table frame) of this launch's first step
    int n_launch_steps;        // steps in this launch
    // run frame
    int64_t run_step0;         // step index within the run of this launch's first step
    int64_t out_stride;
    double h;
    HTableau ht;               // tableau scaled by h
    double event_altitude;
    // state / per-trajectory constants
    double *state;             // SoA [6][n]
    int64_t n;
    const double *cd, *area, *mass;
    double cd_s, area_s, mass_s;
    const int32_t *live;       // [n_live] trajectory ids, ascending
    const int32_t *n_live;     // device scalar
    // outputs
    double *samples;
    int64_t sample_stride, field_stride;
    int32_t *impact_step;
    uint8_t *status;
    // packed sample slabs (RB_FLAG_PACKED_SAMPLES): `samples` is the base of THIS launch's slab, whose first row is
    // sample `sample_row0`, MINUS sample_row0 * sample_stride (row k of the run at samples + k * sample_stride, as in
    // dense mode); the column of a trajectory is its position in the launch's live list (dense mode: its index).
    // Rows are therefore as wide as the live ensemble, not as the initial one.
    int packed;
    int64_t k_first, k_last;   // sample rows this launch writes (k_last < k_first: none); host-computed (no 64-bit division in the kernel)
    int64_t sample_row0;
    int64_t slab_offset;       // element offset of the slab within the arena (for edge_slot)
    int32_t *slab_live;        // device scalar or NULL: live count at launch start (= valid columns of the slab)
    int64_t *edge_slot;        // [2n] packed mode: arena offset + field stride of the row at the end of the crossing step
};

// ---------------------------------------------------------------- PTX helpers (TMA bulk + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

// ---------------------------------------------------------------- init
struct InitParams {
    const double *y0;
    int64_t y0_traj_stride, y0_comp_stride;
    double *state;
    int64_t n;
    int32_t *live;
    int32_t *impact_step;
    double *impact_time;
    uint8_t *status;
    double *samples;
    int64_t sample_stride, field_stride;
};

__global__ void k_init_resume(int32_t *live, int64_t n) {   // RB_FLAG_RESUME: the buffers ARE the checkpoint
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) live[i] = (int32_t)i;
}

__global__ void k_init(const InitParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    double y[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        y[c] = p.y0[i * p.y0_traj_stride + c * p.y0_comp_stride];
        p.state[c * p.n + i] = y[c];
    }
    p.live[i] = (int32_t)i;
    p.status[i] = RB_TRAJ_ALIVE;
    if (p.impact_step) p.impact_step[i] = -1;
    if (p.impact_time) p.impact_time[i] = nan("");
    if (p.samples) {
        double *s = p.samples + i;
#pragma unroll
        for (int c = 0; c < 6; ++c) st_stream(s + c * p.field_stride, y[c]);
        st_stream(s + 6 * p.field_stride, sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]) - RB_EARTH_R);
        st_stream(s + 7 * p.field_stride, sqrt(y[3] * y[3] + y[4] * y[4] + y[5] * y[5]));
    }
}

// packed slabs: rows first..last (sample indices) of a column that will not be written any more get NaN, so that a
// slab needs no prefill (a lane retires once; at most steps_per_launch / out_stride rows are left in its launch)
// (scalars, not the parameter struct: taking its address would make the kernel copy it to local memory)
__device__ __noinline__ void nan_rows_at(double *column, int64_t sample_stride, int64_t field_stride, int64_t n_rows) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t k = 0; k < n_rows; ++k) {
        double *sp = column + k * sample_stride;
#pragma unroll
        for (int c = 0; c < RB_SAMPLE_FIELDS; ++c) st_stream(sp + c * field_stride, qnan);
    }
}
#define nan_rows(p, col, first, last) \
    nan_rows_at((p).samples + (first) * (p).sample_stride + (col), (p).sample_stride, (p).field_stride, (last) - (first) + 1)

// ---------------------------------------------------------------- the hot kernel
// Shared memory per block: [2 x tile][pairs: Ka (x, y) of 6 slots + y_n as 3 pairs, 9 x BLOCK x 16 B]
//                          [scalars: Ka z of 6 slots, bc, g_old, 8 x BLOCK x 8 B][lookup-table image, TAB_DOUBLES][2 mbarriers]
// (same size as 26 rows of BLOCK doubles)
constexpr int PARK_ROWS = 8;   // per-thread doubles parked in shared memory: y_n[6], bc, g_old
template <int BLOCK>
__host__ __device__ constexpr size_t propagate_smem_bytes() {
    return sizeof(double) * (TILE_BUFFERS * TILE_DOUBLES + (K_SLOTS * K_ROWS + PARK_ROWS) * BLOCK + TAB_DOUBLES) + 2 * sizeof(uint64_t);
}

// The acceleration evaluation exists ONCE in the kernel, as a subroutine the seven evaluation sites of a step call
// (ptxas allocates registers across the call: no ABI spills).  The earlier form -- one flattened stage loop around an
// inlined evaluation with a `switch` over the stage-state arithmetic -- spent ~40 of its ~430 instructions per stage
// on dispatch (jump table, K-slot select, loop state); a call and a return cost 3.
//   ent   32-bit shared address of the 96-byte time-table entry of the stage time
//   bcp   32-bit shared address of this thread's parked body constant
//   kdp, kdz  32-bit shared addresses where a[0..1] (one 16-byte store) and a[2] go (this thread's place in the K slot)
// Returns |r| - R (the event function's altitude) of the evaluated state.
#ifndef RB_RHS_INLINE
#define RB_RHS_INLINE 1   // 1 (default): the evaluation is inlined at its seven sites; 0: one subroutine, seven calls.
#endif                    // Measured on B200 (profiles/r2a_*): inlined 2499 thread-instructions per trajectory-step in the dense
                          // atmosphere against 2733 for the subroutine (12 argument moves per call, as many as the old flattened
                          // loop spent on its `switch`); its 61 KB of code cost 0.7 no-instruction stalls per issue and it is
                          // still 4.6 % faster (the FP64 issue slots, 2 cycles each, are 72 % of the issue port either way).
#if RB_RHS_INLINE
#define RB_RHS_LINKAGE __forceinline__
#else
#define RB_RHS_LINKAGE __noinline__
#endif
//   tabs  32-bit shared address of the block's lookup-table image (rb_math.cuh TabShared)
//   a     the acceleration, also handed back in registers: the next stage's state uses it without reloading it
template <int BLOCK, bool SKIP>
__device__ RB_RHS_LINKAGE double rhs_eval(uint32_t ent, uint32_t bca, uint32_t kdp, uint32_t kdz, uint32_t tabs, double x, double y,
                                        double z, double vx, double vy, double vz, double a[3]) {
    const double *e = reinterpret_cast<const double *>(__cvta_shared_to_generic((size_t)ent));
    const double r[3] = {x, y, z}, v[3] = {vx, vy, vz};
    const double *bcp = reinterpret_cast<const double *>(__cvta_shared_to_generic((size_t)bca));
    double alt;
    acceleration<false, SKIP, TabShared>(e, r, v, bcp, a, nullptr, &alt, TabShared{tabs});
    sts_f64x2(kdp, a[0], a[1]); sts_f64(kdz, a[2]);
    return alt;
}

template <int BLOCK, int MIN_BLOCKS, bool SKIP>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_propagate(const StepParams p) {
    // (STATIC shared memory was tried so that table addresses would be link-time constants: ptxas then rebuilds the shared
    // base from a special register in every stage -- S2R + LEA + MOV -- instead of holding it in one register: no gain)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);
    double *Ks = tiles + TILE_BUFFERS * TILE_DOUBLES;
    double *Park = Ks + K_SLOTS * K_ROWS * BLOCK;
    double *Tab = Park + PARK_ROWS * BLOCK;
    uint64_t *bars = reinterpret_cast<uint64_t *>(Tab + TAB_DOUBLES);

    const int n_live = *p.n_live;
    const int64_t i0 = (int64_t)blockIdx.x * BLOCK;
    if (i0 >= n_live) return;  // whole block beyond the live list
    const int64_t i = i0 + threadIdx.x;
    bool alive = i < n_live;

    const int n_tiles = (p.n_launch_steps + TILE_STEPS - 1) / TILE_STEPS;
    auto tile_steps_of = [&](int t) { return min(TILE_STEPS, p.n_launch_steps - t * TILE_STEPS); };
    auto issue_tile = [&](int t) {  // thread 0 only
        const uint32_t bytes = (uint32_t)((RB_STAGES_PER_STEP * tile_steps_of(t) + 1) * HOT_DOUBLES * sizeof(double));
        const double *src = p.table + (RB_STAGES_PER_STEP * (p.table_step0 + (int64_t)t * TILE_STEPS)) * HOT_DOUBLES;
        uint64_t *bar = &bars[t % TILE_BUFFERS];
        mbar_expect_tx(bar, bytes);
        bulk_g2s(tiles + (t % TILE_BUFFERS) * TILE_DOUBLES, src, bytes, bar);
    };
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
        fence_proxy_async();
        issue_tile(0);
        if (TILE_BUFFERS > 1 && n_tiles > 1) issue_tile(1);
    }

    // per-trajectory state
    int32_t tid = 0;
    double y[6] = {0, 0, 0, 0, 0, 0};
    double b_cd = p.cd_s, b_area = p.area_s, b_mass = p.mass_s;
    bool loaded = false;
    if (alive) {
        tid = p.live[i];
        alive = p.status[tid] == RB_TRAJ_ALIVE;  // without compaction the list still names retired lanes
    }
    if (alive) {
        loaded = true;
#pragma unroll
        for (int c = 0; c < 6; ++c) y[c] = p.state[c * p.n + tid];
        if (p.cd) b_cd = p.cd[tid];
        if (p.area) b_area = p.area[tid];
        if (p.mass) b_mass = p.mass[tid];
    }
    // column of this lane's samples: trajectory index (dense rows) or position in the live list (packed slabs; the
    // host passes `samples` biased so that row k of the run is at samples + k * sample_stride in both modes).
    // Recomputed where it is used (sample steps, retirement) rather than carried in registers through the stages.
#define RB_SAMPLE_COL() (p.packed ? (int64_t)blockIdx.x * BLOCK + threadIdx.x : (int64_t)tid)
    if (p.packed) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && p.slab_live) *p.slab_live = n_live;
        // a listed lane that is already retired (only without compaction) owns columns nobody else writes
        if (!alive && i < n_live && p.samples) nan_rows(p, RB_SAMPLE_COL(), p.k_first, p.k_last);
    }
    // 32-bit shared addresses of this thread's columns, kept opaque so that they live in ONE register each instead of
    // being rebuilt (thread id, shared window base) at every use
    uint32_t Kp = smem_u32(Ks) + threadIdx.x * 16;    // pairs: Ka (x, y) of slot s at Kp + s*BLOCK*16, y_n pairs at slots 6..8
    asm volatile("" : "+r"(Kp));
    uint32_t Kz = smem_u32(Ks) + 9 * BLOCK * 16 + threadIdx.x * 8;   // scalars: Ka z of slot s at Kz + s*BLOCK*8, bc row 6, g_old row 7
    asm volatile("" : "+r"(Kz));
    constexpr uint32_t PB = BLOCK * 16, ZB = BLOCK * 8;   // bytes per pair row / per scalar row
    sts_f64(Kz + 6 * ZB, make_body(b_cd, b_area, b_mass).bc);
    sts_f64(Kz + 7 * ZB, event_g(y, p.event_altitude));
    // the block's copy of the lookup tables (2^(j/32), atan reference angles, atmosphere layer records)
    for (int j = threadIdx.x; j < TAB_DOUBLES; j += BLOCK)
        Tab[j] = j < TAB_ATAN ? g_exp_tab[j] : (j < TAB_LAYER ? (&g_atan_tab[0][0])[j - TAB_ATAN] : (&g_layer[0][0])[j - TAB_LAYER]);
    uint32_t Tb = smem_u32(Tab);
    asm volatile("" : "+r"(Tb));
    __syncthreads();  // barrier init and table image visible to all

    int to_sample = (int)(p.k_first * p.out_stride - p.run_step0);  // steps until the next stored sample
    int64_t next_k = p.k_first;
    double ka[3] = {0.0, 0.0, 0.0};   // the acceleration evaluated last (in registers as well as in its K slot)
    // (A ROTATED form of this loop -- each iteration opening with the FSAL evaluation, which removes the separate first
    //  evaluation and one of the seven inlined copies of the acceleration code, 51 KB instead of 57 -- measured 0.5 %
    //  SLOWER (10.53 vs 10.58 G trajectory-steps/s): with the lookup tables in shared memory the instruction cache no
    //  longer limits, and the rotation adds a per-step test.  It also ran into a ptxas 12.9 defect worth knowing: a loop
    //  variable whose INITIAL value is the address of the dynamic shared memory itself is taken for "the shared-memory
    //  base" and other shared addresses are rematerialised from its register inside the loop, after it has moved on.
    //  Below, `ent` is derived per tile and never equals that constant symbolically.)
    int t = 0;
    for (; t < n_tiles; ++t) {
        mbar_wait(&bars[t % TILE_BUFFERS], (uint32_t)((t / TILE_BUFFERS) & 1));
        uint32_t ent = smem_u32(tiles + (t % TILE_BUFFERS) * TILE_DOUBLES);   // entry 5 m of the tile = t_n of its step m
        const int ts = tile_steps_of(t);
        // the launch's very first evaluation: Ka_0 = a(t_n, y_n) (later steps inherit it from the previous step: FSAL)
        if (t == 0 && alive) rhs_eval<BLOCK, SKIP>(ent, Kz + 6 * ZB, Kp, Kz, Tb, y[0], y[1], y[2], y[3], y[4], y[5], ka);
        for (int m = 0; m < ts; ++m) {
            if (alive) {
                constexpr uint32_t EB = HOT_DOUBLES * 8;   // bytes per table entry
                double ys[6];
                stage_state<1, BLOCK>(Kp, Kz, y, p.ht, ys, ka);
                rhs_eval<BLOCK, SKIP>(ent + 1 * EB, Kz + 6 * ZB, Kp + 1 * PB, Kz + 1 * ZB, Tb, ys[0], ys[1], ys[2], ys[3], ys[4], ys[5], ka);
                stage_state<2, BLOCK>(Kp, Kz, y, p.ht, ys, ka);
                rhs_eval<BLOCK, SKIP>(ent + 2 * EB, Kz + 6 * ZB, Kp + 2 * PB, Kz + 2 * ZB, Tb, ys[0], ys[1], ys[2], ys[3], ys[4], ys[5], ka);
                stage_state<3, BLOCK>(Kp, Kz, y, p.ht, ys, ka);
                rhs_eval<BLOCK, SKIP>(ent + 3 * EB, Kz + 6 * ZB, Kp + 3 * PB, Kz + 3 * ZB, Tb, ys[0], ys[1], ys[2], ys[3], ys[4], ys[5], ka);
                stage_state<4, BLOCK>(Kp, Kz, y, p.ht, ys, ka);
                rhs_eval<BLOCK, SKIP>(ent + 4 * EB, Kz + 6 * ZB, Kp + 4 * PB, Kz + 4 * ZB, Tb, ys[0], ys[1], ys[2], ys[3], ys[4], ys[5], ka);
                stage_state<5, BLOCK>(Kp, Kz, y, p.ht, ys, ka);
                rhs_eval<BLOCK, SKIP>(ent + 5 * EB, Kz + 6 * ZB, Kp + 5 * PB, Kz + 5 * ZB, Tb, ys[0], ys[1], ys[2], ys[3], ys[4], ys[5], ka);
#pragma unroll
                for (int c = 0; c < 3; ++c) sts_f64x2(Kp + (6 + c) * PB, y[2 * c], y[2 * c + 1]);
                stage_state<6, BLOCK>(Kp, Kz, y, p.ht, ys, ka);
#pragma unroll
                for (int c = 0; c < 6; ++c) y[c] = ys[c];
                const double alt_new = rhs_eval<BLOCK, SKIP>(ent + 5 * EB, Kz + 6 * ZB, Kp, Kz, Tb, y[0], y[1], y[2], y[3], y[4], y[5], ka);
                const double g_new = alt_new - p.event_altitude;
                const double g_old = lds_f64(Kz + 7 * ZB);
                const int64_t step_abs = p.table_step0 + (int64_t)t * TILE_STEPS + m + 1;
                const bool finite = (__double2hiint(alt_new) & 0x7ff00000) != 0x7ff00000;
                const bool hit = g_old >= 0.0 && g_new <= 0.0;
                if (!finite || hit) {
                    alive = false;
                    p.status[tid] = finite ? RB_TRAJ_IMPACTED : RB_TRAJ_NONFINITE;
                    if (p.impact_step) p.impact_step[tid] = (int32_t)step_abs;
#pragma unroll
                    for (int c = 0; c < 3; ++c) lds_f64x2(Kp + (6 + c) * PB, y[2 * c], y[2 * c + 1]);
                    if (p.packed && p.samples) {
                        if (finite && to_sample == 1) {
                            p.edge_slot[2 * (int64_t)tid] = p.slab_offset + (next_k - p.sample_row0) * p.sample_stride + RB_SAMPLE_COL();
                            p.edge_slot[2 * (int64_t)tid + 1] = p.field_stride;
                        }
                        nan_rows(p, RB_SAMPLE_COL(), next_k, p.k_last);
                    }
                } else {
                    sts_f64(Kz + 7 * ZB, g_new);
                    if (to_sample == 1 && p.samples) {
                        double *sp = p.samples + next_k * p.sample_stride + RB_SAMPLE_COL();
#pragma unroll
                        for (int c = 0; c < 6; ++c) st_stream(sp + c * p.field_stride, y[c]);
                        st_stream(sp + 6 * p.field_stride, alt_new);
                        const double v2 = fma(y[5], y[5], fma(y[3], y[3], y[4] * y[4]));
                        st_stream(sp + 7 * p.field_stride, sqrt(v2));
                    }
                }
            }
            ent += RB_STAGES_PER_STEP * HOT_DOUBLES * 8;
            if (--to_sample == 0) { to_sample = (int)p.out_stride; ++next_k; }
            if (!__any_sync(0xffffffffu, alive) && (threadIdx.x >= 32 || t + TILE_BUFFERS >= n_tiles)) goto done;
        }
        if (t + TILE_BUFFERS < n_tiles) {
            __syncthreads();
            if (threadIdx.x == 0) issue_tile(t + TILE_BUFFERS);
        }
    }
done:
    // (left through the warp vote) a tile issued ahead but not consumed must land before the block's shared memory is
    // released: the bulk copy and its mbarrier complete_tx would otherwise hit whatever block the SM schedules next
    if (TILE_BUFFERS > 1 && threadIdx.x == 0 && t + 1 < n_tiles) mbar_wait(&bars[(t + 1) % TILE_BUFFERS], (uint32_t)(((t + 1) / TILE_BUFFERS) & 1));
    if (loaded) {
#pragma unroll
        for (int c = 0; c < 6; ++c) p.state[c * p.n + tid] = y[c];
    }
#undef RB_SAMPLE_COL
}

// ---------------------------------------------------------------- compaction of the live list
constexpr int COMPACT_BLOCK = 1024;

__global__ void __launch_bounds__(COMPACT_BLOCK) k_count_live(const int32_t *live_in, const int32_t *n_live,
                                                              const uint8_t *status, int32_t *block_counts) {
    const int n = *n_live;
    const int64_t i = (int64_t)blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    if ((int64_t)blockIdx.x * COMPACT_BLOCK >= n) { if (threadIdx.x == 0) block_counts[blockIdx.x] = 0; return; }
    const int keep = (i < n) && (status[live_in[i]] == RB_TRAJ_ALIVE);
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// single block: exclusive scan of block_counts[0..n_blocks) in place; writes the total to n_live_out and, for the
// host's lagged early-exit poll, into mapped pinned memory (a posted store: a D2H memcpy on the compute stream would
// queue behind whatever sample traffic occupies the link / the copy engine and stall the next launch)
__global__ void __launch_bounds__(1024) k_scan_blocks(int32_t *block_counts, int n_blocks, int32_t *n_live_out,
                                                      int32_t *host_poll) {
    __shared__ int32_t warp_sums[32];
    __shared__ int32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = (idx < n_blocks) ? block_counts[idx] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int warp_off = (warp == 0) ? 0 : warp_sums[warp - 1];
        if (idx < n_blocks) block_counts[idx] = carry + warp_off + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *n_live_out = carry_s;
        if (host_poll) *host_poll = carry_s;
    }
}

__global__ void k_copy_i32(int32_t *dst, const int32_t *src) { *dst = *src; }

__global__ void __launch_bounds__(COMPACT_BLOCK) k_scatter_live(const int32_t *live_in, const int32_t *n_live,
                                                                const uint8_t *status, const int32_t *block_offsets,
                                                                int32_t *live_out) {
    __shared__ int32_t warp_sums[32];
    const int n = *n_live;
    if ((int64_t)blockIdx.x * COMPACT_BLOCK >= n) return;
    const int64_t i = (int64_t)blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t tid = 0;
    bool keep = false;
    if (i < n) { tid = live_in[i]; keep = status[tid] == RB_TRAJ_ALIVE; }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int rank_in_warp = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    if (keep) {
        const int off = block_offsets[blockIdx.x] + ((warp == 0) ? 0 : warp_sums[warp - 1]) + rank_in_warp;
        live_out[off] = tid;
    }
}

// ---------------------------------------------------------------- impact refinement + n_samples
struct RefineParams {
    const double *table;
    int64_t table_steps;
    double h, event_altitude;
    HTableau ht;
    int skip_drag;
    double *state;  // in: y at the start of the crossing step; out: dense-output state at the impact time
    int64_t n;
    const double *cd, *area, *mass;
    double cd_s, area_s, mass_s;
    const int32_t *impact_step;
    double *impact_time;
    const uint8_t *status;
    int32_t *n_samples;
    double *samples;
    int64_t sample_stride, field_stride, out_stride;
    int64_t run_first_step, run_n_steps;
    int refine;
    int32_t *edge_count;
    const int64_t *edge_slot;   // packed sample slabs: `samples` is the arena, the row's place comes from the step kernel
};

struct Dense {
    double Q[6][4];
    double y_old[6];
    double t_old, h, event_altitude;
    __device__ void eval(double t, double yo[6]) const {  // scipy rk.py:723-737
        const double x = (t - t_old) / h;
        const double p0 = x, p1 = p0 * x, p2 = p1 * x, p3 = p2 * x;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = Q[c][0] * p0;
            s += Q[c][1] * p1; s += Q[c][2] * p2; s += Q[c][3] * p3;
            yo[c] = h * s + y_old[c];
        }
    }
    __device__ double g(double t) const {
        double yo[6];
        eval(t, yo);
        return event_g(yo, event_altitude);
    }
};

// Brent's root finder with scipy's tolerances (xtol = rtol = 4 eps, ivp.py:76-77)
template <class DenseT>
__device__ __noinline__ double brent_event(const DenseT &d, double xa, double xb) {
    const double xtol = 4 * DBL_EPSILON, rtol = 4 * DBL_EPSILON;
    double xpre = xa, xcur = xb, xblk = 0., fblk = 0., spre = 0., scur = 0.;
    double fpre = d.g(xpre), fcur = d.g(xcur);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return nan("");
    for (int it = 0; it < 100; ++it) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) { xblk = xpre; fblk = fpre; spre = scur = xcur - xpre; }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2;
        const double sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) stry = -fcur * (xcur - xpre) / (fcur - fpre);
            else {
                const double dpre = (fpre - fcur) / (xpre - xcur), dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = d.g(xcur);
    }
    return xcur;
}

__global__ void __launch_bounds__(128) k_refine(const RefineParams p) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= p.n) return;
    const uint8_t st = p.status[tid];
    int32_t ns = (int32_t)(p.run_n_steps / p.out_stride + 1);
    if (st != RB_TRAJ_ALIVE && p.impact_step) {
        const int64_t local_prev = (int64_t)p.impact_step[tid] - 1 - p.run_first_step;  // last completed step
        ns = (int32_t)(local_prev / p.out_stride + 1);
    }
    // (an impact refined by an earlier chunk of a resumed run has a finite impact_time and is left alone)
    if (st == RB_TRAJ_IMPACTED && p.refine && p.impact_step && p.impact_time && p.impact_time[tid] != p.impact_time[tid]) {
        const int64_t n_abs = (int64_t)p.impact_step[tid] - 1;
        const double *e0 = p.table + (RB_STAGES_PER_STEP * n_abs) * HOT_DOUBLES;
        const Body body = make_body(p.cd ? p.cd[tid] : p.cd_s, p.area ? p.area[tid] : p.area_s, p.mass ? p.mass[tid] : p.mass_s);
        double y[6], K[7][6], ys[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) y[c] = p.state[c * p.n + tid];
        const double h = p.h;
#pragma unroll 1
        for (int s = 0; s <= 6; ++s) {
            double k[6][3];
            for (int j = 0; j < 6; ++j)
                for (int c = 0; c < 3; ++c) k[j][c] = (j < s) ? K[j][3 + c] : 0.0;
            switch (s) {   // the step kernel's own stage arithmetic
                case 0: for (int c = 0; c < 6; ++c) ys[c] = y[c]; break;
                case 1: stage_update<1>(p.ht, k, y, ys); break;
                case 2: stage_update<2>(p.ht, k, y, ys); break;
                case 3: stage_update<3>(p.ht, k, y, ys); break;
                case 4: stage_update<4>(p.ht, k, y, ys); break;
                case 5: stage_update<5>(p.ht, k, y, ys); break;
                default: stage_update<6>(p.ht, k, y, ys); break;
            }
            double a[3];
            if (p.skip_drag) acceleration<false, true>(e0 + ((s == 6) ? 5 : s) * HOT_DOUBLES, ys, ys + 3, body, a, nullptr);
            else acceleration<false, false>(e0 + ((s == 6) ? 5 : s) * HOT_DOUBLES, ys, ys + 3, body, a, nullptr);
            K[s][0] = ys[3]; K[s][1] = ys[4]; K[s][2] = ys[5];
            K[s][3] = a[0];  K[s][4] = a[1];  K[s][5] = a[2];
        }
        Dense d;
        for (int c = 0; c < 6; ++c) {
            for (int k = 0; k < 4; ++k) {
                double q = 0.0;
                for (int j = 0; j < 7; ++j) q += K[j][c] * c_P[j][k];  // Q = K^T P, rk.py:173
                d.Q[c][k] = q;
            }
            d.y_old[c] = y[c];
        }
        d.t_old = e0[H_T];
        d.h = h;
        d.event_altitude = p.event_altitude;
        const double t_new = d.t_old + h;
        const double te = brent_event(d, d.t_old, t_new);
        double yf[6];
        d.eval(te, yf);
        p.impact_time[tid] = te;
#pragma unroll
        for (int c = 0; c < 6; ++c) p.state[c * p.n + tid] = yf[c];
        // a grid sample that coincides with the event time is kept (scipy t_eval <= t semantics)
        const int64_t local_new = n_abs + 1 - p.run_first_step;
        if (te >= t_new && (local_new % p.out_stride) == 0) {
            if (p.samples) {
                double *sp = p.samples + (local_new / p.out_stride) * p.sample_stride + tid;
                int64_t fs = p.field_stride;
                if (p.edge_slot) { sp = p.samples + p.edge_slot[2 * tid]; fs = p.edge_slot[2 * tid + 1]; }
                for (int c = 0; c < 6; ++c) sp[c * fs] = ys[c];
                sp[6 * fs] = event_g(ys, 0.0);
                sp[7 * fs] = sqrt(ys[3] * ys[3] + ys[4] * ys[4] + ys[5] * ys[5]);
            }
            ns += 1;
            if (p.edge_count) atomicAdd(p.edge_count, 1);
        }
    }
    if (p.n_samples) p.n_samples[tid] = ns;
}

// ---------------------------------------------------------------- time entry (shared by host builder and k_eom)
// Sun / Moon on fixed Keplerian ellipses: spacecraft_model.py:270-300, :312-344
RB_HD void kepler_position(double M, double ecc, double a, double Om, double w, double inc, double out[3]) {
    double E = M;
    for (int k = 0; k < 10; ++k) E = E - (E - ecc * sin(E) - M) / (1 - ecc * cos(E));
    const double f = 2 * atan2(sqrt(1 + ecc) * sin(E / 2), sqrt(1 - ecc) * cos(E / 2));
    const double r = a * (1 - ecc * cos(E));
    const double xp = r * cos(f), yp = r * sin(f);
    const double cO = cos(Om), sO = sin(Om), cw = cos(w), sw = sin(w), ci = cos(inc), si = sin(inc);
    out[0] = xp * (cO * cw - sO * sw * ci) - yp * (sO * cw + cO * sw * ci);
    out[1] = xp * (sO * cw + cO * sw * ci) + yp * (cO * cw - sO * sw * ci);
    out[2] = xp * sw * si + yp * cw * si;
}

RB_HD void time_entry(double t, double epoch0, double gmst0, double e[ENTRY_DOUBLES]) {
    const double jd = epoch0 + t / 86400.0;            // spacecraft_model.py:442
    const double gmst = gmst0 + RB_EARTH_OMEGA * t;    // :443
    const double td = jd - RB_JD_AT_0;
    double sun[3], moon[3];
    kepler_position(RB_MOON_M0 + RB_MOON_MMOTION_RAD * td, RB_MOON_E, RB_MOON_A, RB_MOON_OMEGA, RB_MOON_W, RB_MOON_I, moon);
    const double L = RB_SUN_L0 + (2 * RB_PI / RB_YEAR_D) * td;
    kepler_position(L - RB_SUN_OMEGA - RB_SUN_W, RB_SUN_E, RB_SUN_A, RB_SUN_OMEGA, RB_SUN_W, RB_SUN_I, sun);
    const double sn = sqrt(sun[0] * sun[0] + sun[1] * sun[1] + sun[2] * sun[2]);
    const double mn = sqrt(moon[0] * moon[0] + moon[1] * moon[1] + moon[2] * moon[2]);
    const double sn3 = sn * (sn * sn), mn3 = mn * (mn * mn);
    for (int c = 0; c < 3; ++c) {
        e[E_SUN + c] = sun[c];
        e[E_MOON + c] = moon[c];
        e[E_SUN_IND + c] = RB_SUN_K * sun[c] / sn3;     // :362
        e[E_MOON_IND + c] = RB_MOON_K * moon[c] / mn3;
    }
    e[E_CG] = cos(gmst);
    e[E_SG] = sin(gmst);
    // time part of solar_activity_factor, :151-159 (python-sign modulo)
    const double months = (jd - RB_SOLAR_JD_MIN) / RB_DAYS_PER_MONTH;
    double md = fmod(months, RB_SOLAR_CYCLE_MONTHS);
    if (md != 0.0 && md < 0.0) md += RB_SOLAR_CYCLE_MONTHS;
    const double phase = md / RB_SOLAR_CYCLE_MONTHS * 2 * RB_PI;
    const double f107 = RB_SOLAR_F107_AVG + RB_F107_AMPLITUDE * sin(phase);
    e[E_SOLAR] = 1 + (f107 - RB_SOLAR_F107_AVG) / RB_SOLAR_F107_AVG;
    e[E_T] = t;
}

// ---------------------------------------------------------------- batched get_initial_state
// spacecraft_model.py:412-435 with geodetic_to_spheroid (cc:71-79), enu_to_ecef (cc:105-116), ecef_to_eci (cc:50-68)
__device__ __forceinline__ void initial_state(double v, double lat, double lon, double alt, double az, double gam,
                                              double sg, double cg, double y[6]) {
    const double la = RB_DEG_TO_RAD * lat, lo = RB_DEG_TO_RAD * lon;
    double sla, cla, slo, clo;
    sincos(la, &sla, &cla);
    sincos(lo, &slo, &clo);
    const double N = RB_EARTH_R / sqrt(1 - RB_EARTH_E2 * (sla * sla));
    const double xe = (N + alt) * cla * clo, ye = (N + alt) * cla * slo, ze = ((1 - RB_EARTH_E2) * N + alt) * sla;
    double saz, caz, sga, cga;
    sincos(RB_DEG_TO_RAD * az, &saz, &caz);
    sincos(RB_DEG_TO_RAD * gam, &sga, &cga);
    const double ve = v * saz * cga, vn = v * caz * cga, vu = v * sga;
    const double vxe = -slo * ve + -sla * clo * vn + cla * clo * vu;
    const double vye = clo * ve + -sla * slo * vn + cla * slo * vu;
    const double vze = cla * vn + sla * vu;
    y[0] = cg * xe - sg * ye; y[1] = sg * xe + cg * ye; y[2] = ze;
    y[3] = cg * vxe - sg * vye; y[4] = sg * vxe + cg * vye; y[5] = vze;
}

__global__ void k_initial_states(int64_t n, const double *v, const double *lat, const double *lon, const double *alt,
                                 const double *az, const double *gam, double gmst, double *y0, int64_t ts, int64_t cs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sg, cg, y[6];
    sincos(gmst, &sg, &cg);
    initial_state(v[i], lat[i], lon[i], alt[i], az[i], gam[i], sg, cg, y);
#pragma unroll
    for (int c = 0; c < 6; ++c) y0[i * ts + c * cs] = y[c];
}

// ---------------------------------------------------------------- N3: counter-based dispersion generator
// Philox4x32-10 keyed by the seed, counter = (global trajectory index, draw): the dispersion of
// trajectory i does not depend on how the ensemble is sharded over GPUs.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct DispersionParams {
    double nominal[6], sigma[6];
};

__global__ void k_dispersed_states(int64_t n, int64_t first_index, uint64_t seed, DispersionParams dp, double gmst,
                                   double *params /* SoA [6][n] or NULL */, double *y0, int64_t ts, int64_t cs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t idx = (uint64_t)(first_index + i);
    double q[6];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        uint32_t r[4];
        philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)d, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double u1 = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs2;
        sincos(2.0 * RB_PI * u2, &sn, &cs2);
        q[2 * d] = dp.nominal[2 * d] + dp.sigma[2 * d] * (rad * cs2);
        q[2 * d + 1] = dp.nominal[2 * d + 1] + dp.sigma[2 * d + 1] * (rad * sn);
    }
    if (params) {
#pragma unroll
        for (int c = 0; c < 6; ++c) params[c * n + i] = q[c];
    }
    double sg, cg, y[6];
    sincos(gmst, &sg, &cg);
    initial_state(q[0], q[1], q[2], q[3], q[4], q[5], sg, cg, y);
#pragma unroll
    for (int c = 0; c < 6; ++c) y0[i * ts + c * cs] = y[c];
}

// ---------------------------------------------------------------- diagnostics (equations_of_motion dict)
__global__ void __launch_bounds__(128) k_eom(int64_t n, const double *t, const double *y, double epoch_jd, double gmst0,
                                             const double *cd, const double *area, const double *mass, double cd_s,
                                             double area_s, double mass_s, int has_thermal, Thermal th, double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e[ENTRY_DOUBLES];
    time_entry(t[i], epoch_jd, gmst0, e);
    double r[3] = {y[0 * n + i], y[1 * n + i], y[2 * n + i]};
    double v[3] = {y[3 * n + i], y[4 * n + i], y[5 * n + i]};
    const double m_i = mass ? mass[i] : mass_s;
    const Body body = make_body(cd ? cd[i] : cd_s, area ? area[i] : area_s, m_i);
    double a[3];
    RhsOut d;
    acceleration<true>(e, r, v, body, a, &d);
    double *o = out + i;
    for (int c = 0; c < 3; ++c) {
        o[(0 + c) * n] = a[c];
        o[(3 + c) * n] = d.a_grav[c];
        o[(6 + c) * n] = d.a_j2[c];
        o[(9 + c) * n] = d.a_moon[c];
        o[(12 + c) * n] = d.a_sun[c];
        o[(15 + c) * n] = d.a_drag[c];
    }
    o[18 * n] = d.altitude;
    o[19 * n] = d.latitude_deg;
    o[20 * n] = d.rho;
    o[21 * n] = d.T;
    o[22 * n] = d.v_rel_norm;
    if (has_thermal) {
        double q[6];
        const double adn = sqrt(d.a_drag[0] * d.a_drag[0] + d.a_drag[1] * d.a_drag[1] + d.a_drag[2] * d.a_drag[2]);
        spacecraft_temperature(d.v_rel_norm, d.T, adn, m_i, th, q);
        for (int c = 0; c < 6; ++c) o[(23 + c) * n] = q[c];
    }
}

// ---------------------------------------------------------------- N2: ground-track post-pass
// app.py:225-243 (ECI -> ECEF -> geodetic per sample), :302-308 (cumulative haversine down-range),
// spacecraft_visualization.py:562-577 (altitude-threshold crossings).  One thread per trajectory walks
// its samples in time order (the down-range sum is sequential); rows are [k][field][traj] -> coalesced.
struct GroundTrackParams {
    int64_t n, n_out;
    const double *samples;
    int64_t sample_stride, field_stride;
    const int32_t *n_samples;   // valid samples per trajectory, or NULL (= n_out)
    double t0, sample_dt, gmst0, threshold;
    double *geo;                // [n_out][3][n] lat deg, lon deg, h; or NULL
    double *downrange;          // [n_out][n] or NULL
    double *cross_time;         // [n] time of the first threshold crossing (NaN none) or NULL
    double *cross_downrange;    // [n] or NULL
    int32_t *n_cross;           // [n] number of crossings or NULL
};

__global__ void __launch_bounds__(128) k_ground_track(const GroundTrackParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int ns = p.n_samples ? p.n_samples[i] : (int)p.n_out;
    double lat0 = 0, lon0 = 0, dr = 0.0, alt_prev = 0, ct = nan(""), cd = nan("");
    int nc = 0;
    for (int k = 0; k < ns; ++k) {
        const double *s = p.samples + k * p.sample_stride + i;
        const double x = s[0], y = s[p.field_stride], z = s[2 * p.field_stride], alt = s[6 * p.field_stride];
        const double t = p.t0 + p.sample_dt * k;
        double sg, cg;
        sincos(p.gmst0 + RB_EARTH_OMEGA * t, &sg, &cg);           // app.py:225
        const double xe = cg * x + sg * y, ye = cg * y - sg * x;   // eci_to_ecef
        double lat, lon, h;
        geodetic_full(xe, ye, z, lat, lon, h);
        if (k > 0) dr += haversine(lat0, lon0, lat, lon);
        if (k > 0 && ((alt >= p.threshold && alt_prev < p.threshold) || (alt <= p.threshold && alt_prev > p.threshold))) {
            if (nc == 0) { ct = t; cd = dr; }
            ++nc;
        }
        if (p.geo) {
            double *g = p.geo + (int64_t)k * 3 * p.n + i;
            g[0] = lat; g[p.n] = lon; g[2 * p.n] = h;
        }
        if (p.downrange) p.downrange[(int64_t)k * p.n + i] = dr;
        lat0 = lat; lon0 = lon; alt_prev = alt;
    }
    if (p.cross_time) p.cross_time[i] = ct;
    if (p.cross_downrange) p.cross_downrange[i] = cd;
    if (p.n_cross) p.n_cross[i] = nc;
}

// geodetic impact point (or end point) of every trajectory: final_state SoA [6][n] at time
// impact_time[i] (impacted) or t_end (others); out SoA [3][n] = lat deg, lon deg, h
__global__ void k_impact_points(int64_t n, const double *final_state, const double *impact_time, const uint8_t *status,
                                double t_end, double gmst0, double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double t = (status[i] == RB_TRAJ_IMPACTED && impact_time && impact_time[i] == impact_time[i]) ? impact_time[i] : t_end;
    double sg, cg;
    sincos(gmst0 + RB_EARTH_OMEGA * t, &sg, &cg);
    const double x = final_state[i], y = final_state[n + i], z = final_state[2 * n + i];
    double lat, lon, h;
    geodetic_full(cg * x + sg * y, cg * y - sg * x, z, lat, lon, h);
    out[i] = lat; out[n + i] = lon; out[2 * n + i] = h;
}

// ---------------------------------------------------------------- N4: the reference's integrator as shipped
// scipy.integrate.solve_ivp(method='RK45', rtol, atol, t_eval, events=[altitude_event]) as called by
// run_simulation (spacecraft_model.py:516), one thread per trajectory, every trajectory on its own time
// axis: select_initial_step (scipy _ivp/common.py:68-134), RungeKutta._step_impl (rk.py:111-176: RMS error
// norm, SAFETY 0.9, factor clamp 0.2..10, no growth after a rejection, min step 10 ulp), events and t_eval
// sampling on the quartic dense output (ivp.py:659-733).  Time-dependent inputs (Sun/Moon ephemeris, solar
// factor) are evaluated at every stage time on the device (time_entry), since no two lanes share a time grid.
// ABI table (16 doubles / entry) -> the step kernel's 96-byte entries (rb_ctx_set_time_table)
__global__ void k_hot_table(const double *__restrict__ abi, double *__restrict__ hot, int64_t n_entries) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_entries) hot_entry(abi + i * ENTRY_DOUBLES, hot + i * HOT_DOUBLES);
}

struct AdaptiveParams {
    int64_t n;
    const double *y0;
    int64_t y0_traj_stride, y0_comp_stride;
    const double *cd, *area, *mass;
    double cd_s, area_s, mass_s;
    double epoch_jd, gmst0, t0, t_bound, rtol, atol, t_eval0, t_eval_dt, event_altitude;
    int64_t n_eval, max_steps;
    double *samples;
    int64_t sample_stride, field_stride;
    double *final_state;   // SoA [6][n]: state at the event time / at t_bound
    double *impact_time;   // [n] NaN = none
    uint8_t *status;       // [n] rb_traj_status
    int32_t *n_samples, *n_accepted, *n_rejected;
    double *step_log;      // [step_capacity][7][n]: (t_old, y_old) of every accepted step, or nullptr
    int64_t step_capacity;
    // ephemeris grid (optional): hot entries at t = eph_t0 + k / eph_inv_dt, k = 0 .. eph_n - 1
    const double *eph;
    double eph_t0, eph_inv_dt;
    int64_t eph_n;
};

// Nodes of the adaptive mode's ephemeris grid: the reference's own Sun / Moon / solar-factor functions at the node
// times, in the step kernel's entry form.
__global__ void k_eph_nodes(double *eph, int64_t n_nodes, double t_first, double dt, double epoch_jd, double gmst0) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) return;
    double e[ENTRY_DOUBLES];
    time_entry(t_first + (double)k * dt, epoch_jd, gmst0, e);
    hot_entry(e, eph + k * HOT_DOUBLES);
}

// Time-only inputs of the RHS at an arbitrary time: 4-point Lagrange interpolation on the grid.  The interpolation
// error is (3/128) dt^4 |d4f/dt4|: 3e-11 m on the Moon's position for dt = 16 s (R w^4 = 1.9e-14 m/s^4), far below
// the rounding of a 3.8e8 m coordinate; the Sun, the indirect terms and the solar factor are smoother still.
__device__ __forceinline__ void eph_interpolate(const double *eph, double eph_t0, double eph_inv_dt, int64_t eph_n, double t,
                                                double hot[HOT_DOUBLES]) {
    const double u = (t - eph_t0) * eph_inv_dt;
    int64_t k = (int64_t)floor(u);
    k = k < 1 ? 1 : (k > eph_n - 3 ? eph_n - 3 : k);
    const double x = u - (double)k;               // in [0, 1] (a little outside at the clamped ends)
    const double xm = x - 1.0, xp = x + 1.0, x2 = x - 2.0;
    const double w0 = -(x * xm) * x2 * (1.0 / 6.0), w1 = (xp * xm) * x2 * 0.5, w2 = -(xp * x) * x2 * 0.5, w3 = (xp * x) * xm * (1.0 / 6.0);
    const double *e0 = eph + (k - 1) * HOT_DOUBLES;
#pragma unroll
    for (int j = 0; j < HOT_DOUBLES; ++j)
        hot[j] = fma(w0, __ldg(e0 + j), fma(w1, __ldg(e0 + HOT_DOUBLES + j), fma(w2, __ldg(e0 + 2 * HOT_DOUBLES + j), w3 * __ldg(e0 + 3 * HOT_DOUBLES + j))));
    hot[H_T] = t;
}

__device__ __forceinline__ double rms6(const double x[6]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s = fma(x[i], x[i], s);
    return sqrt(s) * 0.408248290463863 /* 1/sqrt(6) */;
}

// scipy's explicit Runge-Kutta pairs as data (rk.py: class RK45 :541-565 -> method 0, class RK23 :432-452 -> method 1)
struct RkMethod {
    double C[7], A[7][6], B[7], E[8], P[8][4];
    double err_exp, init_exp;   // -1 / (error_estimator_order + 1), 1 / (error_estimator_order + 1)
};
__constant__ RkMethod c_rk[2] = {
    {{0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1}, {RB_DP_A_ROWS},
     {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84},
     {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40}, {RB_DP_P_ROWS}, -1.0 / 5, 1.0 / 5},
    {{0, 1.0 / 2, 3.0 / 4}, {{0}, {1.0 / 2}, {0, 3.0 / 4}}, {2.0 / 9, 1.0 / 3, 4.0 / 9},
     {5.0 / 72, -1.0 / 12, -1.0 / 9, 1.0 / 8}, {{1, -4.0 / 3, 5.0 / 9}, {0, 1, -2.0 / 3}, {0, 4.0 / 3, -8.0 / 9}, {0, -1, 1}},
     -1.0 / 3, 1.0 / 3}};
template <int METHOD> struct RkShape;
template <> struct RkShape<0> { static constexpr int S = 6, NP = 4, KROWS = 7; };    // RK45: 6 stages + FSAL, quartic dense output
template <> struct RkShape<1> { static constexpr int S = 3, NP = 3, KROWS = 4; };    // RK23: 3 stages + FSAL, cubic dense output
template <> struct RkShape<2> { static constexpr int S = 12, NP = 0, KROWS = 16; };  // DOP853: 12 stages + FSAL + 3 for the dense output

// Method 2: scipy's DOP853 (rk.py class DOP853; coefficients dop853_coefficients.py -> include/rb_dop853_tables.h).  The error
// norm combines a 5th- and a 3rd-order estimate (_estimate_error_norm); the order-7 dense output (_dense_output_impl) costs three
// more RHS evaluations and solve_ivp builds it on EVERY accepted step of the reference's call, because run_simulation's
// progress_event (spacecraft_model.py:505-511) returns 0, i.e. is "active" on every step (ivp.py:677-681).
struct DopTables {
    double C[16], A[16][15], E3[13], E5[13], D[4][16];
};
__constant__ DopTables c_dop = {RB_DOP_C_INIT, RB_DOP_A_INIT, RB_DOP_E3_INIT, RB_DOP_E5_INIT, RB_DOP_D_INIT};

struct DenseDop {   // Dop853DenseOutput, rk.py:740-757
    double F[7][6];
    double y_old[6];
    double t_old, h, event_altitude;
    __device__ void eval(double t, double yo[6]) const {
        const double x = (t - t_old) / h, x1 = 1 - x;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double a = F[6][c] * x;
            a = (a + F[5][c]) * x1; a = (a + F[4][c]) * x; a = (a + F[3][c]) * x1;
            a = (a + F[2][c]) * x; a = (a + F[1][c]) * x1; a = (a + F[0][c]) * x;
            yo[c] = a + y_old[c];
        }
    }
    __device__ double g(double t) const {
        double yo[6];
        eval(t, yo);
        return event_g(yo, event_altitude);
    }
};

constexpr int ADAPTIVE_BLOCK = 64;
template <int METHOD>
constexpr size_t adaptive_smem_bytes() { return sizeof(double) * RkShape<METHOD>::KROWS * 6 * ADAPTIVE_BLOCK; }

// The stage derivatives K_0 .. K_S: a local array (default; a 672-byte stack frame, of which the compiler keeps what it can
// in its 168 registers) or, with RB_ADAPTIVE_K_SHARED = 1, shared memory ([stage][component][thread], conflict-free, 352-byte
// frame).  Measured on B200, 100 k trajectories x 1024 s: interpolated-ephemeris mode 11.4 ms local against 12.8 ms shared
// (1.79 against 1.59 G step attempts/s); per-stage ephemeris mode 95 against 93.5 ms.  Local it stays.
#ifndef RB_ADAPTIVE_K_SHARED
#define RB_ADAPTIVE_K_SHARED 0
#endif
template <int METHOD>
__global__ void __launch_bounds__(ADAPTIVE_BLOCK, 6) k_adaptive(const AdaptiveParams p) {   // 6 blocks: 168 registers (8: 128 registers and 316 B of spills)
    constexpr int S = RkShape<METHOD>::S, NP = RkShape<METHOD>::NP, KROWS = RkShape<METHOD>::KROWS;
    constexpr bool DOP = METHOD == 2;
    extern __shared__ __align__(16) double Ksh[];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= p.n) return;
    const RkMethod &M = c_rk[DOP ? 0 : METHOD];
    const double err_exp = DOP ? -1.0 / 8 : M.err_exp, init_exp = DOP ? 1.0 / 8 : M.init_exp;   // error_estimator_order 7 for DOP853
    auto coef_a = [&](int st, int j) { return DOP ? c_dop.A[st][j] : M.A[st][j]; };
    auto coef_c = [&](int st) { return DOP ? c_dop.C[st] : M.C[st]; };
    auto coef_b = [&](int j) { return DOP ? c_dop.A[12][j] : M.B[j]; };
#if RB_ADAPTIVE_K_SHARED
    double *Kt = Ksh + threadIdx.x;
#define RB_K(j, i) Kt[((j) * 6 + (i)) * ADAPTIVE_BLOCK]
#else
    double Kloc[KROWS][6];
#define RB_K(j, i) Kloc[j][i]
#endif
    const Body body = make_body(p.cd ? p.cd[tid] : p.cd_s, p.area ? p.area[tid] : p.area_s, p.mass ? p.mass[tid] : p.mass_s);
    double y[6], f[6], ynew[6], fnew[6], ent[ENTRY_DOUBLES];
    auto rhs = [&](double t, const double *yy, double *dy) {
        double hot[HOT_DOUBLES];
        if (p.eph) eph_interpolate(p.eph, p.eph_t0, p.eph_inv_dt, p.eph_n, t, hot);
        else {   // the reference's ephemeris functions at the stage time itself
            time_entry(t, p.epoch_jd, p.gmst0, ent);
            hot_entry(ent, hot);
        }
        double a[3];
        acceleration<false, false, false>(hot, yy, yy + 3, body, a, nullptr);   // (full-precision arithmetic, see rb_physics.cuh)
        dy[0] = yy[3]; dy[1] = yy[4]; dy[2] = yy[5]; dy[3] = a[0]; dy[4] = a[1]; dy[5] = a[2];
    };
#pragma unroll
    for (int c = 0; c < 6; ++c) y[c] = p.y0[tid * p.y0_traj_stride + c * p.y0_comp_stride];
    double t = p.t0;
    rhs(t, y, f);
    double h_abs;
    {   // select_initial_step (common.py:68-134), order = error_estimator_order, direction +1, max_step inf
        const double interval = fabs(p.t_bound - p.t0);
        if (interval == 0.0) h_abs = 0.0;
        else {
            double scale[6], a[6], b[6], y1[6], f1[6];
            for (int i = 0; i < 6; ++i) { scale[i] = p.atol + fabs(y[i]) * p.rtol; a[i] = y[i] / scale[i]; b[i] = f[i] / scale[i]; }
            const double d0 = rms6(a), d1 = rms6(b);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = fmin(h0, interval);
            for (int i = 0; i < 6; ++i) y1[i] = fma(h0, f[i], y[i]);
            rhs(t + h0, y1, f1);
            for (int i = 0; i < 6; ++i) a[i] = (f1[i] - f[i]) / scale[i];
            const double d2 = rms6(a) / h0;
            const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), init_exp);
            h_abs = fmin(fmin(100 * h0, h1), interval);
        }
    }
    double g_old = event_g(y, p.event_altitude);
    int64_t ei = 0;
    int32_t acc = 0, rej = 0;
    uint8_t status = RB_TRAJ_ALIVE;
    double t_event = nan("");
    while (true) {
        if (t == p.t_bound) break;
        if ((int64_t)acc + rej >= p.max_steps) { status = RB_TRAJ_STEP_FAILED; break; }
        const double min_step = 10 * fabs(nextafter(t, (double)INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        bool rejected = false, failed = false;
        double h = 0, t_new = t;
        for (;;) {   // RungeKutta._step_impl, rk.py:111-176
            if (h_abs < min_step) { failed = true; break; }
            t_new = t + h_abs;
            if (t_new - p.t_bound > 0) t_new = p.t_bound;
            h = t_new - t;
            h_abs = fabs(h);
#pragma unroll
            for (int i = 0; i < 6; ++i) RB_K(0, i) = f[i];
#pragma unroll 1
            for (int s = 1; s < S; ++s) {   // rk_step, rk.py:61-65
                double ys[6], ks[6];
                for (int i = 0; i < 6; ++i) {
                    double dy = coef_a(s, 0) * RB_K(0, i);
                    for (int j = 1; j < s; ++j) dy = fma(coef_a(s, j), RB_K(j, i), dy);
                    ys[i] = fma(dy, h, y[i]);
                }
                rhs(t + coef_c(s) * h, ys, ks);
                for (int i = 0; i < 6; ++i) RB_K(s, i) = ks[i];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double sacc = coef_b(0) * RB_K(0, i);
#pragma unroll
                for (int j = 1; j < S; ++j) sacc = fma(coef_b(j), RB_K(j, i), sacc);
                ynew[i] = fma(h, sacc, y[i]);
            }
            rhs(t + h, ynew, fnew);
            double error_norm;
            if constexpr (DOP) {   // DOP853._estimate_error_norm
                double e5n = 0.0, e3n = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    RB_K(S, i) = fnew[i];
                    double e5 = RB_K(0, i) * c_dop.E5[0], e3 = RB_K(0, i) * c_dop.E3[0];
#pragma unroll
                    for (int j = 1; j <= S; ++j) { e5 = fma(RB_K(j, i), c_dop.E5[j], e5); e3 = fma(RB_K(j, i), c_dop.E3[j], e3); }
                    const double scale = p.atol + fmax(fabs(y[i]), fabs(ynew[i])) * p.rtol;
                    e5 /= scale; e3 /= scale;
                    e5n = fma(e5, e5, e5n); e3n = fma(e3, e3, e3n);
                }
                error_norm = (e5n == 0 && e3n == 0) ? 0.0 : fabs(h) * e5n / sqrt((e5n + 0.01 * e3n) * 6.0);
            } else {
                double en[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    RB_K(S, i) = fnew[i];
                    double e = RB_K(0, i) * M.E[0];
#pragma unroll
                    for (int j = 1; j < S; ++j) e = fma(RB_K(j, i), M.E[j], e);
                    e = fma(fnew[i], M.E[S], e);
                    const double scale = p.atol + fmax(fabs(y[i]), fabs(ynew[i])) * p.rtol;
                    en[i] = e * h / scale;
                }
                error_norm = rms6(en);
            }
            if (error_norm < 1) {
                double factor = (error_norm == 0) ? 10.0 : fmin(10.0, 0.9 * pow(error_norm, err_exp));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                break;
            }
            if (!(error_norm >= 1)) { failed = true; status = RB_TRAJ_NONFINITE; break; }   // NaN error norm
            h_abs *= fmax(0.2, 0.9 * pow(error_norm, err_exp));
            rejected = true;
            ++rej;
        }
        if (failed) { if (status == RB_TRAJ_ALIVE) status = RB_TRAJ_STEP_FAILED; break; }
        if (acc < p.step_capacity) {   // progress_event's "root" and state of this step: its start
            double *sl = p.step_log + ((int64_t)acc * 7) * p.n + tid;
            sl[0] = t;
            for (int c = 0; c < 6; ++c) sl[(c + 1) * p.n] = y[c];
        }
        ++acc;
        typename std::conditional<DOP, DenseDop, Dense>::type d;
        if constexpr (DOP) {   // _dense_output_impl: stages 13..15, then F
#pragma unroll 1
            for (int s = S + 1; s < 16; ++s) {
                double ys[6], ks[6];
                for (int i = 0; i < 6; ++i) {
                    double dy = c_dop.A[s][0] * RB_K(0, i);
                    for (int j = 1; j < s; ++j) dy = fma(c_dop.A[s][j], RB_K(j, i), dy);
                    ys[i] = fma(dy, h, y[i]);
                }
                rhs(t + c_dop.C[s] * h, ys, ks);
                for (int i = 0; i < 6; ++i) RB_K(s, i) = ks[i];
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double delta = ynew[c] - y[c];
                d.F[0][c] = delta;
                d.F[1][c] = h * RB_K(0, c) - delta;
                d.F[2][c] = 2 * delta - h * (fnew[c] + RB_K(0, c));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double q = c_dop.D[k][0] * RB_K(0, c);
#pragma unroll
                    for (int j = 1; j < 16; ++j) q = fma(c_dop.D[k][j], RB_K(j, c), q);
                    d.F[3 + k][c] = h * q;
                }
                d.y_old[c] = y[c];
            }
        } else {   // dense output of the accepted step: Q = K^T P (rk.py:173), NP columns
#pragma unroll
            for (int c = 0; c < 6; ++c) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double q = 0.0;
                    if (k < NP) {
#pragma unroll
                        for (int j = 0; j <= S; ++j) q += RB_K(j, c) * M.P[j][k];
                    }
                    d.Q[c][k] = q;
                }
                d.y_old[c] = y[c];
            }
        }
        d.t_old = t; d.h = h; d.event_altitude = p.event_altitude;
        const double g_new = event_g(ynew, p.event_altitude);
        double t_lim = t_new;
        bool terminate = false;
        if (g_old >= 0.0 && g_new <= 0.0) {
            t_event = brent_event(d, t, t_new);
            terminate = true;
            t_lim = t_event;
        }
        while (ei < p.n_eval) {
            const double te = p.t_eval0 + p.t_eval_dt * (double)ei;
            if (te > t_lim) break;
            if (p.samples) {
                double ye[6];
                d.eval(te, ye);
                double *sp = p.samples + ei * p.sample_stride + tid;
                for (int c = 0; c < 6; ++c) sp[c * p.field_stride] = ye[c];
                sp[6 * p.field_stride] = sqrt(ye[0] * ye[0] + ye[1] * ye[1] + ye[2] * ye[2]) - RB_EARTH_R;
                sp[7 * p.field_stride] = sqrt(ye[3] * ye[3] + ye[4] * ye[4] + ye[5] * ye[5]);
            }
            ++ei;
        }
        if (terminate) {
            d.eval(t_event, y);
            status = RB_TRAJ_IMPACTED;
            break;
        }
        for (int c = 0; c < 6; ++c) { y[c] = ynew[c]; f[c] = fnew[c]; }
        t = t_new; g_old = g_new;
    }
#undef RB_K
    for (int c = 0; c < 6; ++c) p.final_state[c * p.n + tid] = y[c];
    if (p.impact_time) p.impact_time[tid] = t_event;
    p.status[tid] = status;
    if (p.n_samples) p.n_samples[tid] = (int32_t)ei;
    if (p.n_accepted) p.n_accepted[tid] = acc;
    if (p.n_rejected) p.n_rejected[tid] = rej;
}

// ---------------------------------------------------------------- self-test of rb_math.cuh
__global__ void k_math_probe(int op, int64_t n, const double *x, const double *y, double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = x[i], b = y ? y[i] : 0.0;
    double r;
    switch (op) {
        case 0: r = rb_rcp(a); break;
        case 1: r = rb_rsqrt(a); break;
        case 2: r = rb_exp(a); break;
        case 3: r = rb_log_near1(a); break;
        case 4: r = rb_atan2_unit(a, b); break;
        case 5: r = rb_pow_near1(a, b); break;
        case 6: r = rb_log1p_small(a); break;
        case 7: r = rb_exp_bounded(b * rb_log1p_small(a)); break;   // the pressure law as density() evaluates it
        // the budgeted variants of the hot RHS (FAST), each with its own stated bound
        case 8: r = rb_exp_bounded<TabGlobal, true>(a); break;
        case 9: r = rb_log1p_small<true>(a); break;
        case 10: r = rb_sqrt_q(a); break;
        case 11: r = third_body_k3(b, a); break;
        case 12: r = rb_rsqrt_q(a); break;
        case 13: r = rb_rcp_q(a); break;
        case 14: r = rb_atan2_unit<true>(a, b); break;
        case 15: r = rb_sqrt(a); break;
        default: r = nan("");
    }
    out[i] = r;
}

// ---------------------------------------------------------------- FP64 FMA peak probe
__global__ void __launch_bounds__(256) k_dfma_peak(double *sink, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;  // never true; keeps the chains alive
}
constexpr int DFMA_PER_THREAD_ITER = 64;

}  // namespace rb
