// rb_api.cu -- C ABI (include/reentry_b200.h) of the B200-native ensemble propagator.
// Host orchestration: context + workspace, time-table build/upload, launch sequence of the
// step kernel with live-list compaction between launches, lagged live-count polling,
// impact refinement, host-buffer convenience path with overlapped sample download.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <new>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "rb_kernels.cuh"
#include "reentry_b200.h"

using namespace rb;

// Launch shape of the step kernel: 128 threads x 6 resident blocks = 24 warps/SM at 80 registers, no spills
// (swept on B200: 256x2 5.9, 128x4 6.0, 128x5 6.7, 256x3 6.7, 128x6 6.8, 128x7 6.8 G traj-steps/s; profiles/).
#ifndef RB_BLOCK
#define RB_BLOCK 128
#endif
#ifndef RB_MIN_BLOCKS
#define RB_MIN_BLOCKS 6
#endif
// (Narrower blocks for ensembles that do not fill one wave of the machine -- 32- and 64-thread blocks with 4-step tiles,
//  which spread the warps to within one or two per SM -- were built and measured on the two sub-wave configurations of
//  BASELINE.json: C3 (100 k trajectories) 9.42 G trajectory-steps/s against 10.1 G with 128-thread blocks, C4 (50 k)
//  9.46 against 9.53 (profiles/r2e_configs_c3_c4.jsonl).  One shape it is.)
template <bool SKIP>
static void launch_step(const StepParams &sp, int64_t live_bound, cudaStream_t s) {
    const unsigned grid = (unsigned)((live_bound + RB_BLOCK - 1) / RB_BLOCK);
    k_propagate<RB_BLOCK, RB_MIN_BLOCKS, SKIP><<<grid, RB_BLOCK, propagate_smem_bytes<RB_BLOCK>(), s>>>(sp);
}

struct rb_ctx {
    int device = 0;
    char err[512] = {0};
    // time table
    double *d_table = nullptr;      // the step kernel's table (HOT_DOUBLES per entry)
    double *d_table_abi = nullptr;  // the caller's table as uploaded (ENTRY_DOUBLES per entry), converted on the device
    double *d_eph = nullptr;        // adaptive mode: ephemeris grid nodes (HOT_DOUBLES per node)
    int64_t eph_capacity = 0;
    int64_t table_capacity = 0;  // entries allocated
    int64_t table_steps = -1;    // steps covered
    double t0 = 0, dt = 0;
    // workspace
    int32_t *d_live[2] = {nullptr, nullptr};
    int32_t *d_block_counts = nullptr;
    int32_t *d_n_live = nullptr;  // per partition q: [4q] live count, [4q+1] fresh count; [8] edge-sample counter
    int64_t ws_n = 0;
    int64_t *d_edge_slot = nullptr;   // packed sample slabs: [2 n] place of the row at the end of the crossing step
    int64_t edge_n = 0;
    int64_t table_generation = 0;
    int32_t *h_poll = nullptr;  // pinned ring
    int32_t *d_poll = nullptr;  // device address of the same memory
    cudaEvent_t poll_ev[16] = {};
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr, span_ev[2] = {nullptr, nullptr};
    cudaStream_t aux_stream = nullptr;  // second partition
    bool events_ready = false;
    // host-path device buffers (grow-only)
    void *d_host_bufs[9] = {nullptr};
    size_t d_host_bytes[9] = {0};
    cudaStream_t copy_stream = nullptr;
    cudaStream_t compute_stream = nullptr;
    cudaStream_t up_stream = nullptr;   // chunked table upload of the packed host path
    double *h_table = nullptr;  // pinned staging for the table
    int64_t h_table_entries = 0;
    rb_stats stats = {0, 0, 0, -1, 0.0, 0};
};

static int32_t cuda_fail(rb_ctx *ctx, cudaError_t e, const char *what) {
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, cudaGetErrorString(e));
    return RB_ERR_CUDA;
}
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);     \
    } while (0)

extern "C" {

int32_t rb_abi_version(void) { return RB_ABI_VERSION; }

const char *rb_strerror(int32_t code) {
    switch (code) {
        case RB_OK: return "ok";
        case RB_ERR_INVALID_ARG: return "invalid argument";
        case RB_ERR_CUDA: return "CUDA error (see rb_ctx_last_error)";
        case RB_ERR_NO_TIME_TABLE: return "no time table staged (rb_ctx_set_time_table)";
        case RB_ERR_TABLE_RANGE: return "run exceeds the staged time table";
        case RB_ERR_NOMEM: return "out of memory";
        case RB_ERR_NO_DEVICE: return "no CUDA device";
        case RB_ERR_SAMPLE_CAPACITY: return "packed sample arena or slab list too small";
        default: return "unknown error";
    }
}

const char *rb_ctx_last_error(const rb_ctx *ctx) { return ctx ? ctx->err : ""; }

int32_t rb_ctx_create(rb_ctx **out, int32_t device) {
    if (!out) return RB_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return RB_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return RB_ERR_INVALID_ARG;
    rb_ctx *ctx = new (std::nothrow) rb_ctx();
    if (!ctx) return RB_ERR_NOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_n_live, 12 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_poll, 16 * sizeof(int32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&ctx->d_poll, ctx->h_poll, 0);
    for (int i = 0; i < 16 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->poll_ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->span_ev[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->span_ev[1]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        ctx->events_ready = true;
        e = cudaFuncSetAttribute(k_propagate<RB_BLOCK, RB_MIN_BLOCKS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)propagate_smem_bytes<RB_BLOCK>());
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_propagate<RB_BLOCK, RB_MIN_BLOCKS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)propagate_smem_bytes<RB_BLOCK>());
    }
    if (e != cudaSuccess) {
        fprintf(stderr, "rb_ctx_create: %s\n", cudaGetErrorString(e));
        delete ctx;
        return RB_ERR_CUDA;
    }
    *out = ctx;
    return RB_OK;
}

int32_t rb_ctx_destroy(rb_ctx *ctx) {
    if (!ctx) return RB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_table);
    cudaFree(ctx->d_table_abi);
    cudaFree(ctx->d_eph);
    cudaFree(ctx->d_live[0]);
    cudaFree(ctx->d_live[1]);
    cudaFree(ctx->d_block_counts);
    cudaFree(ctx->d_edge_slot);
    cudaFree(ctx->d_n_live);
    cudaFreeHost(ctx->h_poll);
    cudaFreeHost(ctx->h_table);
    for (auto &b : ctx->d_host_bufs) cudaFree(b);
    if (ctx->events_ready)
        for (auto &ev : ctx->poll_ev) cudaEventDestroy(ev);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
    for (auto &ev : ctx->span_ev) if (ev) cudaEventDestroy(ev);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->compute_stream) cudaStreamDestroy(ctx->compute_stream);
    if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
    delete ctx;
    return RB_OK;
}

int64_t rb_ctx_table_generation(const rb_ctx *ctx) { return ctx ? ctx->table_generation : -1; }

int32_t rb_ctx_stats(const rb_ctx *ctx, rb_stats *out) {
    if (!ctx || !out) return RB_ERR_INVALID_ARG;
    *out = ctx->stats;
    return RB_OK;
}

// ------------------------------------------------------------------ time table
int64_t rb_time_table_entries(int64_t n_steps) { return n_steps < 0 ? 0 : RB_STAGES_PER_STEP * n_steps + 1; }

int32_t rb_build_time_table(double epoch_jd, double gmst0, double t0, double dt, int64_t n_steps, double *host_table,
                            int32_t n_threads) {
    if (!host_table || n_steps < 0 || !(dt > 0)) return RB_ERR_INVALID_ARG;
    // stage times exactly as scipy forms them: t_{n+1} = t_n + h (rk.py t_new = t + h), t_n + C_s*h (rk.py:64)
    const double C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0};
    std::vector<double> tn((size_t)n_steps + 1);
    double t = t0;
    for (int64_t n = 0; n <= n_steps; ++n) { tn[(size_t)n] = t; t = t + dt; }
    const int64_t n_entries = rb_time_table_entries(n_steps);
#ifdef _OPENMP
    // torchrun exports OMP_NUM_THREADS=1 to every rank; the table build is a short burst, so size the team from
    // the machine (capped) rather than from that variable
    if (n_threads <= 0) n_threads = std::max(omp_get_max_threads(), std::min(omp_get_num_procs(), 16));
#pragma omp parallel for schedule(static) num_threads(n_threads) if (n_entries > 512)
#endif
    for (int64_t e = 0; e < n_entries; ++e) {
        double te;
        if (e == 0) te = tn[0];
        else {
            const int64_t n = (e - 1) / RB_STAGES_PER_STEP;
            const int s = (int)((e - 1) % RB_STAGES_PER_STEP) + 1;
            te = (s == 5) ? tn[(size_t)n + 1] : tn[(size_t)n] + C[s] * dt;
        }
        time_entry(te, epoch_jd, gmst0, host_table + e * ENTRY_DOUBLES);
    }
    (void)n_threads;
    return RB_OK;
}

int32_t rb_ctx_set_time_table(rb_ctx *ctx, const double *host_table, int64_t n_steps, double t0, double dt,
                              void *stream) {
    if (!ctx || !host_table || n_steps < 0) return RB_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    const int64_t n_entries = rb_time_table_entries(n_steps);
    if (n_entries > ctx->table_capacity) {
        CK(cudaStreamSynchronize((cudaStream_t)stream));
        cudaFree(ctx->d_table);
        cudaFree(ctx->d_table_abi);
        ctx->d_table = ctx->d_table_abi = nullptr;
        ctx->table_capacity = 0;
        CK(cudaMalloc(&ctx->d_table, (size_t)n_entries * HOT_DOUBLES * sizeof(double)));
        CK(cudaMalloc(&ctx->d_table_abi, (size_t)n_entries * ENTRY_DOUBLES * sizeof(double)));
        ctx->table_capacity = n_entries;
    }
    CK(cudaMemcpyAsync(ctx->d_table_abi, host_table, (size_t)n_entries * ENTRY_DOUBLES * sizeof(double),
                       cudaMemcpyHostToDevice, (cudaStream_t)stream));
    k_hot_table<<<(unsigned)((n_entries + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctx->d_table_abi, ctx->d_table, n_entries);
    CK(cudaGetLastError());
    ctx->table_steps = n_steps;
    ctx->t0 = t0;
    ctx->dt = dt;
    ctx->table_generation++;
    return RB_OK;
}

// ------------------------------------------------------------------ initial states
int32_t rb_initial_states(rb_ctx *ctx, int64_t n, const double *v, const double *lat, const double *lon,
                          const double *alt, const double *azimuth, const double *gamma, double gmst, double *y0,
                          int64_t y0_traj_stride, int64_t y0_comp_stride, void *stream) {
    if (!ctx || n < 0 || !v || !lat || !lon || !alt || !azimuth || !gamma || !y0) return RB_ERR_INVALID_ARG;
    if (n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    const int B = 256;
    k_initial_states<<<(unsigned)((n + B - 1) / B), B, 0, (cudaStream_t)stream>>>(n, v, lat, lon, alt, azimuth, gamma, gmst,
                                                                                 y0, y0_traj_stride, y0_comp_stride);
    CK(cudaGetLastError());
    return RB_OK;
}

int32_t rb_dispersed_initial_states(rb_ctx *ctx, int64_t n, int64_t first_index, uint64_t seed, const double *nominal6,
                                    const double *sigma6, double gmst, double *params, double *y0,
                                    int64_t y0_traj_stride, int64_t y0_comp_stride, void *stream) {
    if (!ctx || n < 0 || first_index < 0 || !nominal6 || !sigma6 || !y0) return RB_ERR_INVALID_ARG;
    if (n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    DispersionParams dp;
    for (int c = 0; c < 6; ++c) { dp.nominal[c] = nominal6[c]; dp.sigma[c] = sigma6[c]; }
    const int B = 256;
    k_dispersed_states<<<(unsigned)((n + B - 1) / B), B, 0, (cudaStream_t)stream>>>(n, first_index, seed, dp, gmst, params, y0,
                                                                                   y0_traj_stride, y0_comp_stride);
    CK(cudaGetLastError());
    return RB_OK;
}

int32_t rb_equations_of_motion(rb_ctx *ctx, int64_t n, const double *t, const double *y, double epoch_jd, double gmst0,
                               const double *cd, const double *area, const double *mass, double cd_scalar,
                               double area_scalar, double mass_scalar, const rb_thermal *thermal, double *out,
                               void *stream) {
    if (!ctx || n < 0 || !t || !y || !out) return RB_ERR_INVALID_ARG;
    if (n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    Thermal th{};
    if (thermal) {
        if (!(thermal->iter_fact > 0)) return RB_ERR_INVALID_ARG;
        th = Thermal{thermal->capsule_length, thermal->dt, thermal->thermal_conductivity, thermal->specific_heat_capacity,
                     thermal->emissivity, thermal->ablation_efficiency, thermal->iter_fact};
    }
    k_eom<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, t, y, epoch_jd, gmst0, cd, area, mass, cd_scalar,
                                                                         area_scalar, mass_scalar, thermal ? 1 : 0, th, out);
    CK(cudaGetLastError());
    return RB_OK;
}

// ------------------------------------------------------------------ propagation
static int32_t ensure_workspace(rb_ctx *ctx, int64_t n, cudaStream_t stream) {
    if (n <= ctx->ws_n) return RB_OK;
    CK(cudaStreamSynchronize(stream));
    cudaFree(ctx->d_live[0]); cudaFree(ctx->d_live[1]); cudaFree(ctx->d_block_counts);
    ctx->d_live[0] = ctx->d_live[1] = ctx->d_block_counts = nullptr;
    ctx->ws_n = 0;
    CK(cudaMalloc(&ctx->d_live[0], (size_t)n * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->d_live[1], (size_t)n * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->d_block_counts, (size_t)((n + COMPACT_BLOCK - 1) / COMPACT_BLOCK + 8) * sizeof(int32_t)));
    ctx->ws_n = n;
    return RB_OK;
}

// Called around the launches of a run (host-buffer paths).
struct LaunchHook {
    // before the launch covering table steps [.., step_end) is enqueued; streams = the partition streams it will use
    virtual int32_t before_launch(int64_t step_end, const cudaStream_t *streams, int n_streams) { (void)step_end; (void)streams; (void)n_streams; return RB_OK; }
    // after the launch that completed `steps_done` steps of the run has been enqueued on all partition streams
    virtual void after_launch(int64_t steps_done, const cudaStream_t *streams, int n_streams) { (void)steps_done; (void)streams; (void)n_streams; }
    // packed slabs: slab `index` of partition q, written by launch number `launch` (-1: the initial-state slab of k_init),
    // has just been enqueued on `stream` (its kernel is the last thing in the stream)
    virtual void after_slab(int64_t index, rb_sample_slab &slab, int q, int64_t launch, cudaStream_t stream) { (void)index; (void)slab; (void)q; (void)launch; (void)stream; }
    // the host has just learnt (lagged poll) how many trajectories of partition q were alive when launch `launch` started
    virtual void live_count(int q, int64_t launch, int64_t count) { (void)q; (void)launch; (void)count; }
    // no more launches will be enqueued
    virtual void finish_slabs() {}
    virtual ~LaunchHook() {}
};

// rb_propagate_host with a pinned sample buffer.  Sample rows reach the host two ways:
//   * while at least RB_DMA_MIN_LIVE_PERCENT of a partition's trajectories are alive, the step kernel writes the rows
//     into a device staging buffer (NaN-prefilled) and the copy engine moves each finished row slice to the host:
//     the SMs never touch the link, and the copy engine sustains the full 57 GB/s beside the kernel;
//   * below that, rows are mostly padding, so the kernel stores the few valid entries straight into the host buffer
//     (zero-copy).  Those stores are scattered 8-byte PCIe writes that the issuing warps wait on: used for the whole
//     run they cost 16 ms on C2 (167.6 ms), all of it in the phase where trajectories retire; the copy engine for every
//     row costs 184 ms (7.7 GB of padded rows).  Switch point swept on C2: 60 % 163.2, 40 % 161.0, 20 % 158.1,
//     10 % 158.2, 5 % 160.2 ms per pass.
#ifndef RB_DMA_MIN_LIVE_PERCENT
#define RB_DMA_MIN_LIVE_PERCENT 20
#endif
struct SampleRoute {
    double *staging = nullptr;   // device, same [row][field][trajectory] layout as the host buffer
    double *host = nullptr;      // the caller's pinned buffer (host address)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};   // per partition
    int64_t rows_via_dma = 0;
    // back-pressure: done[q][i & 7] completes when the i-th row slice of partition q has reached the host
    cudaEvent_t done[2][8] = {};
    int64_t n_dma[2] = {0, 0};
    int64_t launches_backlogged = 0;
};
// The copy engine is used only while the link keeps up: if the slice sent RB_DMA_MAX_BACKLOG launches ago has not
// arrived, the link (or the host's ingest: 8 ranks of a VM share ~88 GB/s) is the bottleneck, and padded rows would
// only add bytes to it -- the launch stores its valid samples directly instead.
#ifndef RB_DMA_MAX_BACKLOG
#define RB_DMA_MAX_BACKLOG 4
#endif

__global__ void k_fill_nan_2d(double *p, int64_t width, int64_t height, int64_t pitch) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < width * height) p[(i / width) * pitch + (i % width)] = nan("");
}

// One partition of the ensemble: its own live list window, counters and stream.  Two partitions on
// two streams let the ragged tail of one partition's launch (last partial wave, retired lanes)
// overlap with the body of the other's (+3.6 % on C2, profiles/r1_variant_sweep.txt).
struct Part {
    int64_t start = 0, count = 0, live_bound = 0;
    int cur = 0;
    bool finished = false;
    int64_t seen_live = -1;          // live count of the last poll (lagged host view); -1 = nothing retired yet
    cudaStream_t s = nullptr;
    int32_t *n_live = nullptr;       // device [2]: live count, fresh count
    int32_t *block_counts = nullptr; // device, this partition's slice
    int32_t *h_poll = nullptr;       // pinned ring [8]
    int32_t *d_poll = nullptr;       // its device mapping
    cudaEvent_t *poll_ev = nullptr;  // [8]
};

__global__ void k_set_counts(int32_t *a, int32_t va, int32_t *b, int32_t vb) {
    a[0] = va; a[1] = va;
    if (b) { b[0] = vb; b[1] = vb; }
}

static int32_t propagate_impl(rb_ctx *ctx, const rb_in *in, const rb_run *run, const rb_out *out, cudaStream_t stream,
                              LaunchHook *hook, int64_t *steps_done_out, SampleRoute *route = nullptr) {
    if (!ctx || !in || !run || !out) return RB_ERR_INVALID_ARG;
    const int64_t n = in->n;
    if (n < 0 || n > 0x7fffffffLL) return RB_ERR_INVALID_ARG;
    const bool resume = (run->flags & RB_FLAG_RESUME) != 0;
    const bool packed = (run->flags & RB_FLAG_PACKED_SAMPLES) != 0 && out->samples;
    if (run->n_steps < 0 || run->first_step < 0 || run->out_stride < 1 || run->out_stride > 0x7fffffffLL)
        return RB_ERR_INVALID_ARG;
    if ((!in->y0 && !resume) || !out->final_state || !out->status || !out->impact_step) return RB_ERR_INVALID_ARG;
    if (resume && (run->first_step % run->out_stride) != 0) return RB_ERR_INVALID_ARG;
    if (packed && (resume || !out->slabs || !out->n_slabs || out->slab_capacity < 1 || out->packed_capacity < 0)) return RB_ERR_INVALID_ARG;
    if (ctx->table_steps < 0) return RB_ERR_NO_TIME_TABLE;
    if (run->first_step + run->n_steps > ctx->table_steps) return RB_ERR_TABLE_RANGE;
    // sample frame: a resumed run indexes samples by the ABSOLUTE (table-frame) step, so that all chunks fill one buffer
    const int64_t frame0 = (run->flags & RB_FLAG_RESUME) ? run->first_step : 0;
    if (out->samples && !packed && out->n_out_capacity < (frame0 + run->n_steps) / run->out_stride + 1) return RB_ERR_INVALID_ARG;
    if (packed) *out->n_slabs = 0;
    ctx->stats = rb_stats{0, 0, 0, -1, 0.0, 0};
    if (steps_done_out) *steps_done_out = 0;
    if (n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    int32_t rc = ensure_workspace(ctx, n, stream);
    if (rc != RB_OK) return rc;
    int64_t packed_used = 0, n_slabs = 0;
    if (packed) {
        if (n > ctx->edge_n) {
            CK(cudaStreamSynchronize(stream));
            cudaFree(ctx->d_edge_slot);
            ctx->d_edge_slot = nullptr; ctx->edge_n = 0;
            CK(cudaMalloc(&ctx->d_edge_slot, (size_t)n * 2 * sizeof(int64_t)));
            ctx->edge_n = n;
        }
        // slab 0: the initial states (sample 0) of the whole ensemble, written by k_init
        const int64_t W0 = (n + 15) & ~(int64_t)15;
        if (RB_SAMPLE_FIELDS * W0 > out->packed_capacity) return RB_ERR_SAMPLE_CAPACITY;
        out->slabs[0] = rb_sample_slab{0, 0, 1, 0, n, W0, -1, n};
        packed_used = RB_SAMPLE_FIELDS * W0;
        n_slabs = 1;
    }

    const int64_t spl = run->steps_per_launch > 0 ? run->steps_per_launch : 32;
    const bool compact = (run->flags & RB_FLAG_COMPACT) != 0;
    const bool early = (run->flags & RB_FLAG_EARLY_EXIT) != 0;
    const bool profile = (run->flags & RB_FLAG_PROFILE) != 0;
    const bool skip_drag = (run->flags & RB_FLAG_SKIP_NEGLIGIBLE_DRAG) != 0;

    // partitions: two when each half still fills the machine several times over
    const int n_parts = (n >= RB_OVERLAP_MIN_TRAJECTORIES && !(run->flags & RB_FLAG_NO_OVERLAP)) ? 2 : 1;
    Part parts[2];
    for (int q = 0; q < n_parts; ++q) {
        Part &pt = parts[q];
        pt.start = (q == 0) ? 0 : n / 2;
        pt.count = (n_parts == 1) ? n : (q == 0 ? n / 2 : n - n / 2);
        pt.live_bound = pt.count;
        pt.s = (q == 0) ? stream : ctx->aux_stream;
        pt.n_live = ctx->d_n_live + 4 * q;
        pt.block_counts = ctx->d_block_counts + ((q == 0) ? 0 : (n / 2 + COMPACT_BLOCK - 1) / COMPACT_BLOCK + 1);
        pt.h_poll = ctx->h_poll + 8 * q;
        pt.d_poll = ctx->d_poll + 8 * q;
        pt.poll_ev = ctx->poll_ev + 8 * q;
    }

    InitParams ip{};
    ip.y0 = in->y0; ip.y0_traj_stride = in->y0_traj_stride; ip.y0_comp_stride = in->y0_comp_stride;
    ip.state = out->final_state; ip.n = n; ip.live = ctx->d_live[0];
    ip.impact_step = out->impact_step; ip.impact_time = out->impact_time; ip.status = out->status;
    ip.samples = out->samples; ip.sample_stride = out->sample_stride; ip.field_stride = out->field_stride;
    if (packed) { ip.sample_stride = RB_SAMPLE_FIELDS * out->slabs[0].width; ip.field_stride = out->slabs[0].width; }
    if (route && out->samples && !resume) ip.samples = route->staging;   // row 0 travels by copy engine too
    if (resume) k_init_resume<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ctx->d_live[0], n);
    else k_init<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ip);
    if (route && out->samples && !resume) {
        CK(cudaEventRecord(route->ev[0], stream));
        CK(cudaStreamWaitEvent(route->copy_stream, route->ev[0], 0));
        CK(cudaMemcpyAsync(route->host, route->staging, (size_t)RB_SAMPLE_FIELDS * n * sizeof(double), cudaMemcpyDeviceToHost,
                           route->copy_stream));
        route->rows_via_dma += 1;
    }
    k_set_counts<<<1, 1, 0, stream>>>(parts[0].n_live, (int32_t)parts[0].count, n_parts > 1 ? parts[1].n_live : nullptr,
                                     n_parts > 1 ? (int32_t)parts[1].count : 0);
    CK(cudaMemsetAsync(ctx->d_n_live + 8, 0, sizeof(int32_t), stream));  // edge-sample counter of k_refine
    CK(cudaGetLastError());
    ctx->stats.kernel_launches += 2;
    if (packed && hook) hook->after_slab(0, out->slabs[0], 0, -1, stream);
    if (profile) CK(cudaEventRecord(ctx->span_ev[0], stream));
    if (n_parts > 1) {  // fork
        CK(cudaEventRecord(ctx->fork_ev, stream));
        CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_ev, 0));
    }

    StepParams sp{};
    sp.table = ctx->d_table; sp.out_stride = run->out_stride; sp.h = ctx->dt; sp.ht = make_htableau(ctx->dt); sp.event_altitude = run->event_altitude;
    sp.state = out->final_state; sp.n = n;
    sp.cd = in->cd; sp.area = in->area; sp.mass = in->mass;
    sp.cd_s = in->cd_scalar; sp.area_s = in->area_scalar; sp.mass_s = in->mass_scalar;
    sp.samples = out->samples; sp.sample_stride = out->sample_stride; sp.field_stride = out->field_stride;
    sp.impact_step = out->impact_step; sp.status = out->status;
    sp.packed = packed ? 1 : 0; sp.edge_slot = packed ? ctx->d_edge_slot : nullptr;

    int64_t done = 0;
    int64_t launch_idx = 0;
    while (done < run->n_steps) {
        if (early && launch_idx >= 2) {  // look at the live counts published 2 launches ago (never blocks the GPU)
            const int slot = (int)((launch_idx - 2) & 7);
            int64_t seen_total = 0;
            for (int q = 0; q < n_parts; ++q) {
                Part &pt = parts[q];
                if (pt.finished) continue;
                CK(cudaEventSynchronize(pt.poll_ev[slot]));
                const int32_t seen = pt.h_poll[slot];
                if (compact) pt.live_bound = std::min<int64_t>(pt.live_bound, seen);  // the list only shrinks when compacted
                pt.seen_live = seen;
                if (hook) hook->live_count(q, launch_idx - 1, seen);   // the count after launch idx-2 = alive at the start of idx-1
                if (seen == 0) pt.finished = true;
                seen_total += seen;
            }
            ctx->stats.last_polled_live = seen_total;
            if (seen_total == 0) break;
        }
        const int64_t steps = std::min<int64_t>(spl, run->n_steps - done);
        sp.table_step0 = run->first_step + done;
        sp.run_step0 = frame0 + done;
        sp.n_launch_steps = (int)steps;
        cudaStream_t active[2];
        int n_active = 0;
        for (int q = 0; q < n_parts; ++q)
            if (!parts[q].finished) active[n_active++] = parts[q].s;
        if (hook && (rc = hook->before_launch(run->first_step + done + steps, active, n_active)) != RB_OK) break;
        for (int q = 0; q < n_parts && rc == RB_OK; ++q) {
            Part &pt = parts[q];
            if (pt.finished) continue;
            sp.live = ctx->d_live[pt.cur] + pt.start;
            sp.n_live = pt.n_live;
            // sample rows this launch writes: k with k * out_stride in (run_step0, run_step0 + steps]
            const int64_t k_first = sp.run_step0 / run->out_stride + 1, k_last = (sp.run_step0 + steps) / run->out_stride;
            // (once back-pressure has been seen the link is the bottleneck of this call: from then on only rows without
            //  padding -- every trajectory of the partition alive -- are worth sending by copy engine)
            const int64_t min_pct = (route && route->launches_backlogged > 0) ? 100 : RB_DMA_MIN_LIVE_PERCENT;
            const bool dense_rows = pt.seen_live < 0 || pt.seen_live * 100 >= pt.count * min_pct;
            bool via_dma = route && out->samples && dense_rows && early && k_last >= k_first;
            if (via_dma && route->n_dma[q] >= RB_DMA_MAX_BACKLOG &&
                cudaEventQuery(route->done[q][(route->n_dma[q] - RB_DMA_MAX_BACKLOG) & 7]) == cudaErrorNotReady) {
                via_dma = false;
                route->launches_backlogged++;
            }
            const size_t slice = (size_t)(k_first * RB_SAMPLE_FIELDS * n + pt.start);   // first element of the partition's slice
            sp.samples = out->samples;
            sp.k_first = k_first; sp.k_last = k_last;
            const bool slab = packed && k_last >= k_first;
            if (packed) {
                sp.samples = nullptr;   // a launch without a sample step writes nothing
                if (slab) {
                    const int64_t W = (pt.live_bound + 15) & ~(int64_t)15;
                    const int64_t need = (k_last - k_first + 1) * RB_SAMPLE_FIELDS * W;
                    if (packed_used + need > out->packed_capacity || n_slabs >= out->slab_capacity) { rc = RB_ERR_SAMPLE_CAPACITY; break; }
                    out->slabs[n_slabs] = rb_sample_slab{packed_used, k_first, k_last - k_first + 1, pt.start, pt.count, W, sp.run_step0, -1};
                    // biased so that row k of the run is at samples + k * sample_stride (integer arithmetic: the biased
                    // address may lie below the arena)
                    sp.samples = (double *)((uintptr_t)(out->samples + packed_used) - (uintptr_t)(k_first * RB_SAMPLE_FIELDS * W) * sizeof(double));
                    sp.sample_stride = RB_SAMPLE_FIELDS * W; sp.field_stride = W;
                    sp.sample_row0 = k_first; sp.slab_offset = packed_used;
                    sp.slab_live = out->slab_live ? out->slab_live + n_slabs : nullptr;
                    packed_used += need;
                }
            }
            if (via_dma) {   // (NaN first: a trajectory that retired in the last two launches leaves its entries unset)
                const int64_t hgt = (k_last - k_first + 1) * RB_SAMPLE_FIELDS;
                k_fill_nan_2d<<<(unsigned)((pt.count * hgt + 255) / 256), 256, 0, pt.s>>>(route->staging + slice, pt.count, hgt, n);
                sp.samples = route->staging;
            }
            if (skip_drag) launch_step<true>(sp, pt.live_bound, pt.s);
            else launch_step<false>(sp, pt.live_bound, pt.s);
            CK(cudaGetLastError());
            if (slab) {
                if (hook) hook->after_slab(n_slabs, out->slabs[n_slabs], q, launch_idx, pt.s);
                ++n_slabs;
            }
            if (via_dma) {
                CK(cudaEventRecord(route->ev[q], pt.s));
                CK(cudaStreamWaitEvent(route->copy_stream, route->ev[q], 0));
                CK(cudaMemcpy2DAsync(route->host + slice, (size_t)n * sizeof(double), route->staging + slice, (size_t)n * sizeof(double),
                                     (size_t)pt.count * sizeof(double), (size_t)((k_last - k_first + 1) * RB_SAMPLE_FIELDS),
                                     cudaMemcpyDeviceToHost, route->copy_stream));
                CK(cudaEventRecord(route->done[q][route->n_dma[q] & 7], route->copy_stream));
                route->n_dma[q]++;
                route->rows_via_dma += k_last - k_first + 1;
            }
            ctx->stats.kernel_launches++;
            ctx->stats.propagate_launches++;
            if ((compact || early) && done + steps < run->n_steps) {
                const unsigned cgrid = (unsigned)((pt.live_bound + COMPACT_BLOCK - 1) / COMPACT_BLOCK);
                const int32_t *lin = ctx->d_live[pt.cur] + pt.start;
                k_count_live<<<cgrid, COMPACT_BLOCK, 0, pt.s>>>(lin, pt.n_live, out->status, pt.block_counts);
                const int slot = (int)(launch_idx & 7);
                k_scan_blocks<<<1, 1024, 0, pt.s>>>(pt.block_counts, (int)cgrid, pt.n_live + 1, early ? pt.d_poll + slot : nullptr);
                ctx->stats.kernel_launches += 2;
                if (compact) {
                    k_scatter_live<<<cgrid, COMPACT_BLOCK, 0, pt.s>>>(lin, pt.n_live, out->status, pt.block_counts,
                                                                     ctx->d_live[pt.cur ^ 1] + pt.start);
                    k_copy_i32<<<1, 1, 0, pt.s>>>(pt.n_live, pt.n_live + 1);
                    ctx->stats.kernel_launches += 2;
                    pt.cur ^= 1;
                }
                CK(cudaGetLastError());
                if (early) CK(cudaEventRecord(pt.poll_ev[slot], pt.s));
            }
        }
        if (rc != RB_OK) break;
        ctx->stats.steps_enqueued += steps;
        done += steps;
        if (hook) hook->after_launch(done, active, n_active);
        ++launch_idx;
    }
    if (packed) *out->n_slabs = n_slabs;
    if (hook) hook->finish_slabs();
    if (n_parts > 1) {  // join
        CK(cudaEventRecord(ctx->join_ev, ctx->aux_stream));
        CK(cudaStreamWaitEvent(stream, ctx->join_ev, 0));
    }
    if (profile) CK(cudaEventRecord(ctx->span_ev[1], stream));
    if (rc != RB_OK) return rc;   // (streams joined; the outputs are incomplete)
    if (route && route->rows_via_dma > 0) {   // k_refine may add an event-time sample to a row the copy engine delivered
        CK(cudaEventRecord(route->ev[0], route->copy_stream));
        CK(cudaStreamWaitEvent(stream, route->ev[0], 0));
    }

    RefineParams rp{};
    rp.table = ctx->d_table; rp.table_steps = ctx->table_steps; rp.h = ctx->dt; rp.ht = make_htableau(ctx->dt); rp.event_altitude = run->event_altitude;
    rp.skip_drag = skip_drag ? 1 : 0;
    rp.state = out->final_state; rp.n = n;
    rp.cd = in->cd; rp.area = in->area; rp.mass = in->mass;
    rp.cd_s = in->cd_scalar; rp.area_s = in->area_scalar; rp.mass_s = in->mass_scalar;
    rp.impact_step = out->impact_step; rp.impact_time = out->impact_time; rp.status = out->status;
    rp.n_samples = out->n_samples;
    rp.samples = out->samples; rp.sample_stride = out->sample_stride; rp.field_stride = out->field_stride;
    rp.out_stride = run->out_stride; rp.run_first_step = run->first_step - frame0; rp.run_n_steps = frame0 + run->n_steps;
    rp.refine = (run->flags & RB_FLAG_NO_REFINE) ? 0 : 1;
    rp.edge_count = ctx->d_n_live + 8;
    rp.edge_slot = packed ? ctx->d_edge_slot : nullptr;
    k_refine<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(rp);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    if (steps_done_out) *steps_done_out = done;
    if (profile) {
        CK(cudaStreamSynchronize(stream));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, ctx->span_ev[0], ctx->span_ev[1]));
        ctx->stats.propagate_ms = t;  // span of the step-kernel launch sequence (compaction included, < 0.5 %)
    }
    return RB_OK;
}

int32_t rb_propagate(rb_ctx *ctx, const rb_in *in, const rb_run *run, const rb_out *out, void *stream) {
    return propagate_impl(ctx, in, run, out, (cudaStream_t)stream, nullptr, nullptr);
}

// ------------------------------------------------------------------ host-buffer path
struct EventSet {   // events of one call, destroyed on every exit path
    std::vector<cudaEvent_t> evs;
    cudaError_t make(cudaEvent_t *e) {
        cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
        if (r == cudaSuccess) evs.push_back(*e);
        return r;
    }
    ~EventSet() { for (auto e : evs) cudaEventDestroy(e); }
};

static int32_t host_buf(rb_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes > ctx->d_host_bytes[slot]) {
        cudaFree(ctx->d_host_bufs[slot]);
        ctx->d_host_bufs[slot] = nullptr;
        ctx->d_host_bytes[slot] = 0;
        CK(cudaMalloc(&ctx->d_host_bufs[slot], bytes));
        ctx->d_host_bytes[slot] = bytes;
    }
    *out = ctx->d_host_bufs[slot];
    return RB_OK;
}

__global__ void k_fill_nan(double *p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = nan("");
}

// transposes SoA [6][n] -> AoS [n][6] for the host-facing final_state
__global__ void k_soa_to_aos6(const double *soa, double *aos, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 6; ++c) aos[i * 6 + c] = soa[c * n + i];
}

struct SampleDownloader : LaunchHook {
    rb_ctx *ctx;
    const double *d_samples;
    double *h_samples;
    int64_t n, out_stride, rows_copied = 0;
    cudaStream_t copy_stream;
    cudaEvent_t ev[2];
    cudaError_t err = cudaSuccess;
    void after_launch(int64_t steps_done, const cudaStream_t *streams, int n_streams) override {
        // rows k <= steps_done/out_stride are final once this launch has finished on every partition stream
        const int64_t rows_ready = steps_done / out_stride + 1;
        if (rows_ready <= rows_copied) return;
        cudaError_t e = cudaSuccess;
        for (int q = 0; q < n_streams && e == cudaSuccess; ++q) {
            e = cudaEventRecord(ev[q], streams[q]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(copy_stream, ev[q], 0);
        }
        const size_t row = (size_t)RB_SAMPLE_FIELDS * n * sizeof(double);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync((char *)h_samples + rows_copied * row, (const char *)d_samples + rows_copied * row,
                                (size_t)(rows_ready - rows_copied) * row, cudaMemcpyDeviceToHost, copy_stream);
        if (e != cudaSuccess) err = e;
        rows_copied = rows_ready;
    }
};

int32_t rb_propagate_host(rb_ctx *ctx, int64_t n, const double *y0, const double *cd, const double *area,
                          const double *mass, double cd_scalar, double area_scalar, double mass_scalar, double epoch_jd,
                          double gmst0, double t0, double dt, const rb_run *run, double *samples, int64_t n_out_capacity,
                          double *final_state, int32_t *impact_step, double *impact_time, uint8_t *status,
                          int32_t *n_samples, int64_t *n_samples_used) {
    if (!ctx || !run || n < 0 || !y0 || !final_state || !impact_step || !impact_time || !status) return RB_ERR_INVALID_ARG;
    if (run->first_step != 0 || run->n_steps < 0 || run->out_stride < 1 || !(dt > 0)) return RB_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->compute_stream) CK(cudaStreamCreateWithFlags(&ctx->compute_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->compute_stream;
    const int64_t n_out = run->n_steps / run->out_stride + 1;
    if (samples && n_out_capacity < n_out) return RB_ERR_INVALID_ARG;
    if (n == 0) { if (n_samples_used) *n_samples_used = 0; return RB_OK; }

    // time table: build on the host (all cores), stage through pinned memory
    const int64_t n_entries = rb_time_table_entries(run->n_steps);
    if (n_entries > ctx->h_table_entries) {
        cudaFreeHost(ctx->h_table);
        ctx->h_table = nullptr; ctx->h_table_entries = 0;
        CK(cudaHostAlloc(&ctx->h_table, (size_t)n_entries * ENTRY_DOUBLES * sizeof(double), cudaHostAllocDefault));
        ctx->h_table_entries = n_entries;
    }
    const bool trace = getenv("RB_HOST_TRACE") != nullptr;   // phase timings of this call on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(now() - a).count(); };
    const auto t_begin = now();
    CK(cudaStreamSynchronize(cs));  // previous call may still read the pinned table
    int32_t rc = RB_OK;

    void *d_y0, *d_state, *d_cam = nullptr, *d_istep, *d_itime, *d_status, *d_ns, *d_samples = nullptr;
    if ((rc = host_buf(ctx, 0, (size_t)n * 6 * 8, &d_y0))) return rc;
    if ((rc = host_buf(ctx, 1, (size_t)n * 6 * 8 * 2, &d_state))) return rc;  // SoA state + AoS transpose
    if ((rc = host_buf(ctx, 2, (size_t)n * 3 * 8, &d_cam))) return rc;
    if ((rc = host_buf(ctx, 3, (size_t)n * 4, &d_istep))) return rc;
    if ((rc = host_buf(ctx, 4, (size_t)n * 8, &d_itime))) return rc;
    if ((rc = host_buf(ctx, 5, (size_t)n, &d_status))) return rc;
    if ((rc = host_buf(ctx, 6, (size_t)n * 4, &d_ns))) return rc;
    const size_t row = (size_t)RB_SAMPLE_FIELDS * n * sizeof(double);
    // Pinned (cudaHostAlloc / cudaHostRegister) sample buffers are written DIRECTLY by the step kernel over
    // PCIe (zero-copy streaming stores): the download is fused into the producing kernel, only valid samples
    // cross the link, and nothing is left to copy when the last launch ends.  Pageable buffers go through
    // a NaN-filled device buffer and chunked cudaMemcpyAsync.
    bool zero_copy = false;
    if (samples) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, samples) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer &&
            !(run->flags & RB_FLAG_STAGED_SAMPLES)) {
            zero_copy = true;
            d_samples = attr.devicePointer;
        } else {
            (void)cudaGetLastError();
            if ((rc = host_buf(ctx, 7, (size_t)n_out * row, &d_samples))) return rc;
            const int64_t cnt = n_out * RB_SAMPLE_FIELDS * n;
            k_fill_nan<<<(unsigned)((cnt + 255) / 256), 256, 0, cs>>>((double *)d_samples, cnt);
        }
    }
    CK(cudaMemcpyAsync(d_y0, y0, (size_t)n * 6 * 8, cudaMemcpyHostToDevice, cs));
    double *d_cd = nullptr, *d_area = nullptr, *d_mass = nullptr;
    if (cd) { d_cd = (double *)d_cam; CK(cudaMemcpyAsync(d_cd, cd, (size_t)n * 8, cudaMemcpyHostToDevice, cs)); }
    if (area) { d_area = (double *)d_cam + n; CK(cudaMemcpyAsync(d_area, area, (size_t)n * 8, cudaMemcpyHostToDevice, cs)); }
    if (mass) { d_mass = (double *)d_cam + 2 * n; CK(cudaMemcpyAsync(d_mass, mass, (size_t)n * 8, cudaMemcpyHostToDevice, cs)); }
    // the time table is built on the host while the initial states are on their way to the device
    const auto t_table = now();
    rc = rb_build_time_table(epoch_jd, gmst0, t0, dt, run->n_steps, ctx->h_table, 0);
    const double ms_table = ms_since(t_table);
    if (rc != RB_OK) return rc;
    rc = rb_ctx_set_time_table(ctx, ctx->h_table, run->n_steps, t0, dt, cs);
    if (rc != RB_OK) return rc;

    rb_in in{};
    in.n = n; in.y0 = (const double *)d_y0; in.y0_traj_stride = 6; in.y0_comp_stride = 1;
    in.cd = d_cd; in.area = d_area; in.mass = d_mass;
    in.cd_scalar = cd_scalar; in.area_scalar = area_scalar; in.mass_scalar = mass_scalar;
    rb_out out{};
    out.samples = (double *)d_samples; out.sample_stride = RB_SAMPLE_FIELDS * n; out.field_stride = n;
    out.n_out_capacity = n_out;
    out.final_state = (double *)d_state; out.impact_step = (int32_t *)d_istep; out.impact_time = (double *)d_itime;
    out.status = (uint8_t *)d_status; out.n_samples = (int32_t *)d_ns;

    SampleDownloader dl;
    dl.ctx = ctx; dl.d_samples = (const double *)d_samples; dl.h_samples = samples; dl.n = n;
    dl.out_stride = run->out_stride; dl.copy_stream = ctx->copy_stream;
    EventSet events;   // (destroyed on every exit path)
    cudaEvent_t ev[2] = {nullptr, nullptr};
    const bool staged = samples && !zero_copy;
    if (staged)
        for (int q = 0; q < 2; ++q) { CK(events.make(&ev[q])); dl.ev[q] = ev[q]; }
    rb_run r2 = *run;
    r2.flags |= RB_FLAG_EARLY_EXIT;
    int64_t steps_done = 0;
    SampleRoute route;
    if (zero_copy && !(run->flags & RB_FLAG_DIRECT_SAMPLES)) {
        void *d_staging = nullptr;
        if (host_buf(ctx, 7, (size_t)n_out * row, &d_staging) != RB_OK) { (void)cudaGetLastError(); d_staging = nullptr; }   // no room: direct stores only
      if (d_staging) {
        route.staging = (double *)d_staging; route.host = samples; route.copy_stream = ctx->copy_stream;
        for (int q = 0; q < 2; ++q) { CK(events.make(&ev[q])); route.ev[q] = ev[q]; }
        for (int q = 0; q < 2; ++q)
            for (int i = 0; i < 8; ++i) CK(events.make(&route.done[q][i]));
      }
    }
    const auto t_prop = now();
    rc = propagate_impl(ctx, &in, &r2, &out, cs, staged ? &dl : nullptr, &steps_done, route.staging ? &route : nullptr);
    if (trace && route.staging) fprintf(stderr, "[dma rows %lld, launches kept off the copy engine by back-pressure %lld] ", (long long)route.rows_via_dma, (long long)route.launches_backlogged);
    const double ms_enqueue = ms_since(t_prop);
    if (trace) { cudaStreamSynchronize(cs); fprintf(stderr, "rb_propagate_host: table %.2f ms, setup %.2f ms, propagate enqueue %.2f ms, propagate done %.2f ms", ms_table, std::chrono::duration<double, std::milli>(t_prop - t_begin).count() - ms_table, ms_enqueue, ms_since(t_prop)); }
    const auto t_dl = now();
    if (rc == RB_OK) {
        double *d_aos = (double *)d_state + 6 * n;
        k_soa_to_aos6<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>((const double *)d_state, d_aos, n);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(final_state, d_aos, (size_t)n * 6 * 8, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(impact_step, d_istep, (size_t)n * 4, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(impact_time, d_itime, (size_t)n * 8, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_status, (size_t)n, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && n_samples) e = cudaMemcpyAsync(n_samples, d_ns, (size_t)n * 4, cudaMemcpyDeviceToHost, cs);
        int32_t edge = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&edge, ctx->d_n_live + 8, sizeof(int32_t), cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && staged) {
            dl.after_launch(steps_done, &cs, 1);  // rows not yet sent (k_refine has joined the partitions)
            if (dl.err != cudaSuccess) e = dl.err;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
        if (e == cudaSuccess && staged && edge > 0) {
            // a grid sample coincided with an event time and was added by k_refine after its row had
            // been sent (measure-zero case): send the rows again
            e = cudaMemcpyAsync(samples, d_samples, (size_t)dl.rows_copied * row, cudaMemcpyDeviceToHost, cs);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->copy_stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "rb_propagate_host download");
    }
    if (rc != RB_OK) {   // drain: the copy engine may still be writing the caller's buffers
        cudaStreamSynchronize(cs);
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->aux_stream);
    }
    if (trace) fprintf(stderr, ", download %.2f ms, total %.2f ms\n", ms_since(t_dl), ms_since(t_begin));
    if (n_samples_used) *n_samples_used = !samples ? 0 : (staged ? dl.rows_copied : steps_done / run->out_stride + 1);
    return rc;
}

// ------------------------------------------------------------------ host-buffer path, packed sample slabs
namespace {

// The step kernel's table (96-byte entries) is built on the HOST while the GPU already runs, and only as far as
// launches are actually enqueued: a run "until impact" with a generous step cap pays for the steps it flies.  ONE
// builder thread per call (no OpenMP team: eight ranks of a box share its cores, and a team that has to be woken up
// or is oversubscribed showed 4-13 ms outliers on the critical path) runs ahead of the launch loop; the launch loop
// uploads what has been built before each launch that needs it.
struct TableFeeder : LaunchHook {
    rb_ctx *ctx = nullptr;
    double epoch_jd = 0, gmst0 = 0, t0 = 0, dt = 0;
    int64_t n_steps = 0;
    double *h_hot = nullptr;             // pinned staging, [entries][HOT_DOUBLES]
    cudaStream_t up_stream = nullptr;
    cudaEvent_t up_ev = nullptr;
    cudaError_t err = cudaSuccess;
    std::atomic<int64_t> built_steps{-1};   // entries 0 .. 5 built_steps are final in h_hot
    std::atomic<int64_t> wanted_steps{0};   // the builder stays at most a lookahead beyond this
    std::atomic<bool> stop{false};
    int64_t uploaded_steps = -1;
    std::thread builder;
    double ms_waited = 0.0;
    static constexpr int64_t LOOKAHEAD = 1024;

    void start() {
        builder = std::thread([this] {
            const double C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0};
            double tn = t0;                       // t_n exactly as scipy accumulates it (t_{n+1} = t_n + dt)
            double ent[ENTRY_DOUBLES];
            time_entry(tn, epoch_jd, gmst0, ent);
            hot_entry(ent, h_hot);
            built_steps.store(0, std::memory_order_release);
            for (int64_t n = 0; n < n_steps && !stop.load(std::memory_order_relaxed); ++n) {
                while (n >= wanted_steps.load(std::memory_order_relaxed) + LOOKAHEAD && !stop.load(std::memory_order_relaxed))
                    std::this_thread::yield();
                const double tn1 = tn + dt;
                for (int s = 1; s <= RB_STAGES_PER_STEP; ++s) {
                    const double te = (s == 5) ? tn1 : tn + C[s] * dt;
                    time_entry(te, epoch_jd, gmst0, ent);
                    hot_entry(ent, h_hot + (RB_STAGES_PER_STEP * n + s) * HOT_DOUBLES);
                }
                tn = tn1;
                if (((n + 1) & 15) == 0 || n + 1 == n_steps) built_steps.store(n + 1, std::memory_order_release);
            }
        });
    }
    void finish() {
        stop.store(true);
        if (builder.joinable()) builder.join();
    }
    ~TableFeeder() override { finish(); }

    int32_t before_launch(int64_t step_end, const cudaStream_t *streams, int n_streams) override {
        wanted_steps.store(step_end, std::memory_order_relaxed);
        if (step_end <= uploaded_steps) return RB_OK;
        const auto t_a = std::chrono::steady_clock::now();
        int64_t have;
        while ((have = built_steps.load(std::memory_order_acquire)) < step_end) std::this_thread::yield();
        ms_waited += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a).count();
        const int64_t e0 = uploaded_steps < 0 ? 0 : RB_STAGES_PER_STEP * uploaded_steps + 1, e1 = RB_STAGES_PER_STEP * have + 1;
        cudaError_t e = cudaMemcpyAsync(ctx->d_table + e0 * HOT_DOUBLES, h_hot + e0 * HOT_DOUBLES,
                                        (size_t)(e1 - e0) * HOT_DOUBLES * sizeof(double), cudaMemcpyHostToDevice, up_stream);
        if (e == cudaSuccess) e = cudaEventRecord(up_ev, up_stream);
        for (int q = 0; q < n_streams && e == cudaSuccess; ++q) e = cudaStreamWaitEvent(streams[q], up_ev, 0);
        if (e != cudaSuccess) { err = e; return RB_ERR_CUDA; }
        uploaded_steps = have;
        return RB_OK;
    }
};

// Every slab leaves for the host by copy engine as soon as its launch has finished AND the host knows how many of its
// columns are live: a slab is as wide as the host's (one launch old) view of the live count when its launch was
// enqueued; one loop iteration later the lagged poll tells the exact count at that launch's start, and the transfer
// (a 2-D copy: rows of `count` doubles at the slab's pitch) carries the live columns only.  Bytes on the link = valid
// samples + the columns of lanes that retire inside the launch itself.
struct SlabSender : TableFeeder {
    const double *d_arena = nullptr;
    double *h_arena = nullptr;
    cudaStream_t copy_stream = nullptr;
    static constexpr int RING = 8;
    struct Pending { rb_sample_slab *slab; int64_t launch; cudaEvent_t ev; };
    cudaEvent_t ring[2][RING] = {};
    std::vector<Pending> pending[2];
    int64_t n_queued[2] = {0, 0};
    int64_t bytes_sent = 0;
    void send(const Pending &pd, int64_t cols) {
        if (!h_arena || err != cudaSuccess) return;
        rb_sample_slab &sl = *pd.slab;
        cols = std::min<int64_t>(std::max<int64_t>(cols, 0), sl.width);
        cudaError_t r = cudaStreamWaitEvent(copy_stream, pd.ev, 0);
        if (r == cudaSuccess && cols > 0)
            r = cudaMemcpy2DAsync(h_arena + sl.offset, (size_t)sl.width * sizeof(double), d_arena + sl.offset, (size_t)sl.width * sizeof(double),
                                  (size_t)cols * sizeof(double), (size_t)(sl.n_rows * RB_SAMPLE_FIELDS), cudaMemcpyDeviceToHost, copy_stream);
        if (r != cudaSuccess) err = r;
        bytes_sent += cols * sl.n_rows * RB_SAMPLE_FIELDS * (int64_t)sizeof(double);
    }
    void after_slab(int64_t index, rb_sample_slab &sl, int q, int64_t launch, cudaStream_t stream) override {
        (void)index;
        if (!h_arena || err != cudaSuccess) return;
        cudaEvent_t e = ring[q][n_queued[q]++ % RING];   // (at most two slabs of a partition wait at any time)
        cudaError_t r = cudaEventRecord(e, stream);
        if (r != cudaSuccess) { err = r; return; }
        Pending pd{&sl, launch, e};
        if (launch < 1) send(pd, sl.n_live >= 0 ? sl.n_live : sl.width);   // initial states / first launch: everything is alive
        else pending[q].push_back(pd);
    }
    void live_count(int q, int64_t launch, int64_t count) override {
        auto &v = pending[q];
        size_t k = 0;
        for (; k < v.size() && v[k].launch <= launch; ++k) {
            if (v[k].launch == launch) v[k].slab->n_live = count;
            send(v[k], v[k].launch == launch ? count : v[k].slab->width);
        }
        v.erase(v.begin(), v.begin() + (long)k);
    }
    void finish_slabs() override {   // the last launches: their exact counts never reach the host in time
        for (int q = 0; q < 2; ++q) {
            for (auto &pd : pending[q]) send(pd, pd.slab->width);
            pending[q].clear();
        }
    }
};
}  // namespace

int32_t rb_propagate_host_packed(rb_ctx *ctx, int64_t n, const double *y0, const double *cd, const double *area,
                                 const double *mass, double cd_scalar, double area_scalar, double mass_scalar,
                                 double epoch_jd, double gmst0, double t0, double dt, const rb_run *run,
                                 double *packed, int64_t packed_capacity, rb_sample_slab *slabs, int64_t slab_capacity,
                                 int64_t *n_slabs, int64_t *packed_used, double *final_state, int32_t *impact_step,
                                 double *impact_time, uint8_t *status, int32_t *n_samples) {
    if (!ctx || !run || n < 0 || !y0 || !final_state || !impact_step || !impact_time || !status) return RB_ERR_INVALID_ARG;
    if (run->first_step != 0 || run->n_steps < 0 || run->out_stride < 1 || !(dt > 0)) return RB_ERR_INVALID_ARG;
    if (run->flags & RB_FLAG_RESUME) return RB_ERR_INVALID_ARG;
    if (packed && (!slabs || !n_slabs || slab_capacity < 1 || packed_capacity < 0)) return RB_ERR_INVALID_ARG;
    if (n_slabs) *n_slabs = 0;
    if (packed_used) *packed_used = 0;
    if (n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->compute_stream) CK(cudaStreamCreateWithFlags(&ctx->compute_stream, cudaStreamNonBlocking));
    if (!ctx->up_stream) CK(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->compute_stream;
    const bool trace = getenv("RB_HOST_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    // everything of the previous call on this context has drained (it synchronised), so buffers may be regrown
    int32_t rc = RB_OK;
    void *d_y0, *d_state, *d_cam = nullptr, *d_istep, *d_itime, *d_status, *d_ns, *d_arena = nullptr, *d_slab_live = nullptr;
    if ((rc = host_buf(ctx, 0, (size_t)n * 6 * 8, &d_y0))) return rc;
    if ((rc = host_buf(ctx, 1, (size_t)n * 6 * 8 * 2, &d_state))) return rc;  // SoA state + AoS transpose
    if ((rc = host_buf(ctx, 2, (size_t)n * 3 * 8, &d_cam))) return rc;
    if ((rc = host_buf(ctx, 3, (size_t)n * 4, &d_istep))) return rc;
    if ((rc = host_buf(ctx, 4, (size_t)n * 8, &d_itime))) return rc;
    if ((rc = host_buf(ctx, 5, (size_t)n, &d_status))) return rc;
    if ((rc = host_buf(ctx, 6, (size_t)n * 4, &d_ns))) return rc;
    if (packed) {
        if ((rc = host_buf(ctx, 7, (size_t)packed_capacity * 8, &d_arena))) return rc;
        if ((rc = host_buf(ctx, 8, (size_t)slab_capacity * 4, &d_slab_live))) return rc;
    }
    // time table: device capacity for the whole cap, contents fed chunk by chunk (TableFeeder)
    const int64_t n_entries = rb_time_table_entries(run->n_steps);
    if (n_entries > ctx->table_capacity) {
        cudaFree(ctx->d_table); cudaFree(ctx->d_table_abi);
        ctx->d_table = ctx->d_table_abi = nullptr; ctx->table_capacity = 0;
        CK(cudaMalloc(&ctx->d_table, (size_t)n_entries * HOT_DOUBLES * sizeof(double)));
        CK(cudaMalloc(&ctx->d_table_abi, (size_t)n_entries * ENTRY_DOUBLES * sizeof(double)));
        ctx->table_capacity = n_entries;
    }
    if (n_entries > ctx->h_table_entries) {
        cudaFreeHost(ctx->h_table);
        ctx->h_table = nullptr; ctx->h_table_entries = 0;
        CK(cudaHostAlloc(&ctx->h_table, (size_t)n_entries * ENTRY_DOUBLES * sizeof(double), cudaHostAllocDefault));
        ctx->h_table_entries = n_entries;
    }
    ctx->table_steps = run->n_steps; ctx->t0 = t0; ctx->dt = dt;
    ctx->table_generation++;   // (a caller's cached "my table is staged" no longer holds)

    EventSet events;
    SlabSender hook;
    hook.ctx = ctx; hook.epoch_jd = epoch_jd; hook.gmst0 = gmst0; hook.t0 = t0; hook.dt = dt; hook.n_steps = run->n_steps;
    hook.h_hot = ctx->h_table; hook.up_stream = ctx->up_stream;
    hook.start();   // the builder thread runs ahead from here on (the upload of y0 below overlaps its first entries)
    hook.d_arena = (const double *)d_arena; hook.h_arena = packed; hook.copy_stream = ctx->copy_stream;
    CK(events.make(&hook.up_ev));
    for (int q = 0; q < 2; ++q)
        for (int i = 0; i < SlabSender::RING; ++i) CK(events.make(&hook.ring[q][i]));

    CK(cudaMemcpyAsync(d_y0, y0, (size_t)n * 6 * 8, cudaMemcpyHostToDevice, cs));
    double *d_cd = nullptr, *d_area = nullptr, *d_mass = nullptr;
    if (cd) { d_cd = (double *)d_cam; CK(cudaMemcpyAsync(d_cd, cd, (size_t)n * 8, cudaMemcpyHostToDevice, cs)); }
    if (area) { d_area = (double *)d_cam + n; CK(cudaMemcpyAsync(d_area, area, (size_t)n * 8, cudaMemcpyHostToDevice, cs)); }
    if (mass) { d_mass = (double *)d_cam + 2 * n; CK(cudaMemcpyAsync(d_mass, mass, (size_t)n * 8, cudaMemcpyHostToDevice, cs)); }

    rb_in in{};
    in.n = n; in.y0 = (const double *)d_y0; in.y0_traj_stride = 6; in.y0_comp_stride = 1;
    in.cd = d_cd; in.area = d_area; in.mass = d_mass;
    in.cd_scalar = cd_scalar; in.area_scalar = area_scalar; in.mass_scalar = mass_scalar;
    rb_out out{};
    out.samples = (double *)d_arena; out.packed_capacity = packed_capacity;
    out.slabs = slabs; out.slab_capacity = slab_capacity; out.slab_live = (int32_t *)d_slab_live;
    int64_t ns_local = 0;
    out.n_slabs = &ns_local;
    out.final_state = (double *)d_state; out.impact_step = (int32_t *)d_istep; out.impact_time = (double *)d_itime;
    out.status = (uint8_t *)d_status; out.n_samples = (int32_t *)d_ns;
    rb_run r2 = *run;
    r2.flags |= RB_FLAG_EARLY_EXIT | RB_FLAG_COMPACT | (packed ? RB_FLAG_PACKED_SAMPLES : 0u);
    r2.flags &= ~(uint32_t)RB_FLAG_PROFILE;
    int64_t steps_done = 0;
    const auto t_prop = std::chrono::steady_clock::now();
    rc = propagate_impl(ctx, &in, &r2, &out, cs, &hook, &steps_done);
    const double ms_enqueue = ms_since(t_prop);
    hook.finish();
    cudaError_t e = hook.err;
    if (rc == RB_OK && e == cudaSuccess) {
        double *d_aos = (double *)d_state + 6 * n;
        k_soa_to_aos6<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>((const double *)d_state, d_aos, n);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(final_state, d_aos, (size_t)n * 6 * 8, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(impact_step, d_istep, (size_t)n * 4, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(impact_time, d_itime, (size_t)n * 8, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_status, (size_t)n, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && n_samples) e = cudaMemcpyAsync(n_samples, d_ns, (size_t)n * 4, cudaMemcpyDeviceToHost, cs);
    }
    // drain: the call returns with every buffer final (also on errors: the copy engine may still be writing `packed`)
    int32_t edge = 0;
    std::vector<int32_t> live((size_t)std::max<int64_t>(ns_local, 1));
    if (rc == RB_OK && e == cudaSuccess) {
        e = cudaMemcpyAsync(&edge, ctx->d_n_live + 8, sizeof(int32_t), cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && packed && ns_local > 0)
            e = cudaMemcpyAsync(live.data(), d_slab_live, (size_t)ns_local * 4, cudaMemcpyDeviceToHost, cs);
    }
    cudaError_t e2 = cudaStreamSynchronize(cs);
    if (e == cudaSuccess) e = e2;
    if (rc == RB_OK && e == cudaSuccess && packed && edge > 0) {
        // a grid sample coincided with an event time and was added by k_refine after its slab had been sent
        // (measure-zero case): send the arena again
        const int64_t used = ns_local ? slabs[ns_local - 1].offset + slabs[ns_local - 1].n_rows * RB_SAMPLE_FIELDS * slabs[ns_local - 1].width : 0;
        e = cudaMemcpyAsync(packed, d_arena, (size_t)used * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream);
    }
    e2 = cudaStreamSynchronize(ctx->copy_stream);
    if (e == cudaSuccess) e = e2;
    e2 = cudaStreamSynchronize(ctx->up_stream);
    if (e == cudaSuccess) e = e2;
    e2 = cudaStreamSynchronize(ctx->aux_stream);
    if (e == cudaSuccess) e = e2;
    if (rc == RB_OK && e != cudaSuccess) rc = cuda_fail(ctx, e, "rb_propagate_host_packed");
    ctx->stats.sample_bytes_sent = hook.bytes_sent;
    if (rc == RB_OK && packed) {
        for (int64_t i = 1; i < ns_local; ++i) slabs[i].n_live = live[(size_t)i];   // (exact, from the device; the lagged polls agree)
        *n_slabs = ns_local;
        if (packed_used) *packed_used = ns_local ? slabs[ns_local - 1].offset + slabs[ns_local - 1].n_rows * RB_SAMPLE_FIELDS * slabs[ns_local - 1].width : 0;
    }
    if (trace)
        fprintf(stderr, "rb_propagate_host_packed: waited %.2f ms for the table builder (%lld of %lld steps built), enqueue %.2f ms, total %.2f ms, %lld slabs, %.3f GB sent\n",
                hook.ms_waited, (long long)hook.built_steps.load(), (long long)run->n_steps, ms_enqueue, ms_since(t_begin), (long long)ns_local, hook.bytes_sent * 1e-9);
    return rc;
}

int32_t rb_unpack_samples(const double *packed, const rb_sample_slab *slabs, int64_t n_slabs, int64_t n,
                          const int32_t *impact_step, const uint8_t *status, int64_t n_out, double *dense, int32_t n_threads) {
    if (!packed || !slabs || n_slabs < 0 || n < 0 || !impact_step || !status || n_out < 0 || !dense) return RB_ERR_INVALID_ARG;
    const double qnan = nan("");
    const int64_t total = n_out * RB_SAMPLE_FIELDS * n;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = std::max(omp_get_max_threads(), std::min(omp_get_num_procs(), 16));
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (int64_t i = 0; i < total; ++i) dense[i] = qnan;
    std::vector<int32_t> cols;
    for (int64_t s = 0; s < n_slabs; ++s) {
        const rb_sample_slab &sl = slabs[s];
        if (sl.first_row + sl.n_rows > n_out || sl.first_traj < 0 || sl.first_traj + sl.n_traj > n) return RB_ERR_INVALID_ARG;
        cols.clear();   // column map: trajectories of the partition still alive when the launch started, ascending
        const bool identity = sl.n_live == sl.n_traj;   // nothing retired yet, or a run without compaction (the list never shrinks)
        for (int64_t i = sl.first_traj; i < sl.first_traj + sl.n_traj; ++i)
            if (identity || status[i] == RB_TRAJ_ALIVE || impact_step[i] > sl.first_step) cols.push_back((int32_t)i);
        if ((int64_t)cols.size() > sl.width || (sl.n_live >= 0 && (int64_t)cols.size() != sl.n_live)) return RB_ERR_INVALID_ARG;
        const int64_t nc = (int64_t)cols.size();
        const int32_t *cp = cols.data();
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads) collapse(2)
#endif
        for (int64_t r = 0; r < sl.n_rows; ++r)
            for (int f = 0; f < RB_SAMPLE_FIELDS; ++f) {
                const double *src = packed + sl.offset + (r * RB_SAMPLE_FIELDS + f) * sl.width;
                double *dst = dense + ((sl.first_row + r) * RB_SAMPLE_FIELDS + f) * n;
                for (int64_t j = 0; j < nc; ++j) dst[cp[j]] = src[j];
            }
    }
    (void)n_threads;
    return RB_OK;
}

void *rb_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void rb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ N4: adaptive mode
int32_t rb_propagate_adaptive(rb_ctx *ctx, const rb_in *in, const rb_adaptive_run *run, const rb_adaptive_out *out,
                              void *stream) {
    if (!ctx || !in || !run || !out) return RB_ERR_INVALID_ARG;
    if (in->n < 0 || !in->y0 || !out->final_state || !out->status) return RB_ERR_INVALID_ARG;
    if (!(run->rtol > 0) || !(run->atol >= 0) || !(run->t_bound >= run->t0) || run->n_eval < 0 || run->max_steps < 1)
        return RB_ERR_INVALID_ARG;
    if (run->method < RB_METHOD_RK45 || run->method > RB_METHOD_DOP853) return RB_ERR_INVALID_ARG;
    if (in->n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    AdaptiveParams p{};
    p.n = in->n; p.y0 = in->y0; p.y0_traj_stride = in->y0_traj_stride; p.y0_comp_stride = in->y0_comp_stride;
    p.cd = in->cd; p.area = in->area; p.mass = in->mass;
    p.cd_s = in->cd_scalar; p.area_s = in->area_scalar; p.mass_s = in->mass_scalar;
    p.epoch_jd = run->epoch_jd; p.gmst0 = run->gmst0; p.t0 = run->t0; p.t_bound = run->t_bound;
    p.rtol = std::max(run->rtol, 100 * 2.220446049250313e-16);   // validate_tol, common.py:44-51
    p.atol = run->atol; p.t_eval0 = run->t_eval0; p.t_eval_dt = run->t_eval_dt; p.event_altitude = run->event_altitude;
    p.n_eval = run->n_eval; p.max_steps = run->max_steps;
    p.samples = out->samples; p.sample_stride = out->sample_stride; p.field_stride = out->field_stride;
    p.final_state = out->final_state; p.impact_time = out->impact_time; p.status = out->status;
    p.n_samples = out->n_samples; p.n_accepted = out->n_accepted; p.n_rejected = out->n_rejected;
    if (out->step_log && out->step_capacity < 0) return RB_ERR_INVALID_ARG;
    p.step_log = out->step_log; p.step_capacity = out->step_log ? out->step_capacity : 0;
    if (run->ephemeris_dt > 0.0) {   // time-only inputs from a grid (built on the device) instead of at every stage time
        const double span = run->t_bound - run->t0;
        if (!(span / run->ephemeris_dt < 2.0e8)) return RB_ERR_INVALID_ARG;
        const int64_t n_nodes = (int64_t)ceil(span / run->ephemeris_dt) + 4;   // one node before t0, two after t_bound
        if (n_nodes > ctx->eph_capacity) {
            CK(cudaStreamSynchronize((cudaStream_t)stream));
            cudaFree(ctx->d_eph);
            ctx->d_eph = nullptr; ctx->eph_capacity = 0;
            CK(cudaMalloc(&ctx->d_eph, (size_t)n_nodes * HOT_DOUBLES * sizeof(double)));
            ctx->eph_capacity = n_nodes;
        }
        p.eph = ctx->d_eph; p.eph_t0 = run->t0 - run->ephemeris_dt; p.eph_inv_dt = 1.0 / run->ephemeris_dt; p.eph_n = n_nodes;
        k_eph_nodes<<<(unsigned)((n_nodes + 127) / 128), 128, 0, (cudaStream_t)stream>>>(ctx->d_eph, n_nodes, p.eph_t0, run->ephemeris_dt,
                                                                                         run->epoch_jd, run->gmst0);
    }
    const unsigned agrid = (unsigned)((in->n + ADAPTIVE_BLOCK - 1) / ADAPTIVE_BLOCK);
    if (run->method == RB_METHOD_DOP853) k_adaptive<2><<<agrid, ADAPTIVE_BLOCK, (RB_ADAPTIVE_K_SHARED ? adaptive_smem_bytes<2>() : 0), (cudaStream_t)stream>>>(p);
    else if (run->method == RB_METHOD_RK23) k_adaptive<1><<<agrid, ADAPTIVE_BLOCK, (RB_ADAPTIVE_K_SHARED ? adaptive_smem_bytes<1>() : 0), (cudaStream_t)stream>>>(p);
    else k_adaptive<0><<<agrid, ADAPTIVE_BLOCK, (RB_ADAPTIVE_K_SHARED ? adaptive_smem_bytes<0>() : 0), (cudaStream_t)stream>>>(p);
    CK(cudaGetLastError());
    return RB_OK;
}

// ------------------------------------------------------------------ N2: ground track
int32_t rb_ground_track(rb_ctx *ctx, int64_t n, int64_t n_out, const double *samples, int64_t sample_stride,
                        int64_t field_stride, const int32_t *n_samples, double t0, double sample_dt, double gmst0,
                        double threshold, double *geo, double *downrange, double *cross_time, double *cross_downrange,
                        int32_t *n_cross, void *stream) {
    if (!ctx || n < 0 || n_out < 0 || !samples) return RB_ERR_INVALID_ARG;
    if (n == 0 || n_out == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    GroundTrackParams p{n, n_out, samples, sample_stride, field_stride, n_samples, t0, sample_dt, gmst0, threshold,
                        geo, downrange, cross_time, cross_downrange, n_cross};
    k_ground_track<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
    CK(cudaGetLastError());
    return RB_OK;
}

int32_t rb_impact_points(rb_ctx *ctx, int64_t n, const double *final_state, const double *impact_time,
                         const uint8_t *status, double t_end, double gmst0, double *out, void *stream) {
    if (!ctx || n < 0 || !final_state || !status || !out) return RB_ERR_INVALID_ARG;
    if (n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    k_impact_points<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, final_state, impact_time, status, t_end,
                                                                                  gmst0, out);
    CK(cudaGetLastError());
    return RB_OK;
}

// ------------------------------------------------------------------ math self-test
int32_t rb_math_probe(rb_ctx *ctx, int32_t op, int64_t n, const double *x, const double *y, double *out, void *stream) {
    if (!ctx || n < 0 || !x || !out || op < 0 || op > 15) return RB_ERR_INVALID_ARG;
    if (n == 0) return RB_OK;
    CK(cudaSetDevice(ctx->device));
    k_math_probe<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(op, n, x, y, out);
    CK(cudaGetLastError());
    return RB_OK;
}

// ------------------------------------------------------------------ FP64 peak probe
int32_t rb_fp64_peak_probe(rb_ctx *ctx, double ms, double *tflops_out, void *stream) {
    if (!ctx || !tflops_out) return RB_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    double *sink = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = prop.multiProcessorCount * 8, block = 256;
    double best = 0.0;
    auto work = [&]() -> cudaError_t {   // (one exit below frees what was allocated, whatever fails)
        cudaError_t e = cudaMalloc(&sink, 8);
        if (e == cudaSuccess) e = cudaEventCreate(&a);
        if (e == cudaSuccess) e = cudaEventCreate(&b);
        if (e != cudaSuccess) return e;
        int iters = 2000;
        double elapsed_total = 0.0;
        k_dfma_peak<<<grid, block, 0, s>>>(sink, 200, 1.0);  // warm-up
        for (int rep = 0; rep < 50 && elapsed_total < ms; ++rep) {
            if ((e = cudaEventRecord(a, s)) != cudaSuccess) return e;
            k_dfma_peak<<<grid, block, 0, s>>>(sink, iters, 1.0 + rep);
            if ((e = cudaEventRecord(b, s)) != cudaSuccess) return e;
            if ((e = cudaEventSynchronize(b)) != cudaSuccess) return e;
            float t = 0;
            if ((e = cudaEventElapsedTime(&t, a, b)) != cudaSuccess) return e;
            const double flop = 2.0 * DFMA_PER_THREAD_ITER * (double)iters * grid * block;
            best = std::max(best, flop / (t * 1e-3) / 1e12);
            elapsed_total += t;
            if (t < 5.0f) iters *= 2;
        }
        return cudaGetLastError();
    };
    const cudaError_t e = work();
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    cudaFree(sink);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "rb_fp64_peak_probe");
    *tflops_out = best;
    return RB_OK;
}

}  // extern "C"
