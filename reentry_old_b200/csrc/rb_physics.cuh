// rb_physics.cuh -- per-trajectory FP64 math of the hot path (device inline functions).
//
// What is computed follows the reference line by line (citations below); how it is computed
// is B200-first: everything that depends on time only (Sun/Moon ephemeris, the indirect
// third-body terms, cos/sin of GMST, the time part of the solar-activity factor) comes from
// a 128-byte time-table entry staged in shared memory, shared denominators are inverted
// once, and the whole evaluation lives in registers.
//
// The functions are __host__ __device__ so that tests can run the SAME source on the CPU
// (tests/host_harness.cu) before any GPU time is spent; the product only ever calls them
// from kernels.
#pragma once

#include <math.h>
#include <stdint.h>

#include "reentry_constants.h"
#include "rb_math.cuh"

namespace rb {

constexpr double RB_NEGLIGIBLE_DRAG_ALTITUDE_ = 400000.0;  // == RB_NEGLIGIBLE_DRAG_ALTITUDE of the C ABI

// time-table entry layout of the C ABI (include/reentry_b200.h): what rb_build_time_table writes
enum : int {
    E_SUN = 0, E_MOON = 3, E_SUN_IND = 6, E_MOON_IND = 9, E_CG = 12, E_SG = 13, E_SOLAR = 14, E_T = 15,
    ENTRY_DOUBLES = 16
};
// ... and the 96-byte form the step kernel streams through shared memory (rb_ctx_set_time_table converts on the
// device): only what the hot RHS reads, with the two time-only indirect third-body terms pre-summed and negated
// so that they ride on the gravity multiply (a = k r - ind, one DFMA per component).
enum : int { H_SUN = 0, H_MOON = 3, H_NIND = 6, H_SOLAR = 9, H_T = 10, H_SOLAR_M1 = 11, HOT_DOUBLES = 12 };
RB_HD void hot_entry(const double *__restrict__ e, double *__restrict__ h) {
    for (int c = 0; c < 3; ++c) {
        h[H_SUN + c] = e[E_SUN + c];
        h[H_MOON + c] = e[E_MOON + c];
        h[H_NIND + c] = -(e[E_MOON_IND + c] + e[E_SUN_IND + c]);
    }
    h[H_SOLAR] = e[E_SOLAR];
    h[H_T] = e[E_T];
    h[H_SOLAR_M1] = e[E_SOLAR] - 1.0;   // (factor - 1) of the sigmoid blend, :166
}

// ---- constants of the hot RHS, kept in the constant bank ---------------------------------
// A 64-bit literal that is not a 32-bit-encodable immediate costs two UMOVs per use on sm_100
// (ptxas rematerialises them); constant-bank operands (c[3][off]) ride along with DFMA/DMUL
// for free.  The host copy serves the CPU build of this header (tests/host_harness.cu).
struct Consts {
    double omega, neg_mu, j2c, j2c2, moon_k, sun_k, earth_r, one_m_e2, e2r, one_m_e2r_over_r, tan_tol, tan_tol_sure, rad2deg, latc, latc_rad;
    double sig_ks, sig_x0, blend_alt, top_rho0, rgas, top_h, top_T, top_P, top_k;
    double sig_c, top_c;         // sig_ks sig_x0 ; -top_k top_h  (the exponents as one fma of the altitude)
    double sea_rho, j2f;         // (j2f: diagnostics only -- kept off the 16-byte pairs of the hot constants above)
    double j2c5n;                // -5 j2c
    double bp[RB_N_LAYERS], gr[RB_N_LAYERS], bt[RB_N_LAYERS], bpz[RB_N_LAYERS];
    double iso_k[RB_N_LAYERS];   // gM / (R* T_i)        (isothermal layers, spacecraft_model.py:84)
    double pow_k[RB_N_LAYERS];   // gM / (R* L_i)        (gradient layers, :86); 0 where L_i == 0
    double inv_bt[RB_N_LAYERS];  // 1 / T_i
    // the same per-layer constants packed for the hot path: one 64-byte record per layer (three LDC.128 instead of
    // nine LDC.64 with a lane-dependent index): {h_i, L_i}, {T_i, L_i / T_i}, {k_i, P_i / R_GAS} with k_i = gM/(R* T_i) on
    // isothermal layers and gM/(R* L_i) on gradient layers
    double layer[RB_N_LAYERS][8];
    int32_t bp_hi[RB_N_LAYERS];  // high 32 bits of the altitude breakpoints: `altitude >= h_i` as one integer compare
};

#define RB_GM_ (-RB_EARTH_GRAVITY * RB_AIR_MOLAR_MASS)
#define RB_T_(i) (rb_bt_(i))
constexpr double rb_bp_(int i) { constexpr double a[RB_N_LAYERS] = RB_ALTITUDE_BREAKPOINTS_INIT; return a[i]; }
constexpr double rb_gr_(int i) { constexpr double a[RB_N_LAYERS] = RB_TEMPERATURE_GRADIENTS_INIT; return a[i]; }
constexpr double rb_bt_(int i) { constexpr double a[RB_N_LAYERS] = RB_BASE_TEMPERATURES_INIT; return a[i]; }
constexpr double rb_bpz_(int i) { constexpr double a[RB_N_LAYERS] = RB_BASE_PRESSURES_INIT; return a[i]; }
constexpr double rb_isok_(int i) { return RB_GM_ / (RB_GAS_CONSTANT * rb_bt_(i)); }
constexpr double rb_powk_(int i) { return rb_gr_(i) == 0.0 ? 0.0 : RB_GM_ / (RB_GAS_CONSTANT * rb_gr_(i)); }
#define RB_L8_(f) {f(0), f(1), f(2), f(3), f(4), f(5), f(6), f(7)}
constexpr double rb_invbt_(int i) { return 1.0 / rb_bt_(i); }
constexpr int32_t rb_hi_word_(double x) {   // high 32 bits of a positive normal double (compile time)
    int e = 0;
    double m = x;
    while (m >= 2.0) { m *= 0.5; ++e; }
    while (m < 1.0) { m *= 2.0; --e; }
    const uint64_t frac = (uint64_t)((m - 1.0) * 4503599627370496.0);
    return (int32_t)((((uint64_t)(1023 + e) << 52) | frac) >> 32);
}
constexpr int32_t rb_bphi_(int i) { return i == 0 ? 0 : rb_hi_word_(rb_bp_(i)); }
static_assert(rb_hi_word_(90000.0) == 0x40F5F900 && rb_hi_word_(1.0) == 0x3FF00000, "rb_hi_word_");
#define RB_LAYER_(i) {rb_bp_(i), rb_gr_(i), rb_bt_(i), rb_gr_(i) / rb_bt_(i),                         \
                      rb_gr_(i) == 0.0 ? rb_isok_(i) : rb_powk_(i), rb_bpz_(i) / RB_R_GAS, 0.0, 0.0}

#define RB_CONSTS_INIT                                                                              \
    {RB_EARTH_OMEGA, -RB_EARTH_MU,                                                                   \
     1.5 * RB_EARTH_J2 * (RB_EARTH_R * RB_EARTH_R), 3.0 * RB_EARTH_J2 * (RB_EARTH_R * RB_EARTH_R),    \
     RB_MOON_K, RB_SUN_K, RB_EARTH_R, 1.0 - RB_EARTH_E2, RB_EARTH_E2 * RB_EARTH_R,                         \
     (1.0 - RB_EARTH_E2 * RB_EARTH_R) / RB_EARTH_R, 1.0000000000003333e-06 /* tan(1e-6) */, 100.0 * 1.0000000000003333e-06, 180.0 / RB_PI,           \
     RB_LAT_FACTOR_COEF / 90.0, (RB_LAT_FACTOR_COEF / 90.0) * (180.0 / RB_PI), RB_SIGMOID_K * RB_SIGMOID_SMOOTHNESS, RB_SIGMOID_X0,                 \
     RB_SOLAR_BLEND_ALT, rb_bpz_(7) / (RB_R_GAS * rb_bt_(7)), RB_R_GAS, rb_bp_(7), rb_bt_(7), rb_bpz_(7),                            \
     rb_isok_(7) - 1.0 / RB_SCALE_HEIGHT,                                                            \
     (RB_SIGMOID_K * RB_SIGMOID_SMOOTHNESS) * RB_SIGMOID_X0, -(rb_isok_(7) - 1.0 / RB_SCALE_HEIGHT) * rb_bp_(7),   \
     RB_SEA_LEVEL_RHO, 1.5 * RB_EARTH_MU * RB_EARTH_J2 * (RB_EARTH_R * RB_EARTH_R),                   \
     -7.5 * RB_EARTH_J2 * (RB_EARTH_R * RB_EARTH_R),                                                 \
     RB_L8_(rb_bp_), RB_L8_(rb_gr_), RB_L8_(rb_bt_), RB_L8_(rb_bpz_), RB_L8_(rb_isok_),              \
     RB_L8_(rb_powk_), RB_L8_(rb_invbt_), RB_L8_(RB_LAYER_), RB_L8_(rb_bphi_)}

#if defined(__CUDACC__)
__constant__ Consts c_k = RB_CONSTS_INIT;
// per-layer records read with a LANE-dependent index: through L1 (LDG.128) or the block's shared-memory table image
// (table policy of rb_math.cuh) instead of the constant bank, whose indexed loads replay once per distinct address
__device__ const double __align__(64) g_layer[RB_N_LAYERS][8] = RB_L8_(RB_LAYER_);
__device__ __forceinline__ void tab_layer(const TabGlobal &, int i, double lay[6]) {
    const double2 l01 = __ldg(reinterpret_cast<const double2 *>(&g_layer[i][0]));
    const double2 l23 = __ldg(reinterpret_cast<const double2 *>(&g_layer[i][2]));
    const double2 l45 = __ldg(reinterpret_cast<const double2 *>(&g_layer[i][4]));
    lay[0] = l01.x; lay[1] = l01.y; lay[2] = l23.x; lay[3] = l23.y; lay[4] = l45.x; lay[5] = l45.y;
}
__device__ __forceinline__ void tab_layer(const TabShared &tb, int i, double lay[6]) { tb.layer(i, lay); }
static_assert(RB_N_LAYERS == 8, "table image layout (rb_math.cuh TAB_LAYER) assumes 8 layer records");
#endif
static const Consts h_k = RB_CONSTS_INIT;
#if defined(__CUDA_ARCH__)
#define KC(f) (rb::c_k.f)
#else
#define KC(f) (rb::h_k.f)
#endif

struct Body {  // per-trajectory constant bc = 0.5 Cd A / m (SpacecraftModel.Cd / .A / .m, spacecraft_model.py:394-397, :204-206)
    double bc;
};
struct RhsOut;
// FAST: the budgeted arithmetic of the fixed-step hot path (second-order sqrt / rcp / rsqrt, truncated series tails; every
// variant's bound is stated at its definition and pinned by rb_math_probe): default for the non-diagnostic evaluation.  The
// adaptive kernel evaluates the same table-entry form with FAST = false (its error estimator sees the RHS at full precision).
template <bool DIAG, bool SKIP = false, class TB = TabGlobal, bool FAST = !DIAG>
RB_HD void acceleration(const double *__restrict__ e, const double r[3], const double v[3], const double *bcp,
                        double a[3], RhsOut *diag, double *altitude_out = nullptr, const TB &tb = TB());
template <bool DIAG, bool SKIP = false, bool FAST = !DIAG>
RB_HD void acceleration(const double *__restrict__ e, const double r[3], const double v[3], const Body &b,
                        double a[3], RhsOut *diag, double *altitude_out = nullptr) {
    acceleration<DIAG, SKIP, TabGlobal, FAST>(e, r, v, &b.bc, a, diag, altitude_out, TabGlobal());
}
RB_HD Body make_body(double cd, double area, double mass) { return Body{0.5 * cd * area / mass}; }

struct RhsOut {  // optional diagnostics (dict of spacecraft_model.py:472-480)
    double a_grav[3], a_j2[3], a_moon[3], a_sun[3], a_drag[3];
    double altitude, latitude_deg, rho, T, v_rel_norm;
};

// N1: heat_transfer / spacecraft_temperature (spacecraft_model.py:223-255), stateless per sample.
// out = (Qc, Qr, Q_net, Q, T_s, dT) in the reference's RETURN order.
struct Thermal {
    double capsule_length, dt, thermal_conductivity, specific_heat_capacity, emissivity, ablation_efficiency, iter_fact;
};
RB_HD void spacecraft_temperature(double v_norm, double atmo_T, double a_drag_norm, double mass, const Thermal &th,
                                  double out[6]) {
    double T_s = atmo_T;
    const int idt = (int)(th.dt / th.iter_fact);
    const double dti = (double)idt;
    for (int k = 0; k < 6; ++k) out[k] = nan("");
    const double Q = th.ablation_efficiency * ((mass * a_drag_norm) * v_norm);
    const double Ta2 = atmo_T * atmo_T, Ta4 = Ta2 * Ta2;
    for (int it = 0; it < idt; ++it) {
        const double Qc = th.thermal_conductivity * (T_s - atmo_T) / th.capsule_length;
        const double Ts2 = T_s * T_s;
        const double Qr = th.emissivity * 5.67e-8 * (Ts2 * Ts2 - Ta4);
        const double Q_net = Q - Qc - Qr;
        const double dT = Q_net / (mass * th.specific_heat_capacity) * dti;
        out[0] = Qc; out[1] = Qr; out[2] = Q_net; out[3] = Q; out[5] = dT;
        T_s += dT;
    }
    out[4] = T_s;
}

// |geodetic latitude| in DEGREES of an ECEF point: coordinate_converter.py:81-96 with its quirks
// (theta built with (1 - e2*R); the loop tests |lat - lat_new| < 1e-6 BEFORE lat = lat_new, so
// the PREVIOUS iterate is returned; at most 100 iterations).
//
// B200 form: every angle of the reference's iteration is carried as an (A, B) pair with
// tan(angle) = A / B, B > 0 -- atan2(a, b) followed by sin/cos is a normalisation that cancels:
//   N cos(lat) = R B / sqrt(B^2 + (1 - e2) A^2)                       (N = R / sqrt(1 - e2 sin^2 lat))
//   lat_new    = atan2(z/p, 1 - e2 N / (N + h)),  N + h = p / cos(lat) (coordinate_converter.py:89-90)
//              = pair (z/p, 1 - (e2 R / p) B / sqrt(B^2 + (1 - e2) A^2))
//   |lat - lat_new| < tol  <=>  |A B' - A' B| < tan(tol) (B B' + A A')
// i.e. the same iterates and the same break decision as the reference with ONE rsqrt per
// iteration and ONE inverse-trig call (the returned angle) instead of 2 + n atan2 and 1 + n
// sincos.  Only |lat| is needed by the density (spacecraft_model.py:76).
//   p2 = x^2 + y^2, inv_p = 1/sqrt(p2) are passed in (shared with the caller).
// FAST (the hot RHS, which needs |lat| to 1e-11 rad at most): second-order rsqrt (rb_math.cuh); the same iterates to
// 1e-12 and the same break decisions except within 1e-12 of the tolerance.
template <bool FAST = false>
RB_HD void geodetic_pair(double p2, double inv_p, double z, double &A_out, double &B_out) {
    const double p = p2 * inv_p;
    // theta = atan2(z R, p (1 - e2 R)) -> (sin, cos) by normalisation
    // (scale invariant: (z R, p (1 - e2 R)) is normalised as (z, p (1 - e2 R) / R) -- one multiply and one constant less)
    const double tb = p * KC(one_m_e2r_over_r);
    const double th = FAST ? rb_rsqrt_q(fma(z, z, tb * tb)) : rb_rsqrt(fma(z, z, tb * tb));
    const double st = z * th, ct = tb * th;
    // lat_0 = atan2(z + e2R sin^3 theta, p - e2R cos^3 theta)
    double A = fma(KC(e2r), st * (st * st), z), B = fma(-KC(e2r), ct * (ct * ct), p);
    const double zp = z * inv_p;          // A of every later iterate
    const double ke = KC(e2r) * inv_p;    // e2 R / p
    {
        const double qa = fma(B, B, KC(one_m_e2) * (A * A));
        const double q = FAST ? rb_rsqrt_q(qa) : rb_rsqrt(qa);
        const double Bn = fma(-ke * B, q, 1.0);
        const bool conv = fabs(fma(A, Bn, -(zp * B))) < KC(tan_tol) * fma(B, Bn, A * zp);
        if (!conv) {
            A = zp; B = Bn;
            const double zp2 = zp * zp, g = KC(one_m_e2) * zp2, azp = fabs(zp);
            for (int it = 1; it < RB_GEODETIC_MAX_ITER; ++it) {
                const double q2 = FAST ? rb_rsqrt_q(fma(B, B, g)) : rb_rsqrt(fma(B, B, g));
                const double Bn2 = fma(-ke * B, q2, 1.0);
                const double d = azp * fabs(Bn2 - B), S = fma(B, Bn2, zp2);
                if (d < KC(tan_tol) * S) break;
                B = Bn2;
                // The iteration map B -> 1 - ke B / sqrt(B^2 + g) contracts by |f'| = ke g / (B^2 + g)^1.5
                // <= 1.03 e2 R/|r| < 0.008 (|r| > 0.9 R, 0.99 < B < 1), and the right-hand side of the test
                // shrinks by at most 0.98: d < 122 tol S (100 is used) here PROVES the next pass would break and return
                // this Bn2 -- the pass that only confirms convergence is not computed (same returned iterate).
                // (Peeling the first pass out of the loop was tried: the copy of the loop it leaves behind in each of
                // the step kernel's seven inlined evaluations costs more instruction cache than the peel saves issue slots,
                // and an out-of-line copy makes the callers spill.)
                if (d < 0x1.a36e3p-14 * S) break;   // 1.0e-4 = 100 tan(1e-6), written with a zero low word (immediate)
            }
        }
    }
    A_out = A; B_out = B;
}
template <bool FAST = false, class TB = TabGlobal>
RB_HD double pair_abs_latitude_rad(double A, double B, const TB &tb = TB()) {
    const double nr = rb_rsqrt_seed(fma(A, A, B * B));   // approximate length: atan2 is scale free
    return rb_atan2_unit<FAST>(fabs(A) * nr, B * nr, tb);
}
RB_HD double pair_abs_latitude_deg(double A, double B) { return pair_abs_latitude_rad<false>(A, B) * KC(rad2deg); }
RB_HD double geodetic_abs_latitude_deg(double p2, double inv_p, double z) {
    double A, B;
    geodetic_pair<false>(p2, inv_p, z, A, B);
    return pair_abs_latitude_deg(A, B);
}

// Full ecef_to_geodetic (coordinate_converter.py:81-96): signed latitude, longitude (degrees) and
// the height h = p / cos(lat) - N of the RETURNED iterate, for the ground-track post-pass (N2).
RB_HD void geodetic_full(double x, double y, double z, double &lat_deg, double &lon_deg, double &h) {
    const double p2 = fma(x, x, y * y);
    const double ip = rb_rsqrt(p2);
    double A, B;
    geodetic_pair<false>(p2, ip, z, A, B);
    const double n2 = fma(A, A, B * B);
    const double nr = rb_rsqrt(n2);
    const double a = rb_atan2_unit(fabs(A) * nr, B * nr) * KC(rad2deg);
    lat_deg = (A < 0.0) ? -a : a;
    lon_deg = atan2(y, x) * KC(rad2deg);
    // N = R / sqrt(1 - e2 sin^2) = R sqrt(n2) / sqrt(B^2 + (1 - e2) A^2);  p / cos(lat) = p sqrt(n2) / B
    const double sn = n2 * nr;  // sqrt(n2)
    h = p2 * ip * sn * rb_rcp(B) - KC(earth_r) * sn * rb_rsqrt(fma(B, B, KC(one_m_e2) * (A * A)));
}

// haversine_distance, coordinate_converter.py:125-136 (degrees in, metres out)
RB_HD double haversine(double lat1, double lon1, double lat2, double lon2) {
    const double d2r = RB_PI / 180.0;
    const double la1 = lat1 * d2r, la2 = lat2 * d2r;
    const double sdl = sin((la2 - la1) * 0.5), sdo = sin((lon2 * d2r - lon1 * d2r) * 0.5);
    const double a = sdl * sdl + cos(la1) * cos(la2) * (sdo * sdo);
    return KC(earth_r) * (2.0 * atan2(sqrt(a), sqrt(1.0 - a)));
}

// Density (and temperature) -- atmosphere_model / simplified_nrlmsise_00 / solar_activity_factor
// (spacecraft_model.py:181-188, :71-99, :148-168, :110-127).  solar_time = time-only part of the
// solar factor from the time table.  The two exponentials of the top layer (:84 and :89) are
// one exp of the summed exponent.
// FAST (the hot RHS): the truncated exp / log1p tails of rb_math.cuh (<= 2.2e-15 relative): four FP64 instructions less per
// evaluation.  Everything that multiplies into the drag term keeps full-precision reciprocals and roots.
template <bool FAST = false, class TB = TabGlobal>
RB_HD double density(double altitude, double latitude_factor, double solar_time, double solar_m1, double &T_out, const TB &tb = TB()) {
    // latitude_factor = 1 + 0.01 |lat| / 90 (:76) is formed by the caller (from degrees or radians);
    // solar_m1 = solar_time - 1 (both from the time table)
    // altitude <= 0 as an integer test of the high word (one issue cycle instead of a DSETP's two): |r| - R is either 0
    // or at least 1e-26 in magnitude (never subnormal), so "high word <= 0" is exactly "<= 0.0"; CUDA's canonical NaN has
    // the sign bit set and lands here too -- the lane is non-finite through its gravity term and is retired as such.
#if defined(__CUDA_ARCH__)
    if (FAST ? (__double2hiint(altitude) <= 0) : (altitude <= 0.0)) {
#else
    if (altitude <= 0.0) {   // at or below the sphere of the equatorial radius (never on a live reentry's hot path)
#endif
        T_out = RB_SEA_LEVEL_T;
#if defined(__CUDA_ARCH__)
        // The constant is read by a VOLATILE constant-bank load: as an ordinary literal or constant-bank value the compiler
        // hoists its materialisation (two moves, or a uniform load and two moves) in front of the branch of every
        // evaluation; this way ONE predicated load is issued there.
        static_assert(offsetof(Consts, sea_rho) == 208, "update the offset in the load below");
        double rho0;
        asm volatile("ld.const.f64 %0, [_ZN2rb3c_kE+208];" : "=d"(rho0));
        return rho0;
#else
        return KC(sea_rho);
#endif
    }
    double factor = solar_time;
    // From here on altitude > 0 (or NaN): the thresholds (90 km, the layer breakpoints) have zero low words, so
    // `altitude >= b` is the integer compare of the high words -- one issue cycle instead of a DSETP's two.
    // (CUDA's canonical double NaN has the sign bit set: it compares below every threshold and flows through the
    // blend and layer 0 as NaN.)
#if defined(__CUDA_ARCH__)
    const int ah = __double2hiint(altitude);
#define RB_ALT_GE(b) (ah >= __double2hiint(b))
#else
#define RB_ALT_GE(b) (altitude >= (b))
#endif
    const bool blend = !RB_ALT_GE(RB_SOLAR_BLEND_ALT);   // below 90 km: sigmoid blend of the solar factor, :163-166
    // rho = c0 e^arg (latitude_factor factor): every layer funnels into ONE exp, so that a warp whose lanes sit in
    // different layers (isothermal / gradient / top) does not execute one exponential per layer kind
    double arg, c0, T;
    if (RB_ALT_GE(rb_bp_(7))) {   // top layer (i == 7): isothermal + the extra scale-height tail
        T = KC(top_T);            // P7 / (R_GAS T7) is a constant of the layer
        arg = rb_exp_clamp(fma(KC(top_k), altitude, KC(top_c)));   // unbounded below as the altitude grows
        c0 = KC(top_rho0);
    } else {
        // i = searchsorted(ALTITUDE_BREAKPOINTS, altitude, side='right') - 1 in 0..6 (spacecraft_model.py:79);
        // the exponents of these layers lie in [-3.2, 0]
        int i = RB_ALT_GE(rb_bp_(4)) ? 4 : 0;
#if defined(__CUDA_ARCH__)
        i += (ah >= KC(bp_hi[i + 2])) ? 2 : 0;      // high words of the breakpoints (their low words are zero)
        i += (ah >= KC(bp_hi[i + 1])) ? 1 : 0;
#else
        i += (altitude >= KC(bp[i + 2])) ? 2 : 0;
        i += (altitude >= KC(bp[i + 1])) ? 1 : 0;
#endif
#if defined(__CUDA_ARCH__)
        double lay[6];
        tab_layer(tb, i, lay);
#else
        const double *__restrict__ lay = KC(layer[i]);
#endif
        const double dh = altitude - lay[0];
        const double L = lay[1];
        T = fma(L, dh, lay[2]);
        const double k = lay[4];
        arg = dh * k;                                                   // isothermal: P_i e^(k dh), :84
#if defined(__CUDA_ARCH__)
        const bool gradient = FAST ? ((__double2hiint(L) << 1) != 0) : (L != 0.0);   // (table values: 0 or |L| > 1e-4; integer test)
#else
        const bool gradient = L != 0.0;
#endif
        if (gradient) arg = k * rb_log1p_small<FAST>(lay[3] * dh);      // gradient: P_i (T/T_i)^k, T/T_i = 1 + (L/T_i) dh, :86
        c0 = lay[5] * rb_rcp(T);                                        // P / (R_GAS T), :88
    }
    T_out = T;
    if (blend)   // (the sigmoid's exponent lies in [-2.6, 0.8] for 0 < altitude < 90 km: no clamp)
    {   // 1 + (f - 1) / (1 + e^(-k (x - x0)))
        const double sg = 1.0 + rb_exp_bounded<TB, FAST>(fma(-KC(sig_ks), altitude, KC(sig_c)), tb);
        factor = fma(solar_m1, rb_rcp(sg), 1.0);
    }
    // (evaluating the two exponentials interleaved, to load their coefficients once, measured no gain: 10.14 vs 10.15 G)
    const double ex = rb_exp_bounded<TB, FAST>(arg, tb);
    return (c0 * ex) * (latitude_factor * factor);
#undef RB_ALT_GE
}
// k / |d|^3 from d2 = |d|^2 for the direct third-body terms of the hot RHS: with the MUFU seed y and e = 1 - d2 y^2,
// |d|^-3 = y^3 (1 - e)^(-3/2) = y^3 (1 + 1.5 e + 1.875 e^2 ..): the first-order form has relative error 1.9 e^2 <= 6e-12 of
// accelerations of 1e-6 m/s^2 -- six FP64 instructions instead of the seven of a second-order rsqrt followed by the cube.
RB_HD double third_body_k3(double k, double d2) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d2));
    const double y2 = y * y;
    const double e = fma(-d2, y2, 1.0);
    const double y3k = (k * y) * y2;
    return fma(y3k, e * 1.5, y3k);
#else
    const double id = 1.0 / sqrt(d2);
    return k * (id * id * id);
#endif
}
// bcp: where the body constant bc = Cd A / (2 m) is read from AT ITS USE (the very end of the evaluation): the step
// kernel parks it in shared memory instead of holding two registers through the whole evaluation.
template <bool DIAG, bool SKIP, class TB, bool FAST>
RB_HD void acceleration(const double *__restrict__ e, const double r[3], const double v[3], const double *bcp,
                        double a[3], RhsOut *diag, double *altitude_out, const TB &tb) {
    const double x = r[0], y = r[1], z = r[2];
    const double p2 = fma(x, x, y * y);
    const double r2 = fma(z, z, p2);
    const double ir = rb_rsqrt(r2);                   // 1/|r|
    const double ir2 = ir * ir, ir3 = ir2 * ir;
    // ground velocity (coordinate_converter.py:29-48, :447) and v_rel = v_ground - omega x r_ecef (:448);
    // rotating omega x r_ecef back is omega x r_eci, so v_rel is formed in ECI and |v_rel| is frame free
    const double wx = fma(KC(omega), y, v[0]), wy = fma(-KC(omega), x, v[1]), wz = v[2];
    // point-mass gravity, :451 : -mu r / |r|^3, and J2, :377-388 : f = 1.5 mu J2 R^2 / r^5 ;
    // a = f r (5 z^2/r^2 - {1,1,3}).  Both are (scalar) * r: the scalars are summed first.
    const double gk = KC(neg_mu) * ir3;
    const double zr2 = ir2 * (z * z);
    const double q = 5.0 * zr2;                       // (DIAG form)
    double ax, ay, az;
    if (DIAG) {   // the two terms separately, as the reference reports them
        const double f = KC(j2f) * (ir3 * ir2);
        const double fxy = fma(q, f, -f), fz = fma(q, f, -3.0 * f);
        diag->a_grav[0] = gk * x; diag->a_grav[1] = gk * y; diag->a_grav[2] = gk * z;
        diag->a_j2[0] = fxy * x; diag->a_j2[1] = fxy * y; diag->a_j2[2] = fz * z;
        ax = diag->a_grav[0] + diag->a_j2[0]; ay = diag->a_grav[1] + diag->a_j2[1]; az = diag->a_grav[2] + diag->a_j2[2];
    } else {
        // f = -gk j with j = 1.5 J2 R^2 / r^2:  k_xy = gk (1 + j (1 - q)),  k_z = k_xy + 2 gk j   (q = 5 z^2 / r^2)
        //   = gk + j2c (gk/r^2)(1 - q),  k_z = k_xy + 2 j2c (gk/r^2): one multiply less than forming gj first
        // (written without the factor (1 - q): a DFMA takes ONE immediate, and fma(-5, zr2, 1) made ptxas build 5.0 in a
        // register pair with two moves per evaluation)
        const double gku = gk * ir2;
        const double kxy = fma(KC(j2c5n), gku * zr2, fma(KC(j2c), gku, gk));
        const double kz = fma(KC(j2c2), gku, kxy);
        ax = fma(kxy, x, e[H_NIND + 0]); ay = fma(kxy, y, e[H_NIND + 1]); az = fma(kz, z, e[H_NIND + 2]);
    }
    // third bodies as a lambda: evaluated inside the drag block (fills the latency of its serial tail)
    // or on their own when the drag block is skipped
    auto third_bodies = [&]() {
        // third bodies, :355-363 : k d/|d|^3 - k s/|s|^3 (the second term is time-only -> table)
        if (DIAG) {
            {
                const double dx = e[E_MOON + 0] - x, dy = e[E_MOON + 1] - y, dz = e[E_MOON + 2] - z;
                const double id = rb_rsqrt(fma(dz, dz, fma(dx, dx, dy * dy)));
                const double k3 = KC(moon_k) * (id * id * id);
                const double mx = fma(k3, dx, -e[E_MOON_IND + 0]), my = fma(k3, dy, -e[E_MOON_IND + 1]), mz = fma(k3, dz, -e[E_MOON_IND + 2]);
                ax += mx; ay += my; az += mz;
                diag->a_moon[0] = mx; diag->a_moon[1] = my; diag->a_moon[2] = mz;
            }
            {
                const double dx = e[E_SUN + 0] - x, dy = e[E_SUN + 1] - y, dz = e[E_SUN + 2] - z;
                const double id = rb_rsqrt(fma(dz, dz, fma(dx, dx, dy * dy)));
                const double k3 = KC(sun_k) * (id * id * id);
                const double sx = fma(k3, dx, -e[E_SUN_IND + 0]), sy = fma(k3, dy, -e[E_SUN_IND + 1]), sz = fma(k3, dz, -e[E_SUN_IND + 2]);
                ax += sx; ay += sy; az += sz;
                diag->a_sun[0] = sx; diag->a_sun[1] = sy; diag->a_sun[2] = sz;
            }
        } else {   // the indirect terms are already in (ax, ay, az): accumulate the direct ones
            {
                const double dx = e[H_MOON + 0] - x, dy = e[H_MOON + 1] - y, dz = e[H_MOON + 2] - z;
                const double d2 = fma(dz, dz, fma(dx, dx, dy * dy));
                const double id = FAST ? 0.0 : rb_rsqrt(d2);
                const double k3 = FAST ? third_body_k3(KC(moon_k), d2) : KC(moon_k) * (id * id * id);
                ax = fma(k3, dx, ax); ay = fma(k3, dy, ay); az = fma(k3, dz, az);
            }
            {
                const double dx = e[H_SUN + 0] - x, dy = e[H_SUN + 1] - y, dz = e[H_SUN + 2] - z;
                const double d2 = fma(dz, dz, fma(dx, dx, dy * dy));
                const double id = FAST ? 0.0 : rb_rsqrt(d2);
                const double k3 = FAST ? third_body_k3(KC(sun_k), d2) : KC(sun_k) * (id * id * id);
                ax = fma(k3, dx, ax); ay = fma(k3, dy, ay); az = fma(k3, dz, az);
            }
        }
    };
    // drag, :458-465 : spherical altitude, geodetic latitude (deg), density,
    // a = -(1/2 rho Cd A |v_rel|^2 / |v_rel|) v_rel / m = -(rho bc |v_rel|) v_rel
    const double altitude = fma(r2, ir, -KC(earth_r));   // |r| - R with |r| = r2 / sqrt(r2)  (:458)
    if (altitude_out) *altitude_out = altitude;          // (the step kernel's event function reuses it)
    if (!SKIP || altitude <= RB_NEGLIGIBLE_DRAG_ALTITUDE_) {
        // the ECEF rotation is about z: p = sqrt(xe^2 + ye^2) = sqrt(x^2 + y^2), ze = z
        const double ip = FAST ? rb_rsqrt_q(p2) : rb_rsqrt(p2);
        double gA, gB;
        geodetic_pair<FAST>(p2, ip, z, gA, gB);
        third_bodies();
        double T, rho, lat_deg = 0.0;
        if (DIAG) {
            lat_deg = pair_abs_latitude_deg(gA, gB);
            rho = density<false>(altitude, fma(KC(latc), lat_deg, 1.0), e[E_SOLAR], e[E_SOLAR] - 1.0, T, tb);
        } else {   // radians straight into the latitude factor (the degree conversion folds into the coefficient)
            rho = density<FAST>(altitude, fma(KC(latc_rad), pair_abs_latitude_rad<FAST>(gA, gB, tb), 1.0), e[H_SOLAR], e[H_SOLAR_M1], T, tb);
        }
        const double w2 = fma(wz, wz, fma(wx, wx, wy * wy));
        // (|v_rel|, 1/T and the sigmoid's reciprocal keep the full forms also in the hot RHS: second-order forms -- 1e-12
        // relative of the DRAG term -- were measured: +2 % throughput and reentry parity unchanged (1.9e-13), but on
        // orbits whose perigee dips into the atmosphere the sensitivity of the trajectory amplifies them to 2.4e-9 where a
        // rounding-level implementation stays at 1e-10: not taken)
        const double vn = rb_sqrt(w2);
        const double kd = -(rho * *bcp) * vn;
        if (DIAG) {
            const double ddx = kd * wx, ddy = kd * wy, ddz = kd * wz;
            diag->a_drag[0] = ddx; diag->a_drag[1] = ddy; diag->a_drag[2] = ddz;
        }
        ax = fma(kd, wx, ax); ay = fma(kd, wy, ay); az = fma(kd, wz, az);
        if (DIAG) {
            diag->altitude = altitude; diag->latitude_deg = lat_deg; diag->rho = rho; diag->T = T; diag->v_rel_norm = vn;
        }
    } else {
        // RB_FLAG_SKIP_NEGLIGIBLE_DRAG: above the threshold rho bc v^2 < 1e-40 m/s^2 (density tail of the model)
        third_bodies();
        if (DIAG) {
            diag->a_drag[0] = diag->a_drag[1] = diag->a_drag[2] = 0.0;
            diag->altitude = altitude; diag->latitude_deg = nan(""); diag->rho = 0.0; diag->T = KC(top_T); diag->v_rel_norm = nan("");
        }
    }
    a[0] = ax; a[1] = ay; a[2] = az;
}

// event function of run_simulation.altitude_event (spacecraft_model.py:496-501)
RB_HD double event_g(const double r[3], double event_altitude) {
    const double r2 = fma(r[2], r[2], fma(r[0], r[0], r[1] * r[1]));
    return fma(r2, rb_rsqrt(r2), -KC(earth_r)) - event_altitude;
}

// ---- Dormand-Prince tableau (scipy rk.py:541-565), row 6 of A = B (FSAL) ----
#define RB_DP_A_ROWS                                                                              \
    {0, 0, 0, 0, 0, 0},                                                                           \
    {1.0 / 5, 0, 0, 0, 0, 0},                                                                     \
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},                                                             \
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},                                                   \
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},                        \
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},                 \
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}

#define RB_DP_P_ROWS                                                                                          \
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},                 \
    {0, 0, 0, 0},                                                                                             \
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},             \
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},                  \
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},          \
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},                       \
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}

// The ODE is second order (y' = [v, a(t, r, v)]), so the position rows of every stage vector K_j
// are the stage velocities v_j = v_n + h sum_i a_ji Ka_i.  Substituting them (Nystrom form),
//   v_s = v_n + h   sum_{j<s} a_sj  Ka_j
//   r_s = r_n + h c_s v_n + h^2 sum_{j<s-1} abar_sj Ka_j,   abar_sj = sum_{j<i<s} a_si a_ij,
// is the same Runge-Kutta step (identical in exact arithmetic, rounding-level different from
// scipy's K-matrix form) and only the three acceleration rows of each K_j have to be kept.
struct Tableau {
    double a[7][6];     // a_sj
    double abar[7][6];  // abar_sj
    double c[7];        // c_s = sum_j a_sj
};
constexpr Tableau make_tableau() {
    Tableau t{};
    constexpr double A[7][6] = {RB_DP_A_ROWS};
    for (int s = 0; s < 7; ++s) {
        double cs = 0.0;
        for (int j = 0; j < 6; ++j) { t.a[s][j] = A[s][j]; cs += A[s][j]; t.abar[s][j] = 0.0; }
        t.c[s] = cs;
    }
    for (int s = 0; s < 7; ++s)
        for (int j = 0; j < s; ++j) {
            double v = 0.0;
            for (int i = j + 1; i < s && i < 6; ++i) v += A[s][i] * A[i][j];
            t.abar[s][j] = v;
        }
    // exact stage abscissae of the tableau (the sums above are within 1 ulp of these)
    t.c[0] = 0.0; t.c[1] = 1.0 / 5; t.c[2] = 3.0 / 10; t.c[3] = 4.0 / 5; t.c[4] = 8.0 / 9; t.c[5] = 1.0; t.c[6] = 1.0;
    return t;
}

// The tableau pre-scaled by the (run-constant) step h, built once on the host and passed in the kernel
// parameters (constant bank): stage updates become chains of fma(const, K, acc) with two register operands.
struct HTableau {
    double ha[7][6];    // h a_sj
    double hha[7][6];   // h^2 abar_sj
    double ch[7];       // c_s h
};
inline HTableau make_htableau(double h) {
    const Tableau t = make_tableau();
    HTableau o{};
    for (int s = 0; s < 7; ++s) {
        for (int j = 0; j < 6; ++j) { o.ha[s][j] = h * t.a[s][j]; o.hha[s][j] = (h * h) * t.abar[s][j]; }
        o.ch[s] = t.c[s] * h;
    }
    return o;
}

}  // namespace rb
