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

// time-table entry layout (include/reentry_b200.h)
enum : int {
    E_SUN = 0, E_MOON = 3, E_SUN_IND = 6, E_MOON_IND = 9, E_CG = 12, E_SG = 13, E_SOLAR = 14, E_T = 15,
    ENTRY_DOUBLES = 16
};

struct Body {  // per-trajectory constant bc = 0.5 Cd A / m (SpacecraftModel.Cd / .A / .m, spacecraft_model.py:394-397, :204-206)
    double bc;
};
RB_HD Body make_body(double cd, double area, double mass) { return Body{0.5 * cd * area / mass}; }

struct RhsOut {  // optional diagnostics (dict of spacecraft_model.py:472-480)
    double a_grav[3], a_j2[3], a_moon[3], a_sun[3], a_drag[3];
    double altitude, latitude_deg, rho, T;
};

// USSA-1976-style layer table (constants.py:124-127) and derived per-layer exponents.
// Kept as scalar switch so the values become immediates / constant-bank operands.
struct Layer { double h, L, T, P; };

RB_HD Layer layer_of(double altitude, int &idx) {
    // i = searchsorted(ALTITUDE_BREAKPOINTS, altitude, side='right') - 1   (spacecraft_model.py:79)
    const double bp[RB_N_LAYERS] = RB_ALTITUDE_BREAKPOINTS_INIT;
    const double gr[RB_N_LAYERS] = RB_TEMPERATURE_GRADIENTS_INIT;
    const double bt[RB_N_LAYERS] = RB_BASE_TEMPERATURES_INIT;
    const double bpz[RB_N_LAYERS] = RB_BASE_PRESSURES_INIT;
    int i = 0;
#pragma unroll
    for (int k = 1; k < RB_N_LAYERS; ++k) i += (altitude >= bp[k]) ? 1 : 0;
    idx = i;
    Layer l;
    // select by compare chain (no local-memory indexing)
    l.h = bp[0]; l.L = gr[0]; l.T = bt[0]; l.P = bpz[0];
#pragma unroll
    for (int k = 1; k < RB_N_LAYERS; ++k)
        if (i == k) { l.h = bp[k]; l.L = gr[k]; l.T = bt[k]; l.P = bpz[k]; }
    return l;
}

// |geodetic latitude| in DEGREES of an ECEF point: coordinate_converter.py:81-96 with its quirks
// (theta built with (1 - e2*R); the loop tests |lat - lat_new| < 1e-6 BEFORE lat = lat_new, so
// the PREVIOUS iterate is returned; at most 100 iterations).
//
// B200 form: every angle of the reference's iteration is carried as its (sin, cos) pair.
// atan2(a, b) followed by sin/cos is the normalisation (a, b)/hypot(a, b), and
// |lat - lat_new| < tol is |sin(lat - lat_new)| = |s c' - c s'| < sin(tol) -- the same iterates
// and the same break decision as the reference, with ONE inverse-trig call (the returned
// angle itself) instead of 2 + n atan2 and 1 + n sincos.  Only |lat| is needed by the density
// (spacecraft_model.py:76).  Uses N + h = p / cos(lat) (coordinate_converter.py:89).
//   p2 = x^2 + y^2, inv_p = 1/sqrt(p2) are passed in (shared with the caller).
RB_HD double geodetic_abs_latitude_deg(double p2, double inv_p, double z) {
    const double e2R = RB_EARTH_E2 * RB_EARTH_R;
    const double p = p2 * inv_p;
    // theta = atan2(z R, p (1 - e2 R))
    const double ta = z * RB_EARTH_R, tb = p * (1.0 - e2R);
    const double th = rb_rsqrt(ta * ta + tb * tb);
    const double st = ta * th, ct = tb * th;
    // lat = atan2(z + e2R sin^3 theta, p - e2R cos^3 theta)
    const double la = z + e2R * (st * (st * st)), lb = p - e2R * (ct * (ct * ct));
    const double lh = rb_rsqrt(la * la + lb * lb);
    double s = la * lh, c = lb * lh;
    const double zp = z * inv_p;
    const double sin_tol = 9.999999999998333e-07;  // sin(1e-6)
    for (int k = 0; k < RB_GEODETIC_MAX_ITER; ++k) {
        const double N = RB_EARTH_R * rb_rsqrt(1.0 - RB_EARTH_E2 * (s * s));
        const double b = 1.0 - RB_EARTH_E2 * N * c * inv_p;   // 1 - e2 N / (N + h)
        const double nh = rb_rsqrt(zp * zp + b * b);
        const double s2 = zp * nh, c2 = b * nh;               // lat_new = atan2(z/p, b)
        if (fabs(s * c2 - c * s2) < sin_tol) break;          // |lat - lat_new| < 1e-6
        s = s2; c = c2;
    }
    return atan2(fabs(s), c) * (180.0 / RB_PI);
}

// Density (and temperature) -- atmosphere_model / simplified_nrlmsise_00 / solar_activity_factor
// (spacecraft_model.py:181-188, :71-99, :148-168, :110-127).  solar_time = time-only part of the
// solar factor from the time table.  The two exponentials of the top layer (:84 and :89) are
// one exp of the summed exponent.
RB_HD double density(double altitude, double abs_latitude_deg, double solar_time, double &T_out) {
    if (altitude <= 0.0) { T_out = RB_SEA_LEVEL_T; return RB_SEA_LEVEL_RHO; }
    double factor = solar_time;
    if (altitude < RB_SOLAR_BLEND_ALT) {
        const double ks = RB_SIGMOID_K * RB_SIGMOID_SMOOTHNESS;
        const double normalized = rb_rcp(1.0 + exp(-ks * (altitude - RB_SIGMOID_X0)));
        factor = fma(factor - 1.0, normalized, 1.0);
    }
    const double latitude_factor = fma(RB_LAT_FACTOR_COEF / 90.0, abs_latitude_deg, 1.0);
    const double gM = -RB_EARTH_GRAVITY * RB_AIR_MOLAR_MASS;
    double P, T;
    if (altitude >= 84852.0) {  // top layer (i == 7): isothermal + the extra scale-height tail
        const double dh = altitude - 84852.0;
        const double bt7[RB_N_LAYERS] = RB_BASE_TEMPERATURES_INIT;
        const double bp7[RB_N_LAYERS] = RB_BASE_PRESSURES_INIT;
        T = bt7[7];
        P = bp7[7] * exp(dh * (gM / (RB_GAS_CONSTANT * bt7[7]) - 1.0 / RB_SCALE_HEIGHT));
    } else {
        int i;
        const Layer l = layer_of(altitude, i);
        const double dh = altitude - l.h;
        T = fma(l.L, dh, l.T);
        if (l.L == 0.0) P = l.P * exp(dh * (gM / (RB_GAS_CONSTANT * l.T)));
        else P = l.P * pow(T / l.T, gM / (RB_GAS_CONSTANT * l.L));
    }
    T_out = T;
    return P * latitude_factor * rb_rcp(RB_R_GAS * T) * factor;
}

// Full acceleration of SpacecraftModel.equations_of_motion (spacecraft_model.py:437-470).
// e: time-table entry (16 doubles).  b.mass is used through b.bc = 0.5 Cd A / m.
template <bool DIAG>
RB_HD void acceleration(const double *__restrict__ e, const double r[3], const double v[3], const Body &b,
                        double a[3], RhsOut *diag) {
    const double x = r[0], y = r[1], z = r[2];
    const double p2 = fma(x, x, y * y);
    const double r2 = fma(z, z, p2);
    const double ir = rb_rsqrt(r2);                   // 1/|r|
    const double rn = r2 * ir;                        // |r|  (euclidean_norm, :67-69, :439)
    const double ir2 = ir * ir, ir3 = ir2 * ir;
    const double cg = e[E_CG], sg = e[E_SG];          // cos/sin(gmst0 + EARTH_OMEGA t), :443
    // ground velocity (coordinate_converter.py:29-48, :447) and v_rel = v_ground - omega x r_ecef (:448);
    // rotating omega x r_ecef back is omega x r_eci, so v_rel is formed in ECI and |v_rel| is frame free
    const double wx = v[0] + RB_EARTH_OMEGA * y, wy = v[1] - RB_EARTH_OMEGA * x, wz = v[2];
    // point-mass gravity, :451 : -mu r / |r|^3
    const double gk = -RB_EARTH_MU * ir3;
    double ax = gk * x, ay = gk * y, az = gk * z;
    if (DIAG) { diag->a_grav[0] = ax; diag->a_grav[1] = ay; diag->a_grav[2] = az; }
    // J2, :377-388 : f = 1.5 mu J2 R^2 / r^5 ; a = f r (5 z^2/r^2 - {1,1,3})
    {
        const double f = (1.5 * RB_EARTH_MU * RB_EARTH_J2 * (RB_EARTH_R * RB_EARTH_R)) * (ir3 * ir2);
        const double q = 5.0 * (z * z) * ir2;
        const double fxy = (q - 1.0) * f, fz = (q - 3.0) * f;
        const double jx = fxy * x, jy = fxy * y, jz = fz * z;
        ax += jx; ay += jy; az += jz;
        if (DIAG) { diag->a_j2[0] = jx; diag->a_j2[1] = jy; diag->a_j2[2] = jz; }
    }
    // third bodies, :355-363 : k d/|d|^3 - k s/|s|^3 (the second term is time-only -> table)
    {
        const double dx = e[E_MOON + 0] - x, dy = e[E_MOON + 1] - y, dz = e[E_MOON + 2] - z;
        const double id = rb_rsqrt(fma(dz, dz, fma(dx, dx, dy * dy)));
        const double k3 = RB_MOON_K * (id * id * id);
        const double mx = fma(k3, dx, -e[E_MOON_IND + 0]), my = fma(k3, dy, -e[E_MOON_IND + 1]), mz = fma(k3, dz, -e[E_MOON_IND + 2]);
        ax += mx; ay += my; az += mz;
        if (DIAG) { diag->a_moon[0] = mx; diag->a_moon[1] = my; diag->a_moon[2] = mz; }
    }
    {
        const double dx = e[E_SUN + 0] - x, dy = e[E_SUN + 1] - y, dz = e[E_SUN + 2] - z;
        const double id = rb_rsqrt(fma(dz, dz, fma(dx, dx, dy * dy)));
        const double k3 = RB_SUN_K * (id * id * id);
        const double sx = fma(k3, dx, -e[E_SUN_IND + 0]), sy = fma(k3, dy, -e[E_SUN_IND + 1]), sz = fma(k3, dz, -e[E_SUN_IND + 2]);
        ax += sx; ay += sy; az += sz;
        if (DIAG) { diag->a_sun[0] = sx; diag->a_sun[1] = sy; diag->a_sun[2] = sz; }
    }
    // drag, :458-465 : spherical altitude, geodetic latitude (deg), density,
    // a = -(1/2 rho Cd A |v_rel|^2 / |v_rel|) v_rel / m = -(rho bc |v_rel|) v_rel
    {
        const double altitude = rn - RB_EARTH_R;
        // the ECEF rotation is about z: p = sqrt(xe^2 + ye^2) = sqrt(x^2 + y^2), ze = z
        const double ip = rb_rsqrt(p2);
        const double lat_deg = geodetic_abs_latitude_deg(p2, ip, z);
        double T;
        const double rho = density(altitude, lat_deg, e[E_SOLAR], T);
        const double w2 = fma(wz, wz, fma(wx, wx, wy * wy));
        const double vn = w2 * rb_rsqrt(w2);
        const double kd = -(rho * b.bc) * vn;
        const double ddx = kd * wx, ddy = kd * wy, ddz = kd * wz;
        ax += ddx; ay += ddy; az += ddz;
        if (DIAG) {
            diag->a_drag[0] = ddx; diag->a_drag[1] = ddy; diag->a_drag[2] = ddz;
            diag->altitude = altitude; diag->latitude_deg = lat_deg; diag->rho = rho; diag->T = T;
        }
    }
    (void)cg; (void)sg;
    a[0] = ax; a[1] = ay; a[2] = az;
}

// event function of run_simulation.altitude_event (spacecraft_model.py:496-501)
RB_HD double event_g(const double r[3], double event_altitude) {
    return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) - RB_EARTH_R - event_altitude;
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

}  // namespace rb
