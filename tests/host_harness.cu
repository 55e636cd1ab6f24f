// TEST INFRASTRUCTURE -- compiles the product's __host__ __device__ physics (rb_physics.cuh,
// time_entry of rb_kernels.cuh) for the HOST so that `-m "not gpu"` tests can check the same
// source against the oracle without a GPU.  Never loaded by the product.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "rb_kernels.cuh"

using namespace rb;


extern "C" {

void hh_time_entry(double t, double epoch0, double gmst0, double *e) { time_entry(t, epoch0, gmst0, e); }

void hh_acceleration(const double *e, const double *r, const double *v, double cd, double area, double mass, double *a,
                     double *diag22) {
    const Body b = make_body(cd, area, mass);
    RhsOut d;
    acceleration<true>(e, r, v, b, a, &d);
    if (diag22) {
        for (int c = 0; c < 3; ++c) {
            diag22[c] = a[c]; diag22[3 + c] = d.a_grav[c]; diag22[6 + c] = d.a_j2[c];
            diag22[9 + c] = d.a_moon[c]; diag22[12 + c] = d.a_sun[c]; diag22[15 + c] = d.a_drag[c];
        }
        diag22[18] = d.altitude; diag22[19] = d.latitude_deg; diag22[20] = d.rho; diag22[21] = d.T;
    }
}

double hh_geodetic_latitude_deg(double x, double y, double z) {
    const double p2 = x * x + y * y;
    return geodetic_abs_latitude_deg(p2, rb_rsqrt(p2), z);
}
double hh_density(double alt, double lat, double solar_time, double *T) {
    return density(alt, fma(KC(latc), lat, 1.0), solar_time, solar_time - 1.0, *T);
}

// Same stage arithmetic as k_propagate (Nystrom form, fma chains, table entries), on the host,
// one trajectory.  table: [(5*n_steps+1)][16].  ys_out: [n_steps+1][6].  Returns the impact step or -1.
int64_t hh_propagate(const double *table, const double *y0, double cd, double area, double mass, double h,
                     int64_t n_steps, double event_altitude, double *ys_out) {
    const HTableau ht = make_htableau(h);
    const Body b = make_body(cd, area, mass);
    double y[6], Ka[7][3], ys[6];
    memcpy(y, y0, sizeof y);
    memcpy(ys_out, y, sizeof y);
    double g_old = event_g(y, event_altitude);
    int k0 = 0;
    for (int64_t n = 0; n < n_steps; ++n) {
        for (int s = (n == 0 ? 0 : 1); s <= 6; ++s) {
            const int slot = (s == 0) ? k0 : ((s == 6) ? (6 - k0) : s);
            double k[6][3];
            for (int j = 0; j < 6; ++j) memcpy(k[j], Ka[j == 0 ? k0 : j], sizeof k[j]);
            switch (s) {
                case 0: memcpy(ys, y, sizeof y); break;
                case 1: stage_update<1>(ht, k, y, ys); break;
                case 2: stage_update<2>(ht, k, y, ys); break;
                case 3: stage_update<3>(ht, k, y, ys); break;
                case 4: stage_update<4>(ht, k, y, ys); break;
                case 5: stage_update<5>(ht, k, y, ys); break;
                default: stage_update<6>(ht, k, y, ys); break;
            }
            const int64_t e = 5 * n + ((s == 6) ? 5 : s);
            double hot[HOT_DOUBLES];
            hot_entry(table + e * 16, hot);
            acceleration<false>(hot, ys, ys + 3, b, Ka[slot], nullptr);
        }
        const double g_new = event_g(ys, event_altitude);
        if (g_old >= 0.0 && g_new <= 0.0) return n + 1;
        memcpy(y, ys, sizeof y);
        g_old = g_new;
        k0 = 6 - k0;
        memcpy(ys_out + 6 * (n + 1), y, sizeof y);
    }
    return -1;
}
}
