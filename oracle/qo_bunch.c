/* TEST INFRASTRUCTURE ONLY -- see qpic_oracle.h for scope and pinning.
 *
 * Plain-C, single-threaded restatement of the beam ("bunch") side of System_3d::run (reference
 * src/System_3d.cpp:199-236): the 8-point deposit of rho_bunch / jx_bunch, the 8-point gather of E and B at the
 * beam particles with the quantum parameter chi, and the beam push (qed = false).  Particle loops run in index
 * order.  Pinned against the unmodified reference by tests/golden/tiny_beam.npz (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "qpic_oracle.h"

/* The reference wraps j1 == -1 (j1 += n) and j2 == n (% n) only, i.e. it is defined for -d <= y - origin < l and indexes out of
 * bounds elsewhere (asserts compiled out).  Beam particles are never re-wrapped by the reference, so the oracle uses the
 * periodic continuation: identical wherever the reference is defined. */
static int periodic_index(int j, int n) {
    j %= n;
    return j < 0 ? j + n : j;
}

/* origin-aware 8-point CIC scatter into an (nx, ny, nz) array: periodic in y and z, dropped outside [0, nx-1) in x;
 * src/array_utils.cpp:290-338 */
void qo_deposit3d(double x, double y, double z, double value, double * a, int nx, int ny, int nz, double dx, double dy, double dz,
                  double ox, double oy, double oz) {
    int i1 = (int)floor((x - ox) / dx);
    int i2 = i1 + 1;
    double x_frac = (x - ox) / dx - i1;
    if (i1 < 0 || i2 >= nx) return;
    int j1 = (int)floor((y - oy) / dy);
    double y_frac = (y - oy) / dy - j1;
    j1 = periodic_index(j1, ny);
    int j2 = (j1 + 1) % ny;
    int k1 = (int)floor((z - oz) / dz);
    double z_frac = (z - oz) / dz - k1;
    k1 = periodic_index(k1, nz);
    int k2 = (k1 + 1) % nz;
#define A3(i, j, k) a[((size_t)(i) * ny + (j)) * nz + (k)]
    A3(i1, j1, k1) += value * (1 - x_frac) * (1 - y_frac) * (1 - z_frac);
    A3(i1, j2, k1) += value * (1 - x_frac) * y_frac * (1 - z_frac);
    A3(i1, j1, k2) += value * (1 - x_frac) * (1 - y_frac) * z_frac;
    A3(i1, j2, k2) += value * (1 - x_frac) * y_frac * z_frac;
    A3(i2, j1, k1) += value * x_frac * (1 - y_frac) * (1 - z_frac);
    A3(i2, j2, k1) += value * x_frac * y_frac * (1 - z_frac);
    A3(i2, j1, k2) += value * x_frac * (1 - y_frac) * z_frac;
    A3(i2, j2, k2) += value * x_frac * y_frac * z_frac;
}

/* origin-aware 8-point gather, 0 outside [0, nx-1) in x; src/array_utils.cpp:340-375 */
double qo_gather3d(double x, double y, double z, const double * a, int nx, int ny, int nz, double dx, double dy, double dz, double ox,
                   double oy, double oz) {
    int i1 = (int)floor((x - ox) / dx);
    int i2 = i1 + 1;
    double x_frac = (x - ox) / dx - i1;
    if (i1 < 0 || i2 >= nx) return 0.0;
    int j1 = (int)floor((y - oy) / dy);
    double y_frac = (y - oy) / dy - j1;
    j1 = periodic_index(j1, ny);
    int j2 = (j1 + 1) % ny;
    int k1 = (int)floor((z - oz) / dz);
    double z_frac = (z - oz) / dz - k1;
    k1 = periodic_index(k1, nz);
    int k2 = (k1 + 1) % nz;
    return A3(i1, j1, k1) * (1 - x_frac) * (1 - y_frac) * (1 - z_frac) + A3(i1, j2, k1) * (1 - x_frac) * y_frac * (1 - z_frac) +
           A3(i1, j1, k2) * (1 - x_frac) * (1 - y_frac) * z_frac + A3(i1, j2, k2) * (1 - x_frac) * y_frac * z_frac +
           A3(i2, j1, k1) * x_frac * (1 - y_frac) * (1 - z_frac) + A3(i2, j2, k1) * x_frac * y_frac * (1 - z_frac) +
           A3(i2, j1, k2) * x_frac * (1 - y_frac) * z_frac + A3(i2, j2, k2) * x_frac * y_frac * z_frac;
#undef A3
}

/* System_3d::deposit_bunch_densities, src/System_3d.cpp:926-946 (rho_bunch, jx_bunch: node centred, :176-177) */
void qo_bunch_deposit_densities(const qo_bunch * b, const double * species_charge, int nx, int ny, int nz, double dx, double dy, double dz,
                                double * rho_bunch, double * jx_bunch) {
    const size_t n3 = (size_t)nx * ny * nz;
    memset(rho_bunch, 0, sizeof(double) * n3);
    memset(jx_bunch, 0, sizeof(double) * n3);
    for (size_t pi = 0; pi < b->count; pi++) {
        const double charge_density = species_charge[b->species_id[pi]] * b->n[pi];
        qo_deposit3d(b->x[pi], b->y[pi], b->z[pi], charge_density, rho_bunch, nx, ny, nz, dx, dy, dz, 0, 0, 0);
        qo_deposit3d(b->x[pi], b->y[pi], b->z[pi], charge_density * b->px[pi] / b->gamma[pi], jx_bunch, nx, ny, nz, dx, dy, dz, 0, 0, 0);
    }
}

/* System_3d::deposit_wakefield_to_bunch, src/System_3d.cpp:843-860; staggering of ex, ey, ez, by, bz: :178-182 */
void qo_bunch_gather_fields(qo_bunch * b, int nx, int ny, int nz, double dx, double dy, double dz, const double * ex, const double * ey,
                            const double * ez, const double * by, const double * bz, double field_schwinger) {
    for (size_t i = 0; i < b->count; i++) {
        const double x = b->x[i], y = b->y[i], z = b->z[i];
        b->ex[i] = qo_gather3d(x, y, z, ex, nx, ny, nz, dx, dy, dz, -0.5 * dx, 0, 0);
        b->ey[i] = qo_gather3d(x, y, z, ey, nx, ny, nz, dx, dy, dz, 0, 0.5 * dy, 0);
        b->ez[i] = qo_gather3d(x, y, z, ez, nx, ny, nz, dx, dy, dz, 0, 0, 0.5 * dz);
        b->by[i] = qo_gather3d(x, y, z, by, nx, ny, nz, dx, dy, dz, 0, 0, 0.5 * dz);
        b->bz[i] = qo_gather3d(x, y, z, bz, nx, ny, nz, dx, dy, dz, 0, 0.5 * dy, 0);
        const double fx = b->gamma[i] * b->ex[i] + b->py[i] * b->bz[i] - b->pz[i] * b->by[i];
        const double fy = b->gamma[i] * b->ey[i] - b->px[i] * b->bz[i];
        const double fz = b->gamma[i] * b->ez[i] + b->px[i] * b->by[i];
        const double f_long = b->px[i] * b->ex[i] + b->py[i] * b->ey[i] + b->pz[i] * b->ez[i];
        b->chi[i] = sqrt(fx * fx + fy * fy + fz * fz - f_long * f_long) / field_schwinger;
    }
}

/* System_3d::advance_bunch_particles with qed == false, src/System_3d.cpp:862-886 */
void qo_bunch_advance(qo_bunch * b, const double * species_charge_to_mass, const int * species_is_photon, double dt) {
    for (size_t i = 0; i < b->count; i++) {
        const int s = b->species_id[i];
        if (species_is_photon[s]) {
            b->x[i] += dt * (b->px[i] / b->gamma[i] - 1);
            b->y[i] += dt * b->py[i] / b->gamma[i];
            b->z[i] += dt * b->pz[i] / b->gamma[i];
            b->gamma[i] = sqrt(b->px[i] * b->px[i] + b->py[i] * b->py[i] + b->pz[i] * b->pz[i]);
        } else {
            const double cmr = species_charge_to_mass[s];
            const double fx = cmr * (b->ex[i] + (b->py[i] * b->bz[i] - b->pz[i] * b->by[i]) / b->gamma[i]);
            const double fy = cmr * (b->ey[i] - b->px[i] * b->bz[i] / b->gamma[i]);
            const double fz = cmr * (b->ez[i] + b->px[i] * b->by[i] / b->gamma[i]);
            b->x[i] += dt * (b->px[i] / b->gamma[i] - 1);
            b->y[i] += dt * b->py[i] / b->gamma[i];
            b->z[i] += dt * b->pz[i] / b->gamma[i];
            b->px[i] += dt * fx;
            b->py[i] += dt * fy;
            b->pz[i] += dt * fz;
            b->gamma[i] = sqrt(1 + b->px[i] * b->px[i] + b->py[i] * b->py[i] + b->pz[i] * b->pz[i]);
        }
    }
}

/* ---------------------------------------------------------------- photon emission (qed = true) */

/* W_new, src/monte_carlo.cpp:11-41: differential emission probability per unit time; note the single-precision powf calls.
 * Constants: src/constants.h:3-12. */
double qo_w_new(double g, double r, double chi, double omega) {
    const double PI = 3.141592653589793238;
    const double ELECTRON_MASS_CGS = 9.1093837015e-28, PLANCK_CONST_CGS = 6.62607015e-27, SPEED_OF_LIGHT_CGS = 29979245800.0;
    const double ELEMENTARY_CHARGE_CGS = 4.803204672997e-10;
    const double PLANCK_CONST_BAR_CGS = PLANCK_CONST_CGS / (2 * PI);
    const double FINE_STRUCTURE_CONSTANT = (ELEMENTARY_CHARGE_CGS * ELEMENTARY_CHARGE_CGS / PLANCK_CONST_BAR_CGS / SPEED_OF_LIGHT_CGS);
    if ((1.0 - r) * g < 1.0 || chi == 0.0 || r == 0.0) return 0.0;
    double rm = 1.0;
    if (!(chi > 0.13333)) { /* :29-33 */
        rm = chi / 0.13333;
        r = r * rm;
    }
    const double coef = ELECTRON_MASS_CGS * SPEED_OF_LIGHT_CGS * SPEED_OF_LIGHT_CGS / (PLANCK_CONST_BAR_CGS * omega);
    const double y = r / ((1.0 - r) * chi);
    const double x = powf(y, 2.0 / 3.0);
    const double exp_v = exp(-2.0 * y / 3.0);
    const double iAiryapp = exp_v / (2.0 * sqrt(PI)) * powf(x + 0.80049, -0.75);
    const double dAiryapp = -exp_v / (2.0 * sqrt(PI)) * powf((x + 0.70861 * (1.0 - 0.65 * x / (1.0 + x * x))), 0.25);
    if (chi > 0.13333) return -FINE_STRUCTURE_CONSTANT * coef / (g * g) * (iAiryapp + (2.0 / x + chi * r * sqrt(x)) * dAiryapp);
    return -rm * FINE_STRUCTURE_CONSTANT * coef / (g * g) * (iAiryapp + (2.0 / x + chi * r * sqrt(x)) * dAiryapp);
}

/* System_3d::advance_bunch_particles with qed == true, src/System_3d.cpp:862-921.  The reference draws two rand_double() per
 * charged particle in particle order (one thread); here the caller supplies them: u_emit[i] (the photon energy fraction r, :890)
 * and u_accept[i] (compared with the emission probability, :897), entries of photon-species particles unused.  Emitted photons
 * are appended to `photons` (capacity >= b->count; x, y, z, px, py, pz, gamma, n are set, the rest stays as it is) in particle
 * order.  Reproduces `p.py -= photon.px` (:917).  Returns the number of photons; *max_probability as the reference's reduction. */
size_t qo_bunch_advance_qed(qo_bunch * b, const double * species_charge_to_mass, const int * species_is_photon, double dt, double frequency,
                            const double * u_emit, const double * u_accept, qo_bunch * photons, double * max_probability) {
    size_t emitted = 0;
    double max_radiation_probability = 0.0;
    for (size_t i = 0; i < b->count; i++) {
        const int s = b->species_id[i];
        if (species_is_photon[s]) {
            b->x[i] += dt * (b->px[i] / b->gamma[i] - 1);
            b->y[i] += dt * b->py[i] / b->gamma[i];
            b->z[i] += dt * b->pz[i] / b->gamma[i];
            b->gamma[i] = sqrt(b->px[i] * b->px[i] + b->py[i] * b->py[i] + b->pz[i] * b->pz[i]);
            continue;
        }
        const double cmr = species_charge_to_mass[s];
        const double fx = cmr * (b->ex[i] + (b->py[i] * b->bz[i] - b->pz[i] * b->by[i]) / b->gamma[i]);
        const double fy = cmr * (b->ey[i] - b->px[i] * b->bz[i] / b->gamma[i]);
        const double fz = cmr * (b->ez[i] + b->px[i] * b->by[i] / b->gamma[i]);
        b->x[i] += dt * (b->px[i] / b->gamma[i] - 1);
        b->y[i] += dt * b->py[i] / b->gamma[i];
        b->z[i] += dt * b->pz[i] / b->gamma[i];
        b->px[i] += dt * fx;
        b->py[i] += dt * fy;
        b->pz[i] += dt * fz;
        b->gamma[i] = sqrt(1 + b->px[i] * b->px[i] + b->py[i] * b->py[i] + b->pz[i] * b->pz[i]);
        double r = u_emit[i];
        const double w = qo_w_new(b->gamma[i], r, b->chi[i], frequency);
        const double radiation_probability = w * dt * b->gamma[i];
        if (radiation_probability > max_radiation_probability) max_radiation_probability = radiation_probability;
        if (u_accept[i] < radiation_probability) {
            if (b->chi[i] < 0.13333) {
                double rm = b->chi[i] / 0.13333;
                r = r * rm;
            }
            const double qx = r * b->px[i], qy = r * b->py[i], qz = r * b->pz[i];
            photons->x[emitted] = b->x[i];
            photons->y[emitted] = b->y[i];
            photons->z[emitted] = b->z[i];
            photons->px[emitted] = qx;
            photons->py[emitted] = qy;
            photons->pz[emitted] = qz;
            photons->n[emitted] = b->n[i];
            photons->gamma[emitted] = sqrt(qx * qx + qy * qy + qz * qz);
            emitted++;
            b->px[i] -= qx;
            b->py[i] -= qx; /* sic, :917 */
            b->pz[i] -= qz;
            b->gamma[i] = sqrt(1 + b->px[i] * b->px[i] + b->py[i] * b->py[i] + b->pz[i] * b->pz[i]);
        }
    }
    photons->count = emitted;
    if (max_probability) *max_probability = max_radiation_probability;
    return emitted;
}
