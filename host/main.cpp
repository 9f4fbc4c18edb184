// quasistatic_pic_b200 <input.toml> -- drop-in driver for laser-wake runs of QuetzalPIC on one B200 (C++14 host code over
// the qpic_b200 C ABI).
//
// Keeps the reference's process-level contracts: the CLI (reference src/quasistatic_pic.cpp:8-46),
// the TOML keys / defaults / stdout echo (src/Config_reader.cpp), the banner printed by System_3d's constructor
// (src/System_3d.cpp:91-152), the per-iteration messages of System_3d::run (:199-236) and the output datasets
// (src/Output_writer.cpp; see output_npy.h for the file format).  solve_wakefield itself runs on the device.
// Beam drivers ([bunch] / [[bunch]]): the particles are built here as System_3d::init_bunch_particles does and live on the device
// (qpic_bunch_*); System_3d::run's time loop (src/System_3d.cpp:199-236) is the loop below, including photon emission (qed = true).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <iostream>
#include <vector>

#include "../include/qpic_b200.h"
#include "config.h"
#include "output_npy.h"

using namespace qhost;

static std::string memory_formatter(long bytes) {  // "{:5.1f} {}" with binary prefixes (src/System_3d.cpp:19-28)
    double res = static_cast<double>(bytes);
    const char * prefixes[] = {"  B", "KiB", "MiB", "GiB", "TiB", "PiB"};
    int index = 0;
    while (res > 1024) {
        res /= 1024;
        index++;
    }
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%5.1f %s", res, prefixes[index]);
    return buf;
}

static void check(int rc, qpic_ctx * ctx) {
    if (rc != QPIC_OK) throw std::runtime_error(qpic_last_error(ctx));
}

int main(int argc, char const * argv[]) {
    if (argc < 2) {
        std::cout << "Provide a path to the input .toml file as the first argument. Aborting." << std::endl;
        return -1;
    }
    if (argc > 2) {
        std::cout << "WARNING: more than one command-line argument are not supported. Additional arguments will be ignored" << std::endl;
    }
    const std::string config_filename = argv[1];
    std::ostream & out = std::cout;
    out << "----------------------------------------" << std::endl;
    const System_parameters params = Config_reader(config_filename, out).get_parameters();
    out << "----------------------------------------" << std::endl;
    out << "Device: one CUDA GPU through libqpic_b200 (ABI " << qpic_abi_version() << ")" << std::endl;
    out << "----------------------------------------" << std::endl;
    const auto t_begin = std::chrono::high_resolution_clock::now();

    if (params.magnetic_field_iterations < 0) throw std::invalid_argument("Magnetic field iterations should be non-negative");
    if (params.ppcy <= 0) throw std::invalid_argument("ppcy should be positive");
    if (params.ppcz <= 0) throw std::invalid_argument("ppcz should be positive");

    // System_3d::System_3d (src/System_3d.cpp:87-104)
    const int nx = static_cast<int>(params.lx / params.dx), ny = static_cast<int>(params.ly / params.dy), nz = static_cast<int>(params.lz / params.dz);
    out << "Spatial dimensions:  [" << nx << ", " << ny << ", " << nz << "]" << std::endl;
    const double dx = params.lx / nx, dy = params.ly / ny, dz = params.lz / nz;
    out << "Steps:               [" << fmt_double(dx) << ", " << fmt_double(dy) << ", " << fmt_double(dz) << "]" << std::endl;
    out << "Simulation box size: [" << fmt_double(params.lx) << ", " << fmt_double(params.ly) << ", " << fmt_double(params.lz) << "]" << std::endl;
    const int time_iterations = static_cast<int>(params.t_end / params.dt) + 1;
    const double t_end = params.dt * (time_iterations - 1);
    out << "Timestep: " << fmt_double(params.dt) << ", end time: " << fmt_double(t_end) << ", iterations: " << time_iterations << std::endl;
    out << "----------------------------------------" << std::endl;
    const Plasma_units units(params.base_frequency_SI);
    out << "PLASMA UNITS\n";
    out << "Frequency " << fmt_g4(units.frequency) << " rad/s\n";
    out << "Density " << fmt_g4(units.density) << " cm^-3\n";
    out << "Wavelength " << fmt_g4(units.wavelength) << " cm\n";
    out << "Schwinger field in plasma units: " << fmt_g4(units.field_schwinger) << "\n";
    out << "----------------------------------------" << std::endl;
    out << "REGISTERED SPECIES\n";
    for (const auto & s : params.species.get_species()) {
        out << (s.is_photon ? s.name + " (is photon)" : s.name) << ": charge " << fmt_g4(s.charge) << ", mass " << fmt_g4(s.mass)
            << ", charge_to_mass_ratio " << fmt_g4(s.charge_to_mass_ratio) << "\n";
    }
    out << "----------------------------------------" << std::endl;

    // beam particles, System_3d::count_bunch_particles / init_bunch_particles (src/System_3d.cpp:750-803): a lattice of ppc points per
    // cell, kept where rho > 0, x-major order, bunch after bunch
    struct {
        std::vector<double> x, y, z, px, py, pz, gamma, n;
        std::vector<int> species_id;
    } beam;
    for (const Bunch_parameters & b : params.bunch_parameters_array) {
        const bool photon = params.species.get_species()[b.species_id].is_photon;
        for (int i = 0; i < b.ppcx * nx; i++) {
            for (int j = 0; j < b.ppcy * ny; j++) {
                for (int k = 0; k < b.ppcz * nz; k++) {
                    const double x = (i + 0.5) * dx / b.ppcx;
                    const double y = (j + 0.5) * dy / b.ppcy;
                    const double z = (k + 0.5) * dz / b.ppcz;
                    const double value = b.rho(x, y, z);
                    if (value > 0.0) {
                        beam.x.push_back(x);
                        beam.y.push_back(y);
                        beam.z.push_back(z);
                        beam.px.push_back(photon ? b.gamma : std::sqrt(b.gamma * b.gamma - 1));
                        beam.py.push_back(0);
                        beam.pz.push_back(0);
                        beam.gamma.push_back(b.gamma);
                        beam.n.push_back(value / b.ppcx / b.ppcy / b.ppcz);
                        beam.species_id.push_back(b.species_id);
                    }
                }
            }
        }
    }
    const size_t beam_count = beam.x.size();

    // plasma macro-particle lattice, System_3d::count_wake_particles / init_wake_particles (src/System_3d.cpp:708-748)
    std::vector<double> py, pz, pn;
    for (int i = 0; i < params.ppcy * ny; i++) {
        for (int j = 0; j < params.ppcz * nz; j++) {
            const double y = (i + 0.5) * dy / params.ppcy;
            const double z = (j + 0.5) * dz / params.ppcz;
            const double value = params.plasma_profile(y, z);
            if (value != 0) {
                py.push_back(y);
                pz.push_back(z);
                pn.push_back(value / params.ppcy / params.ppcz);
            }
        }
    }

    const long fourier_memory = sizeof(double) * (long)ny * nz + 2 * sizeof(double) * (long)ny * (nz / 2 + 1);
    const long array2d_memory = 12l * sizeof(double) * ny * nz;
    const long array3d_memory = 9l * sizeof(double) * nx * ny * nz;
    const long wake_particle_memory = 88l * static_cast<long>(py.size());  // sizeof(wake_particle_2d), src/containers.h:31-43
    const long bunch_particle_memory = 120l * static_cast<long>(beam_count);  // sizeof(bunch_particle_3d), src/containers.h:45-61
    out << "Expected RAM usage:\n";
    out << "3D arrays:       " << memory_formatter(array3d_memory) << "\n"
        << "2D arrays:       " << memory_formatter(array2d_memory) << "\n"
        << "Fourier:         " << memory_formatter(fourier_memory) << "\n"
        << "Wake particles:  " << memory_formatter(wake_particle_memory) << "\n"
        << "Bunch particles: " << memory_formatter(bunch_particle_memory) << "\n"
        << "Total:           " << memory_formatter(array2d_memory + array3d_memory + wake_particle_memory + bunch_particle_memory + fourier_memory) << std::endl;
    out << "----------------------------------------" << std::endl;

    qpic_params p;
    qpic_default_params(&p);
    p.nx = nx;
    p.ny = ny;
    p.nz = nz;
    p.lx = params.lx;
    p.ly = params.ly;
    p.lz = params.lz;
    p.magnetic_field_iterations = params.magnetic_field_iterations;
    const bool any_output = params.output.output3d || params.output.output_xy;
    // Fields_xy alone needs no (nx,ny,nz) per-slice copies: the device collects the cuts slice by slice and evaluates the
    // cuts of the resident arrays itself (qpic_download_xy_cut); Fields needs the full per-slice arrays
    const bool xy_only = params.output.output_xy && !params.output.output3d;
    p.keep_slices_3d = (any_output && !xy_only) ? 1 : 0;
    p.keep_xy_cuts = xy_only ? 1 : 0;
    p.xy_z0 = params.output.z0;
    qpic_ctx * ctx = nullptr;
    if (qpic_create(&p, &ctx) != QPIC_OK) {
        std::cerr << "qpic_create failed: " << qpic_last_error(nullptr) << std::endl;
        return 1;
    }
    try {
        if (params.laser.present) {
            const Laser_gaussian & L = params.laser;
            check(qpic_set_laser_gaussian(ctx, L.amplitude, L.wx, L.wy, L.wz, L.x0, L.y0, L.z0), ctx);
        }
        const bool with_beam = !params.bunch_parameters_array.empty();
        if (with_beam) {
            std::vector<double> charge, cmr;
            std::vector<int> is_photon;
            for (const auto & sp : params.species.get_species()) {
                charge.push_back(sp.charge);
                cmr.push_back(sp.charge_to_mass_ratio);
                is_photon.push_back(sp.is_photon ? 1 : 0);
            }
            check(qpic_bunch_set_species(ctx, (int)charge.size(), charge.data(), cmr.data(), is_photon.data()), ctx);
            check(qpic_bunch_init(ctx, beam_count, beam.x.data(), beam.y.data(), beam.z.data(), beam.px.data(), beam.py.data(), beam.pz.data(),
                                  beam.gamma.data(), beam.n.data(), beam.species_id.data()), ctx);
        }
        for (int ti = 0; ti < time_iterations; ti++) {  // System_3d::run (src/System_3d.cpp:199-236)
            out << "Iteration " << ti << std::endl;
            out << "Depositing bunch densities..." << std::endl;
            if (with_beam) check(qpic_bunch_deposit_densities(ctx), ctx);
            out << "Solving wakefield" << std::flush;
            Output_writer writer(params.output.output3d, params.output.output_xy, params.output.output_bunch, params.output.z0, ti, nx, ny, nz, dx, dy, dz);
            check(qpic_init_wake_particles(ctx, py.data(), pz.data(), pn.data(), py.size()), ctx);
            check(qpic_begin(ctx), ctx);
            int done = 0;
            while (!done) {
                const int before = qpic_current_slice(ctx);  // first slice this call will start
                check(qpic_step(ctx, 10, &done), ctx);
                const int stopped = qpic_stopped_at(ctx);
                const int last_started = done ? (stopped >= 0 ? stopped : 0) : qpic_current_slice(ctx) + 1;
                for (int i = before; i >= last_started; i--) {  // one dot per started slice whose index is divisible by 10 (:329-331)
                    if (i % 10 == 0) out << "." << std::flush;
                }
            }
            if (qpic_stopped_at(ctx) >= 0) out << "Stop condition encountered, stopping the simulation" << std::endl;
            out << std::endl;
            check(qpic_finish(ctx), ctx);
            if (any_output) {
                std::vector<double> buf(xy_only ? 1 : (size_t)nx * ny * nz);
                struct Entry {
                    qpic_field3d id;
                    Field_meta meta;
                };
                const Entry entries[] = {
                    // the five per-slice arrays (initialize_slice_array / write_slice, src/System_3d.cpp:251-260, :663-665)
                    {QPIC_F3_PSI, {"psi", 0, 0, 0}},
                    {QPIC_F3_RHO, {"rho", 0, 0, 0}},
                    {QPIC_F3_JX, {"jx", 0, 0, 0}},
                    {QPIC_F3_JY, {"jy", -0.5 * dx, 0.5 * dy, 0}},
                    {QPIC_F3_JZ, {"jz", -0.5 * dx, 0, 0.5 * dz}},
                    // the nine whole arrays (src/System_3d.cpp:240-249, :624-626)
                    {QPIC_F3_ASQR, {"aSqr", 0, 0, 0}},
                    {QPIC_F3_SUSCEPTIBILITY, {"susceptibility", 0, 0, 0}},
                    {QPIC_F3_RHO_BUNCH, {"rho_bunch", 0, 0, 0}},
                    {QPIC_F3_JX_BUNCH, {"jx_bunch", 0, 0, 0}},
                    {QPIC_F3_EX, {"ex", -0.5 * dx, 0, 0}},
                    {QPIC_F3_EY, {"ey", 0, 0.5 * dy, 0}},
                    {QPIC_F3_EZ, {"ez", 0, 0, 0.5 * dz}},
                    {QPIC_F3_BY, {"by", 0, 0, 0.5 * dz}},
                    {QPIC_F3_BZ, {"bz", 0, 0.5 * dy, 0}},
                };
                if (xy_only) {
                    std::vector<double> cut((size_t)nx * ny);
                    for (const Entry & e : entries) {
                        const bool bunch = (e.id == QPIC_F3_RHO_BUNCH || e.id == QPIC_F3_JX_BUNCH);  // without a beam: zeros (not resident)
                        if (bunch && !with_beam) {
                            std::fill(cut.begin(), cut.end(), 0.0);
                        } else {
                            check(qpic_download_xy_cut(ctx, e.id, 0, nx, cut.data()), ctx);
                            check(qpic_wait_transfers(ctx), ctx);
                        }
                        writer.write_xy(e.meta, cut.data());
                    }
                } else {
                    for (const Entry & e : entries) {
                        check(qpic_download_field3d(ctx, e.id, buf.data()), ctx);
                        writer.write_array(e.meta, buf.data());
                    }
                }
            }
            out << "Depositing fields to particles..." << std::endl;
            if (with_beam) check(qpic_bunch_gather_fields(ctx, units.field_schwinger), ctx);
            if (params.output.output_bunch) {  // write_bunch_parameters (src/Output_writer.cpp:190-244)
                const char * names[] = {"x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi"};
                const qpic_bunch_field ids[] = {QPIC_B_X, QPIC_B_Y, QPIC_B_Z, QPIC_B_PX, QPIC_B_PY, QPIC_B_PZ, QPIC_B_GAMMA,
                                                QPIC_B_EX, QPIC_B_EY, QPIC_B_EZ, QPIC_B_BY, QPIC_B_BZ, QPIC_B_CHI};
                std::vector<double> col(beam_count ? beam_count : 1);
                for (int a = 0; a < 13; a++) {
                    if (with_beam) check(qpic_bunch_download(ctx, ids[a], col.data()), ctx);
                    writer.write_bunch_dataset(names[a], col.data(), beam_count);
                }
                const double weight_norm = dx * dy * dz * units.number_density_norm;
                for (size_t q = 0; q < beam_count; q++) col[q] = beam.n[q] * weight_norm;
                writer.write_bunch_dataset("weight", col.data(), beam_count);
            }
            out << "Advancing particles..." << std::endl;
            if (ti < time_iterations - 1) {
                if (params.qed) {
                    // advance_bunch_particles with qed (src/System_3d.cpp:887-921).  The uniform numbers are drawn here exactly as a
                    // one-thread run of the reference draws them: two rand_double() (src/monte_carlo.cpp:7-9) per charged particle,
                    // in particle order, from the C library generator that the reference never seeds.
                    double max_radiation_probability = 0.0;
                    size_t photon_count = 0;
                    if (with_beam) {
                        std::vector<double> u_emit(beam_count ? beam_count : 1), u_accept(beam_count ? beam_count : 1);
                        for (size_t q = 0; q < beam_count; q++) {
                            if (params.species.get_species()[beam.species_id[q]].is_photon) continue;
                            u_emit[q] = static_cast<double>(rand()) / RAND_MAX;
                            u_accept[q] = static_cast<double>(rand()) / RAND_MAX;
                        }
                        check(qpic_bunch_advance_qed(ctx, params.dt, units.frequency, u_emit.data(), u_accept.data(), 0, &max_radiation_probability, &photon_count), ctx);
                    }
                    out << "Max radiation probability " << fmt_g3(max_radiation_probability) << std::endl;
                    const char * names[] = {"x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi"};
                    const qpic_bunch_field ids[] = {QPIC_B_X, QPIC_B_Y, QPIC_B_Z, QPIC_B_PX, QPIC_B_PY, QPIC_B_PZ, QPIC_B_GAMMA,
                                                    QPIC_B_EX, QPIC_B_EY, QPIC_B_EZ, QPIC_B_BY, QPIC_B_BZ, QPIC_B_CHI};
                    std::vector<double> col(photon_count ? photon_count : 1);
                    for (int a = 0; a < 13; a++) {
                        if (photon_count) check(qpic_photons_download(ctx, ids[a], col.data()), ctx);
                        writer.write_photon_dataset(names[a], col.data(), photon_count);
                    }
                    if (photon_count) check(qpic_photons_download(ctx, QPIC_B_N, col.data()), ctx);
                    const double weight_norm = dx * dy * dz * units.number_density_norm;
                    for (size_t q = 0; q < photon_count; q++) col[q] *= weight_norm;
                    writer.write_photon_dataset("weight", col.data(), photon_count);
                } else if (with_beam) {
                    check(qpic_bunch_advance(ctx, params.dt), ctx);
                }
            }
            writer.close();
            out << std::endl;
        }
    } catch (const std::exception & e) {
        std::cerr << "error: " << e.what() << std::endl;
        qpic_destroy(ctx);
        return 1;
    }
    qpic_destroy(ctx);
    const auto t_done = std::chrono::high_resolution_clock::now();
    out << "----------------------------------------" << std::endl;
    std::printf("Finished in %ld s\n", (long)std::chrono::duration_cast<std::chrono::seconds>(t_done - t_begin).count());
    return 0;
}
