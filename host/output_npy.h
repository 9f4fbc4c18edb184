// Output contract of the drop-in driver (C++14).
//
// The reference writes HDF5 (src/Output_writer.cpp): Fields_NNN.h5 with (nx,ny,nz) datasets + `steps` / `origin` attributes,
// Fields_xy_NNN.h5 with (nx,ny) cuts at z0 + `steps` / `origin` / `plane` / `plane_coordinate`, Photons_NNN.h5, Bunch_NNN.h5.
// Two back ends, selected by the environment variable QPIC_OUTPUT_FORMAT = h5 (default) | npy | both:
//   h5   the same files, written by the self-contained HDF5 writer of output_h5.h (libhdf5 does not exist in this image);
//   npy  one NumPy file per dataset: Fields_000.h5 -> directory Fields_000.h5.d/ with <dataset>.npy and <dataset>.attrs
//        ("name dtype count values..."), the layout oracle/shims/H5Cpp.h produces for the unmodified reference, so both
//        outputs can be diffed directly.
#pragma once

#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "output_h5.h"

namespace qhost {

struct Field_meta {
    std::string name;
    double ox, oy, oz;  // origin (src/System_3d.cpp:176-193)
};

inline void write_npy(const std::string & path, const double * data, const std::vector<size_t> & shape) {
    FILE * f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    std::string sh = "(";
    size_t total = 1;
    for (size_t d : shape) {
        sh += std::to_string(d) + ",";
        total *= d;
    }
    sh += ")";
    const std::string dict = "{'descr': '<f8', 'fortran_order': False, 'shape': " + sh + ", }";
    char header[128];
    std::memset(header, ' ', sizeof(header));
    std::memcpy(header, "\x93NUMPY\x01\x00", 8);
    header[8] = static_cast<char>((sizeof(header) - 10) & 0xff);
    header[9] = static_cast<char>((sizeof(header) - 10) >> 8);
    std::memcpy(header + 10, dict.data(), dict.size());
    header[sizeof(header) - 1] = '\n';
    std::fwrite(header, 1, sizeof(header), f);
    if (total) std::fwrite(data, sizeof(double), total, f);
    std::fclose(f);
}

inline void write_attrs(const std::string & path, const std::vector<double> & steps, const std::vector<double> & origin, bool with_plane,
                        double plane_coordinate) {
    FILE * f = std::fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot create " + path);
    std::fprintf(f, "steps <f8 %zu", steps.size());
    for (double v : steps) std::fprintf(f, " %.17g", v);
    std::fprintf(f, "\norigin <f8 %zu", origin.size());
    for (double v : origin) std::fprintf(f, " %.17g", v);
    std::fprintf(f, "\n");
    if (with_plane) std::fprintf(f, "plane <i4 1 1\nplane_coordinate <f8 1 %.17g\n", plane_coordinate);
    std::fclose(f);
}

class Output_writer {
public:
    Output_writer(bool output3d, bool output_xy, bool output_bunch, double z0, int count, int nx, int ny, int nz, double dx, double dy, double dz)
        : output3d(output3d), output_xy(output_xy), z0(z0), nx(nx), ny(ny), nz(nz), dx(dx), dy(dy), dz(dz) {
        const char * fmt_env = std::getenv("QPIC_OUTPUT_FORMAT");
        const std::string format = fmt_env ? fmt_env : "h5";
        if (format != "h5" && format != "npy" && format != "both") throw std::runtime_error("QPIC_OUTPUT_FORMAT must be h5, npy or both");
        const bool npy = format != "h5", h5 = format != "npy";
        char name[64];
        // (the reference creates the four files in this order, src/Output_writer.cpp:142-156)
        std::snprintf(name, sizeof(name), "Fields_%03d.h5", count);
        if (output3d && npy) mkdir((dir3d = std::string(name) + ".d").c_str(), 0777);
        if (output3d && h5) h5_3d.reset(new H5Writer(name));
        std::snprintf(name, sizeof(name), "Fields_xy_%03d.h5", count);
        if (output_xy && npy) mkdir((dirxy = std::string(name) + ".d").c_str(), 0777);
        if (output_xy && h5) h5_xy.reset(new H5Writer(name));
        std::snprintf(name, sizeof(name), "Bunch_%03d.h5", count);
        if (output_bunch && npy) mkdir((dirbunch = std::string(name) + ".d").c_str(), 0777);
        if (output_bunch && h5) h5_bunch.reset(new H5Writer(name));
        std::snprintf(name, sizeof(name), "Photons_%03d.h5", count);  // always created (src/System_parameters.h:12); empty without qed
        if (npy) mkdir((dirphotons = std::string(name) + ".d").c_str(), 0777);
        if (h5) h5_photons.reset(new H5Writer(name));
    }

    // every file is complete (metadata + superblock written) only after close(); the destructor closes too
    void close() {
        if (h5_3d) h5_3d->close();
        if (h5_xy) h5_xy->close();
        if (h5_bunch) h5_bunch->close();
        if (h5_photons) h5_photons->close();
    }

    // a full (nx,ny,nz) array: whole to Fields, z-interpolated cut to Fields_xy (src/array_utils.cpp:7-24)
    void write_array(const Field_meta & m, const double * a) {
        if (!dir3d.empty()) {
            write_npy(dir3d + "/" + m.name + ".npy", a, {(size_t)nx, (size_t)ny, (size_t)nz});
            write_attrs(dir3d + "/" + m.name + ".attrs", {dx, dy, dz}, {m.ox, m.oy, m.oz}, false, 0);
        }
        if (h5_3d) {  // write_array, src/Output_writer.cpp:7-26
            h5_3d->add_dataset(m.name, {(uint64_t)nx, (uint64_t)ny, (uint64_t)nz}, a,
                               {H5Attribute{"steps", H5Attribute::F64_ARRAY, {dx, dy, dz}, 0}, H5Attribute{"origin", H5Attribute::F64_ARRAY, {m.ox, m.oy, m.oz}, 0}});
        }
        if (output_xy) {
            const int k0 = static_cast<int>(std::floor((z0 - m.oz) / dz));
            const double z_frac = (z0 - m.oz) / dz - k0;
            std::vector<double> cut((size_t)nx * ny);
            for (int i = 0; i < nx; i++) {
                for (int j = 0; j < ny; j++) {
                    const double * row = a + ((size_t)i * ny + j) * nz;
                    cut[(size_t)i * ny + j] = (1 - z_frac) * row[k0] + z_frac * row[k0 + 1];
                }
            }
            write_xy(m, cut.data());
        }
    }

    // a cut that the device already evaluated (qpic_download_xy_cut): Fields_xy only
    void write_xy(const Field_meta & m, const double * cut) {
        if (!dirxy.empty()) {
            write_npy(dirxy + "/" + m.name + ".npy", cut, {(size_t)nx, (size_t)ny});
            write_attrs(dirxy + "/" + m.name + ".attrs", {dx, dy}, {m.ox, m.oy}, true, z0);
        }
        if (h5_xy) {  // write_array with a Plane::XY slice, src/Output_writer.cpp:28-60: plane = static_cast<int>(Plane::XY) = 1
            h5_xy->add_dataset(m.name, {(uint64_t)nx, (uint64_t)ny}, cut,
                               {H5Attribute{"steps", H5Attribute::F64_ARRAY, {dx, dy}, 0}, H5Attribute{"origin", H5Attribute::F64_ARRAY, {m.ox, m.oy}, 0},
                                H5Attribute{"plane", H5Attribute::I32_SCALAR, {}, 1}, H5Attribute{"plane_coordinate", H5Attribute::F64_SCALAR, {z0}, 0}});
        }
    }

    // one 1-D dataset of Bunch_NNN.h5 (write_particle_parameters, src/Output_writer.cpp:190-236)
    void write_bunch_dataset(const std::string & name, const double * values, size_t count) {
        if (!dirbunch.empty()) write_npy(dirbunch + "/" + name + ".npy", values, {count});
        if (h5_bunch) h5_bunch->add_dataset(name, {(uint64_t)count}, values, {});
    }

    // one 1-D dataset of Photons_NNN.h5 (write_photon_parameters, src/Output_writer.cpp:246-252; only written when qed = true)
    void write_photon_dataset(const std::string & name, const double * values, size_t count) {
        if (!dirphotons.empty()) write_npy(dirphotons + "/" + name + ".npy", values, {count});
        if (h5_photons) h5_photons->add_dataset(name, {(uint64_t)count}, values, {});
    }

private:
    bool output3d, output_xy;
    double z0;
    int nx, ny, nz;
    double dx, dy, dz;
    std::string dir3d, dirxy, dirbunch, dirphotons;
    std::unique_ptr<H5Writer> h5_3d, h5_xy, h5_bunch, h5_photons;
};

}  // namespace qhost
