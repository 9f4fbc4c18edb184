// writes a small file with host/output_h5.h for tests/test_host_and_abi.py (read back by tests/h5mini.py); usage: h5_selftest <dir>
#include <cmath>
#include <string>
#include <vector>

#include "output_h5.h"

using namespace qhost;

int main(int argc, char ** argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    {
        H5Writer w(dir + "/fields.h5");
        std::vector<double> a(3 * 4 * 5), b(3 * 4);
        for (size_t i = 0; i < a.size(); i++) a[i] = std::sin(0.1 * i);
        for (size_t i = 0; i < b.size(); i++) b[i] = -1.5 * i;
        H5Attribute steps{"steps", H5Attribute::F64_ARRAY, {0.1, 0.2, 0.3}, 0}, origin{"origin", H5Attribute::F64_ARRAY, {-0.05, 0.0, 0.15}, 0};
        H5Attribute plane{"plane", H5Attribute::I32_SCALAR, {}, 1}, coord{"plane_coordinate", H5Attribute::F64_SCALAR, {2.5}, 0};
        // deliberately not in name order; the names of a Fields file
        const char * names[] = {"psi", "rho", "jx", "jy", "jz", "aSqr", "susceptibility", "rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz"};
        for (const char * n : names) w.add_dataset(n, {3, 4, 5}, a.data(), {steps, origin});
        w.add_dataset("cut", {3, 4}, b.data(), {steps, origin, plane, coord});
        w.add_dataset("empty", {0}, nullptr, {});
        w.close();
    }
    {
        // the small sample committed as tests/golden/h5_writer_sample.h5 (for readers this image does not have: h5py, h5dump)
        H5Writer w(dir + "/sample.h5");
        const double a[6] = {1.0, 2.0, 3.0, 4.0, 5.0, 6.0}, b[2] = {-0.5, 0.25};
        H5Attribute steps{"steps", H5Attribute::F64_ARRAY, {0.1, 0.2, 0.3}, 0}, origin{"origin", H5Attribute::F64_ARRAY, {-0.05, 0.0, 0.15}, 0};
        H5Attribute plane{"plane", H5Attribute::I32_SCALAR, {}, 1}, coord{"plane_coordinate", H5Attribute::F64_SCALAR, {2.5}, 0};
        w.add_dataset("psi", {1, 2, 3}, a, {steps, origin});
        w.add_dataset("ex", {1, 2}, b, {steps, origin, plane, coord});
        w.add_dataset("x", {0}, nullptr, {});
        w.close();
    }
    {
        H5Writer w(dir + "/nothing.h5");  // Photons_NNN.h5 without qed: a file with an empty root group
        w.close();
    }
    return 0;
}
