// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Stand-in for the HDF5 C++ API (`H5Cpp.h`), which is not installed in this image.  It lets the
// UNMODIFIED reference sources compile into oracle/_ref/ (see oracle/Makefile) and turns every HDF5
// dataset the reference writes into a NumPy `.npy` file, so tests can np.load() reference output:
//
//     H5File("Fields_000.h5")          ->  directory  Fields_000.h5.d/
//     createDataSet("psi", ...)        ->  Fields_000.h5.d/psi.npy        (zero filled)
//     createAttribute("steps", ...)    ->  Fields_000.h5.d/psi.attrs      (text: name type count values...)
//
// API surface = exactly the calls made by the reference (src/Output_writer.cpp:7-236,
// src/System_3d.cpp:4, src/System_3d.h:3).  All hyperslabs the reference selects are contiguous in a
// row-major dataset ({1,n1,n2} at {i,0,0}, {1,n} at {i,0}, or one element), so a selection is a
// (start offset, element count) pair.
#pragma once

#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned long long hsize_t;

#define H5F_ACC_TRUNC 0x0002u
enum H5S_seloper_t { H5S_SELECT_SET = 0 };

namespace H5 {

class PredType {
public:
    enum Kind { DOUBLE_KIND, INT_KIND };
    explicit PredType(Kind k) : kind(k) {}
    Kind kind;
    size_t size() const { return kind == DOUBLE_KIND ? sizeof(double) : sizeof(int); }
    const char * descr() const { return kind == DOUBLE_KIND ? "<f8" : "<i4"; }
    static const PredType & native_double() {
        static const PredType t(DOUBLE_KIND);
        return t;
    }
    static const PredType & native_int() {
        static const PredType t(INT_KIND);
        return t;
    }
// the reference spells these as static data members
#define NATIVE_DOUBLE native_double()
#define NATIVE_INT native_int()
};

class DataSpace {
public:
    DataSpace() : rank_(0), sel_start_(0), sel_count_(1), has_sel_(false) {}
    DataSpace(int rank, const hsize_t * dims) : rank_(rank), dims_(dims, dims + rank), sel_start_(0), has_sel_(false) {
        sel_count_ = total();
    }

    hsize_t total() const {
        hsize_t t = 1;
        for (auto d : dims_) t *= d;
        return t;
    }

    void selectHyperslab(H5S_seloper_t, const hsize_t * count, const hsize_t * start) {
        // row-major linear offset of start[]; contiguity is asserted below
        hsize_t offset = 0, n = 1;
        bool seen_partial = false;
        for (int a = 0; a < rank_; a++) {
            offset = offset * dims_[a] + start[a];
            n *= count[a];
            if (seen_partial && (count[a] != dims_[a] || start[a] != 0)) {
                throw std::runtime_error("H5 shim: non-contiguous hyperslab");
            }
            if (count[a] != 1) seen_partial = true;
        }
        sel_start_ = offset;
        sel_count_ = n;
        has_sel_ = true;
    }

    void selectElements(H5S_seloper_t, size_t num, const hsize_t * coords) {
        if (num != 1) throw std::runtime_error("H5 shim: only single-element selections");
        hsize_t offset = 0;
        for (int a = 0; a < rank_; a++) offset = offset * dims_[a] + coords[a];
        sel_start_ = offset;
        sel_count_ = 1;
        has_sel_ = true;
    }

    int rank() const { return rank_; }
    const std::vector<hsize_t> & dims() const { return dims_; }
    hsize_t sel_start() const { return sel_start_; }
    hsize_t sel_count() const { return sel_count_; }

private:
    int rank_;
    std::vector<hsize_t> dims_;
    hsize_t sel_start_, sel_count_;
    bool has_sel_;
};

class Attribute {
public:
    Attribute() = default;
    Attribute(std::string path, std::string name, hsize_t count) : path_(std::move(path)), name_(std::move(name)), count_(count) {}

    void write(const PredType & type, const void * buf) {
        FILE * f = std::fopen(path_.c_str(), "a");
        if (!f) throw std::runtime_error("H5 shim: cannot open " + path_);
        std::fprintf(f, "%s %s %llu", name_.c_str(), type.descr(), count_);
        for (hsize_t i = 0; i < count_; i++) {
            if (type.kind == PredType::DOUBLE_KIND) {
                std::fprintf(f, " %.17g", static_cast<const double *>(buf)[i]);
            } else {
                std::fprintf(f, " %d", static_cast<const int *>(buf)[i]);
            }
        }
        std::fprintf(f, "\n");
        std::fclose(f);
    }

private:
    std::string path_, name_;
    hsize_t count_ = 1;
};

class DataSet {
public:
    static const long NPY_HEADER = 128;

    DataSet() = default;
    DataSet(std::string path, DataSpace space, size_t elem) : path_(std::move(path)), space_(std::move(space)), elem_(elem) {}

    // creates the zero-filled .npy
    void create(const PredType & type) {
        FILE * f = std::fopen(path_.c_str(), "wb");
        if (!f) throw std::runtime_error("H5 shim: cannot create " + path_);
        std::string shape = "(";
        for (auto d : space_.dims()) shape += std::to_string(d) + ",";
        shape += ")";
        std::string dict = std::string("{'descr': '") + type.descr() + "', 'fortran_order': False, 'shape': " + shape + ", }";
        char header[NPY_HEADER];
        std::memset(header, ' ', sizeof(header));
        std::memcpy(header, "\x93NUMPY\x01\x00", 8);
        const unsigned short hlen = NPY_HEADER - 10;
        header[8] = static_cast<char>(hlen & 0xff);
        header[9] = static_cast<char>(hlen >> 8);
        if (dict.size() > static_cast<size_t>(NPY_HEADER - 11)) throw std::runtime_error("H5 shim: npy header too long");
        std::memcpy(header + 10, dict.data(), dict.size());
        header[NPY_HEADER - 1] = '\n';
        std::fwrite(header, 1, sizeof(header), f);
        std::fclose(f);
        if (truncate(path_.c_str(), NPY_HEADER + static_cast<off_t>(space_.total() * elem_)) != 0) {
            throw std::runtime_error("H5 shim: cannot size " + path_);
        }
        std::remove((attr_path()).c_str());
    }

    void write(const void * buf, const PredType & type) { write_at(buf, type, 0, space_.total()); }

    void write(const void * buf, const PredType & type, const DataSpace & /*mem*/, const DataSpace & file_space) {
        write_at(buf, type, file_space.sel_start(), file_space.sel_count());
    }

    Attribute createAttribute(const std::string & name, const PredType &, const DataSpace & space) {
        return Attribute(attr_path(), name, space.rank() == 0 ? 1 : space.total());
    }

    DataSpace getSpace() const { return space_; }

private:
    std::string attr_path() const { return path_.substr(0, path_.size() - 4) + ".attrs"; }

    void write_at(const void * buf, const PredType & type, hsize_t start, hsize_t count) {
        if (count == 0) return;
        FILE * f = std::fopen(path_.c_str(), "r+b");
        if (!f) throw std::runtime_error("H5 shim: cannot open " + path_);
        std::fseek(f, NPY_HEADER + static_cast<long>(start * type.size()), SEEK_SET);
        std::fwrite(buf, type.size(), count, f);
        std::fclose(f);
    }

    std::string path_;
    DataSpace space_;
    size_t elem_ = 8;
};

class H5File {
public:
    H5File() = default;
    H5File(const std::string & name, unsigned int /*flags*/) : dir_(name + ".d") {
        mkdir(dir_.c_str(), 0777);
        // an index of datasets: name, rank, dims
        std::remove((dir_ + "/_index.txt").c_str());
    }

    DataSet createDataSet(const std::string & name, const PredType & type, const DataSpace & space) {
        DataSet ds(dir_ + "/" + name + ".npy", space, type.size());
        ds.create(type);
        FILE * f = std::fopen((dir_ + "/_index.txt").c_str(), "a");
        if (f) {
            std::fprintf(f, "%s %d", name.c_str(), space.rank());
            for (auto d : space.dims()) std::fprintf(f, " %llu", d);
            std::fprintf(f, "\n");
            std::fclose(f);
        }
        return ds;
    }

    DataSet openDataSet(const std::string & name) {
        // recover the shape from the index
        FILE * f = std::fopen((dir_ + "/_index.txt").c_str(), "r");
        if (!f) throw std::runtime_error("H5 shim: no index in " + dir_);
        char nm[256];
        int rank;
        DataSet result;
        bool found = false;
        while (std::fscanf(f, "%255s %d", nm, &rank) == 2) {
            std::vector<hsize_t> dims(rank);
            for (int a = 0; a < rank; a++) {
                if (std::fscanf(f, "%llu", &dims[a]) != 1) break;
            }
            if (name == nm) {
                result = DataSet(dir_ + "/" + name + ".npy", DataSpace(rank, dims.data()), 8);
                found = true;
            }
        }
        std::fclose(f);
        if (!found) throw std::runtime_error("H5 shim: dataset " + name + " not found");
        return result;
    }

private:
    std::string dir_;
};

}  // namespace H5
