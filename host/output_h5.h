// Minimal HDF5 writer (C++14, no libhdf5 -- the library does not exist in this image): exactly the subset of the format that
// the reference's Output_writer produces (src/Output_writer.cpp): a flat root group of contiguous little-endian datasets
// (float64, any rank) with small attributes (float64 arrays, a scalar float64, a scalar int32).
//
// Written against the published "HDF5 File Format Specification", the oldest and most widely readable structures:
// superblock version 0, version-1 object headers, a version-1 group B-tree with one symbol-table node, a local heap for the
// link names, version-1 dataspace / datatype / attribute messages, version-3 contiguous layout.  Raw data is streamed to the
// file as datasets are added (a Fields file holds 14 arrays of nx*ny*nz doubles, far more than fits in host memory twice);
// all metadata is appended on close() and the superblock at offset 0 is written last.
//
// tests/h5mini.py is an independent reader of the same subset used by the tests; no libhdf5 / h5py exists in this image to
// cross-check against, which DESIGN.md states.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace qhost {

struct H5Attribute {
    std::string name;
    enum Kind { F64_ARRAY, F64_SCALAR, I32_SCALAR } kind;
    std::vector<double> f64;
    int32_t i32;
};

class H5Writer {
public:
    explicit H5Writer(const std::string & path) : path(path) {
        f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot create " + path);
        std::vector<uint8_t> zero(DATA_START, 0);
        put(zero.data(), zero.size());  // room for the superblock; the data region starts 8-byte aligned behind it
    }
    ~H5Writer() {
        if (f) {
            try {
                close();
            } catch (...) {
            }
        }
    }
    H5Writer(const H5Writer &) = delete;
    H5Writer & operator=(const H5Writer &) = delete;

    // contiguous float64 dataset; `data` holds prod(dims) values (may be NULL when that is 0)
    void add_dataset(const std::string & name, const std::vector<uint64_t> & dims, const double * data, const std::vector<H5Attribute> & attrs) {
        if (datasets.size() >= MAX_LINKS) throw std::runtime_error("H5Writer: too many datasets in " + path);
        for (const Dataset & d : datasets) {
            if (d.name == name) throw std::runtime_error("H5Writer: dataset " + name + " exists already in " + path);
        }
        Dataset d;
        d.name = name;
        d.dims = dims;
        d.attrs = attrs;
        uint64_t count = 1;
        for (uint64_t n : dims) count *= n;
        d.bytes = count * 8;
        d.data_addr = d.bytes ? pos : UNDEF;
        if (d.bytes) put(data, d.bytes);
        datasets.push_back(d);
    }

    void close() {
        if (!f) return;
        std::sort(datasets.begin(), datasets.end(), [](const Dataset & a, const Dataset & b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
        // ---- local heap data segment: "" at offset 0, then the link names (NUL terminated, 8-byte aligned), then one free block
        std::vector<uint8_t> heap(8, 0);
        for (Dataset & d : datasets) {
            d.name_offset = heap.size();
            heap.insert(heap.end(), d.name.begin(), d.name.end());
            heap.push_back(0);
            while (heap.size() % 8) heap.push_back(0);
        }
        const uint64_t free_offset = heap.size();
        const uint64_t free_size = 32;
        heap.resize(heap.size() + free_size, 0);
        le(heap.data() + free_offset, 1, 8);              // next free block: 1 = end of the free list (H5HL_FREE_NULL)
        le(heap.data() + free_offset + 8, free_size, 8);  // size of this free block
        // ---- dataset object headers
        align8();
        for (Dataset & d : datasets) {
            d.header_addr = pos;
            const std::vector<uint8_t> h = dataset_header(d);
            put(h.data(), h.size());
        }
        // ---- heap
        align8();
        const uint64_t heap_data_addr = pos;
        put(heap.data(), heap.size());
        const uint64_t heap_addr = pos;
        {
            std::vector<uint8_t> h(32, 0);
            std::memcpy(h.data(), "HEAP", 4);  // version 0, 3 reserved bytes
            le(h.data() + 8, heap.size(), 8);
            le(h.data() + 16, free_offset, 8);
            le(h.data() + 24, heap_data_addr, 8);
            put(h.data(), h.size());
        }
        // ---- symbol table node with every link (capacity 2 * LEAF_K entries of 40 bytes)
        uint64_t snod_addr = UNDEF;
        if (!datasets.empty()) {
            snod_addr = pos;
            std::vector<uint8_t> s(8 + 2 * LEAF_K * 40, 0);
            std::memcpy(s.data(), "SNOD", 4);
            s[4] = 1;  // version
            le(s.data() + 6, datasets.size(), 2);
            for (size_t i = 0; i < datasets.size(); i++) {
                uint8_t * e = s.data() + 8 + 40 * i;
                le(e, datasets[i].name_offset, 8);
                le(e + 8, datasets[i].header_addr, 8);
                // cache type 0 (nothing cached), reserved, 16 bytes of scratch: zeros
            }
            put(s.data(), s.size());
        }
        // ---- group B-tree: one leaf node (level 0) with at most one child
        const uint64_t btree_addr = pos;
        {
            std::vector<uint8_t> b(24 + 2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8, 0);
            std::memcpy(b.data(), "TREE", 4);
            b[4] = 0;  // node type 0: group node
            b[5] = 0;  // level 0
            le(b.data() + 6, datasets.empty() ? 0 : 1, 2);
            le(b.data() + 8, UNDEF, 8);   // left sibling
            le(b.data() + 16, UNDEF, 8);  // right sibling
            le(b.data() + 24, 0, 8);      // key 0: heap offset of "" (smaller than every name)
            if (!datasets.empty()) {
                le(b.data() + 32, snod_addr, 8);                    // child 0
                le(b.data() + 40, datasets.back().name_offset, 8);  // key 1: the largest name in child 0
            }
            put(b.data(), b.size());
        }
        // ---- root group object header: one symbol table message
        const uint64_t root_addr = pos;
        {
            std::vector<uint8_t> m(16, 0);
            le(m.data(), btree_addr, 8);
            le(m.data() + 8, heap_addr, 8);
            std::vector<std::vector<uint8_t>> msgs = {message(0x0011, m)};
            const std::vector<uint8_t> h = object_header(msgs);
            put(h.data(), h.size());
        }
        const uint64_t eof = pos;
        // ---- superblock, version 0
        std::vector<uint8_t> sb(96, 0);
        const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        std::memcpy(sb.data(), sig, 8);
        // versions of superblock, free-space storage, root group symbol table entry, reserved, shared header message format: 0
        sb[13] = 8;  // size of offsets
        sb[14] = 8;  // size of lengths
        le(sb.data() + 16, LEAF_K, 2);
        le(sb.data() + 18, INTERNAL_K, 2);
        le(sb.data() + 20, 0, 4);       // file consistency flags
        le(sb.data() + 24, 0, 8);       // base address
        le(sb.data() + 32, UNDEF, 8);   // free-space info
        le(sb.data() + 40, eof, 8);     // end of file address
        le(sb.data() + 48, UNDEF, 8);   // driver information block
        le(sb.data() + 56, 0, 8);       // root entry: link name offset
        le(sb.data() + 64, root_addr, 8);
        le(sb.data() + 72, 1, 4);       // cache type 1: the scratch pad holds the group's B-tree and heap addresses
        le(sb.data() + 80, btree_addr, 8);
        le(sb.data() + 88, heap_addr, 8);
        if (std::fseek(f, 0, SEEK_SET) != 0) throw std::runtime_error("seek failed on " + path);
        if (std::fwrite(sb.data(), 1, sb.size(), f) != sb.size()) throw std::runtime_error("write failed on " + path);
        const bool ok = std::fclose(f) == 0;
        f = nullptr;
        if (!ok) throw std::runtime_error("close failed on " + path);
    }

private:
    static constexpr uint64_t UNDEF = ~0ull;
    static constexpr uint64_t DATA_START = 2048;
    static constexpr unsigned LEAF_K = 16, INTERNAL_K = 16, MAX_LINKS = 2 * LEAF_K;

    struct Dataset {
        std::string name;
        std::vector<uint64_t> dims;
        std::vector<H5Attribute> attrs;
        uint64_t bytes = 0, data_addr = 0, header_addr = 0, name_offset = 0;
    };

    static void le(uint8_t * p, uint64_t v, int bytes) {
        for (int i = 0; i < bytes; i++) p[i] = static_cast<uint8_t>(v >> (8 * i));
    }
    static void pad8(std::vector<uint8_t> & v) {
        while (v.size() % 8) v.push_back(0);
    }
    void put(const void * p, size_t n) {
        if (n && std::fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed on " + path);
        pos += n;
    }
    void align8() {
        const uint8_t zero[8] = {0};
        if (pos % 8) put(zero, 8 - pos % 8);
    }

    // header message: type, size of the (padded) body, flags 0, 3 reserved bytes, body
    static std::vector<uint8_t> message(unsigned type, std::vector<uint8_t> body) {
        pad8(body);
        std::vector<uint8_t> m(8, 0);
        le(m.data(), type, 2);
        le(m.data() + 2, body.size(), 2);
        m.insert(m.end(), body.begin(), body.end());
        return m;
    }
    // version-1 object header: 12-byte prefix + 4 bytes of alignment, then the messages
    static std::vector<uint8_t> object_header(const std::vector<std::vector<uint8_t>> & msgs) {
        size_t total = 0;
        for (const auto & m : msgs) total += m.size();
        std::vector<uint8_t> h(16, 0);
        h[0] = 1;
        le(h.data() + 2, msgs.size(), 2);
        le(h.data() + 4, 1, 4);      // object reference count
        le(h.data() + 8, total, 4);  // size of the header's message area
        for (const auto & m : msgs) h.insert(h.end(), m.begin(), m.end());
        return h;
    }
    // dataspace message body, version 1 (rank 0 = scalar)
    static std::vector<uint8_t> dataspace(const std::vector<uint64_t> & dims) {
        std::vector<uint8_t> b(8 + 8 * dims.size(), 0);
        b[0] = 1;
        b[1] = static_cast<uint8_t>(dims.size());
        for (size_t i = 0; i < dims.size(); i++) le(b.data() + 8 + 8 * i, dims[i], 8);
        return b;
    }
    // datatype message bodies, version 1: IEEE little-endian float64 / two's complement little-endian int32
    static std::vector<uint8_t> type_f64() {
        const uint8_t t[20] = {0x11, 0x20, 0x3f, 0x00, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xff, 0x03, 0, 0};
        return std::vector<uint8_t>(t, t + 20);
    }
    static std::vector<uint8_t> type_i32() {
        const uint8_t t[12] = {0x10, 0x08, 0x00, 0x00, 4, 0, 0, 0, 0, 0, 32, 0};
        return std::vector<uint8_t>(t, t + 12);
    }
    // attribute message body, version 1: name, datatype and dataspace each padded to 8 bytes, then the value
    static std::vector<uint8_t> attribute(const H5Attribute & a) {
        std::vector<uint8_t> type = a.kind == H5Attribute::I32_SCALAR ? type_i32() : type_f64();
        std::vector<uint8_t> space = dataspace(a.kind == H5Attribute::F64_ARRAY ? std::vector<uint64_t>{a.f64.size()} : std::vector<uint64_t>{});
        std::vector<uint8_t> b(8, 0);
        b[0] = 1;
        le(b.data() + 2, a.name.size() + 1, 2);
        le(b.data() + 4, type.size(), 2);
        le(b.data() + 6, space.size(), 2);
        b.insert(b.end(), a.name.begin(), a.name.end());
        b.push_back(0);
        pad8(b);
        b.insert(b.end(), type.begin(), type.end());
        pad8(b);
        b.insert(b.end(), space.begin(), space.end());
        pad8(b);
        if (a.kind == H5Attribute::I32_SCALAR) {
            uint8_t v[4];
            le(v, static_cast<uint32_t>(a.i32), 4);
            b.insert(b.end(), v, v + 4);
        } else {
            for (double x : a.f64) {
                uint8_t v[8];
                std::memcpy(v, &x, 8);
                b.insert(b.end(), v, v + 8);
            }
        }
        return b;
    }
    static std::vector<uint8_t> dataset_header(const Dataset & d) {
        std::vector<std::vector<uint8_t>> msgs;
        msgs.push_back(message(0x0001, dataspace(d.dims)));
        msgs.push_back(message(0x0003, type_f64()));
        // fill value, version 2: space allocated early, fill written "if set", no fill value defined
        msgs.push_back(message(0x0005, std::vector<uint8_t>{2, 1, 2, 0}));
        std::vector<uint8_t> layout(18, 0);
        layout[0] = 3;  // version 3
        layout[1] = 1;  // contiguous
        le(layout.data() + 2, d.data_addr, 8);
        le(layout.data() + 10, d.bytes, 8);
        msgs.push_back(message(0x0008, layout));
        for (const H5Attribute & a : d.attrs) msgs.push_back(message(0x000C, attribute(a)));
        return object_header(msgs);
    }

    std::string path;
    FILE * f = nullptr;
    uint64_t pos = 0;
    std::vector<Dataset> datasets;
};

}  // namespace qhost
