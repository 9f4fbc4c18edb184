"""The C++14 host driver (host/): TOML input, stdout and output-file contracts of the reference executable.

CPU part: the configuration echo and the System_3d banner must be character-identical to what the UNMODIFIED reference
printed for the same input (recorded in tests/golden/*.npz["stdout"]).  GPU part: the files the driver writes must match
the reference's datasets."""
import os
import subprocess

import numpy as np
import pytest

from cases import slice_rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "host", "_build", "quasistatic_pic_b200")
NAMES = ["aSqr", "susceptibility", "rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz", "psi", "rho", "jx", "jy", "jz"]


def build_driver():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "all"], stdout=subprocess.DEVNULL)


def run_driver(name, workdir, golden_dir):
    build_driver()
    toml = os.path.join(golden_dir, "inputs", name + ".toml")
    subprocess.check_call(["cp", toml, workdir])
    # both output back ends: the HDF5 files (default, what the reference writes) and the .npy directories the goldens were made from
    return subprocess.run([DRIVER, name + ".toml"], cwd=workdir, capture_output=True, text=True, env=dict(os.environ, QPIC_OUTPUT_FORMAT="both"))


def check_h5_matches_npy(workdir, stem, names, attrs=()):
    """the .h5 file holds exactly the datasets (and attributes) of the .npy directory"""
    from h5mini import H5File
    f = H5File(os.path.join(workdir, stem + ".h5"))
    assert f.keys() == sorted(names)
    for nm in names:
        assert np.array_equal(f[nm], np.load(os.path.join(workdir, stem + ".h5.d", nm + ".npy"))), (stem, nm)
        for a in attrs:
            assert a in f.attrs(nm), (stem, nm, a)
    return f


def split_sections(text):
    return text.split("----------------------------------------\n")


@pytest.mark.parametrize("name", ["tiny_laser", "tiny_profile", "laser_test", "tiny_bunch", "tiny_bunch_stop", "tiny_beam", "tiny_qed"])
def test_config_echo_and_banner_match_reference_stdout(name, golden_dir, tmp_path):
    gname = "laser_test_summary" if name == "laser_test" else name
    ref = str(np.load(os.path.join(golden_dir, gname + ".npz"))["stdout"])
    got = run_driver(name, str(tmp_path), golden_dir).stdout
    ref_s, got_s = split_sections(ref), split_sections(got)
    # section 1: "Reading config ..." + every key echoed with its default marker (src/Config_reader.cpp:85,101-104)
    assert got_s[1] == ref_s[1]
    # section 2 is the thread / device line; sections 3-5: grid + units + species banner (src/System_3d.cpp:91-120)
    assert got_s[3:6] == ref_s[3:6]
    # section 6: memory estimate incl. the beam particles the driver builds on the host (src/System_3d.cpp:122-150)
    assert got_s[6] == ref_s[6]


def test_missing_argument_and_bad_config(tmp_path):
    build_driver()
    r = subprocess.run([DRIVER], capture_output=True, text=True)
    assert r.returncode != 0 and "Provide a path to the input .toml file" in r.stdout  # src/quasistatic_pic.cpp:11-13
    bad = tmp_path / "bad.toml"
    bad.write_text("lx = 1\nly = 1\nlz = 1\ndx = 0.1\ndy = 0.1\n")  # dz missing
    r = subprocess.run([DRIVER, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "Value [dz] doesn't exist or is of wrong type" in (r.stdout + r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_laser", "tiny_profile"])
def test_driver_output_files_match_reference(name, golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    r = run_driver(name, str(tmp_path), golden_dir)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Solving wakefield.." in r.stdout and "Finished in" in r.stdout
    for nm in NAMES:
        a = np.load(os.path.join(tmp_path, "Fields_000.h5.d", nm + ".npy"))
        assert a.shape == g["f3/" + nm].shape
        assert slice_rel_l2(a, g["f3/" + nm]) < 1e-9, nm
        xy = np.load(os.path.join(tmp_path, "Fields_xy_000.h5.d", nm + ".npy"))
        ref = g["xy/" + nm]
        # the cut sits on the symmetry plane, where odd fields (ez, by, jz) are pure rounding noise: scale by the 3-D field
        scale = np.abs(g["f3/" + nm]).max()
        assert np.abs(xy - ref).max() <= 1e-9 * scale + 1e-300, nm
        attrs = dict(line.split(None, 3)[0::3] for line in open(os.path.join(tmp_path, "Fields_000.h5.d", nm + ".attrs")))
        assert np.allclose([float(v) for v in attrs["origin"].split()], g[f"f3/{nm}@origin"], rtol=0, atol=1e-15)
        assert np.allclose([float(v) for v in attrs["steps"].split()], g[f"f3/{nm}@steps"], rtol=0, atol=1e-15)
    assert os.path.isdir(os.path.join(tmp_path, "Photons_000.h5.d"))
    # the HDF5 files: same datasets bit for bit, attributes of src/Output_writer.cpp:7-60
    f3 = check_h5_matches_npy(str(tmp_path), "Fields_000", NAMES, ("steps", "origin"))
    assert np.allclose(f3.attrs("jy")["origin"], g["f3/jy@origin"], rtol=0, atol=1e-15) and f3.attrs("jy")["steps"].shape == (3,)
    fxy = check_h5_matches_npy(str(tmp_path), "Fields_xy_000", NAMES, ("steps", "origin", "plane", "plane_coordinate"))
    assert fxy.attrs("psi")["plane"] == 1 and fxy.attrs("psi")["plane_coordinate"] == float(g["xy/psi@plane_coordinate"][0])
    assert fxy.attrs("ez")["steps"].shape == (2,)
    from h5mini import H5File
    assert H5File(os.path.join(tmp_path, "Photons_000.h5")).keys() == []


@pytest.mark.gpu
def test_driver_fields_xy_only_uses_device_side_cuts(golden_dir, tmp_path):
    """fields_xy without fields_3d: no (nx,ny,nz) per-slice copies, the cuts come from qpic_download_xy_cut
    (calculate_xy_slice / calculate_y_slice, src/array_utils.cpp:7-45); same datasets as the reference's Fields_xy file."""
    build_driver()
    name = "tiny_laser"
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    text = open(os.path.join(golden_dir, "inputs", name + ".toml")).read().replace("fields_3d = true", "fields_3d = false")
    with open(os.path.join(tmp_path, name + ".toml"), "w") as f:
        f.write(text)
    r = subprocess.run([DRIVER, name + ".toml"], cwd=str(tmp_path), capture_output=True, text=True, env=dict(os.environ, QPIC_OUTPUT_FORMAT="both"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert not os.path.isdir(os.path.join(tmp_path, "Fields_000.h5.d")) and not os.path.exists(os.path.join(tmp_path, "Fields_000.h5"))
    for nm in NAMES:
        xy = np.load(os.path.join(tmp_path, "Fields_xy_000.h5.d", nm + ".npy"))
        ref = g["xy/" + nm]
        assert xy.shape == ref.shape
        scale = np.abs(g["f3/" + nm]).max()
        assert np.abs(xy - ref).max() <= 1e-9 * scale + 1e-300, nm


@pytest.mark.gpu
def test_driver_beam_time_loop_matches_reference(golden_dir, tmp_path):
    """[[bunch]] input with t_end > 0: System_3d::run's time loop (src/System_3d.cpp:199-236) -- Fields_NNN, Bunch_NNN and the stdout
    of every iteration against the unmodified reference"""
    g = np.load(os.path.join(golden_dir, "tiny_beam.npz"))
    r = run_driver("tiny_beam", str(tmp_path), golden_dir)
    assert r.returncode == 0, r.stdout + r.stderr
    ref_s, got_s = split_sections(str(g["stdout"])), split_sections(r.stdout)
    assert got_s[7] == ref_s[7]  # "Iteration 0 ... Advancing particles..." x 3
    for t in range(3):
        for nm in ["x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi", "weight"]:
            a = np.load(os.path.join(tmp_path, f"Bunch_{t:03d}.h5.d", nm + ".npy"))
            b = g[f"it{t}/bunch/{nm}"]
            assert a.shape == b.shape
            assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(b), (t, nm)
        for nm in NAMES:
            a = np.load(os.path.join(tmp_path, f"Fields_{t:03d}.h5.d", nm + ".npy"))
            assert slice_rel_l2(a, g[f"it{t}/f3/{nm}"]) < 1e-9, (t, nm)
        assert os.path.isdir(os.path.join(tmp_path, f"Photons_{t:03d}.h5.d"))
        check_h5_matches_npy(str(tmp_path), f"Bunch_{t:03d}", ["x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi", "weight"])
        check_h5_matches_npy(str(tmp_path), f"Fields_{t:03d}", NAMES, ("steps", "origin"))


@pytest.mark.gpu
def test_driver_qed_run_matches_reference(golden_dir, tmp_path):
    """qed = true: the driver draws the uniform numbers as a one-thread reference run does (glibc rand(), never seeded), so
    Bunch_NNN, Photons_NNN (same photons, same order) and the "Max radiation probability" lines match the unmodified reference"""
    from h5mini import H5File
    g = np.load(os.path.join(golden_dir, "tiny_qed.npz"))
    r = run_driver("tiny_qed", str(tmp_path), golden_dir)
    assert r.returncode == 0, r.stdout + r.stderr
    ref_s, got_s = split_sections(str(g["stdout"])), split_sections(r.stdout)
    assert got_s[7] == ref_s[7]
    names = ["x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi", "weight"]
    for t in range(3):
        for nm in names:
            a = np.load(os.path.join(tmp_path, f"Bunch_{t:03d}.h5.d", nm + ".npy"))
            b = g[f"it{t}/bunch/{nm}"]
            assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(b), (t, nm)
        f = H5File(os.path.join(tmp_path, f"Photons_{t:03d}.h5"))
        if t == 2:
            assert f.keys() == []  # no push after the last iteration (src/System_3d.cpp:222)
            continue
        assert f.keys() == sorted(names)
        for nm in names:
            b = g[f"it{t}/photons/{nm}"]
            assert f[nm].shape == b.shape and np.linalg.norm(f[nm] - b) <= 1e-9 * np.linalg.norm(b) + 1e-300, (t, nm)


def test_hdf5_writer_subset_reads_back(tmp_path):
    """host/output_h5.h (superblock v0, v1 object headers, group B-tree / symbol table node / local heap, contiguous float64
    datasets, v1 attributes) read back by the independent reader tests/h5mini.py; known-answer bytes of the two datatype messages
    and the structural invariants a libhdf5 reader relies on"""
    from h5mini import H5File
    build_driver()
    subprocess.check_call([os.path.join(ROOT, "host", "_build", "h5_selftest"), str(tmp_path)])
    f = H5File(os.path.join(tmp_path, "fields.h5"))
    names = ["psi", "rho", "jx", "jy", "jz", "aSqr", "susceptibility", "rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz"]
    assert f.keys() == sorted(names + ["cut", "empty"])  # strcmp order: "aSqr" < "by" < ... < "susceptibility"
    ref = np.sin(0.1 * np.arange(60)).reshape(3, 4, 5)
    for nm in names:
        assert np.array_equal(f[nm], ref)
        assert np.array_equal(f.attrs(nm)["steps"], [0.1, 0.2, 0.3]) and np.array_equal(f.attrs(nm)["origin"], [-0.05, 0.0, 0.15])
    assert np.array_equal(f["cut"], (-1.5 * np.arange(12)).reshape(3, 4))
    a = f.attrs("cut")
    assert a["plane"] == 1 and a["plane"].dtype == np.int32 and a["plane_coordinate"] == 2.5
    assert f["empty"].shape == (0,)
    raw = open(os.path.join(tmp_path, "fields.h5"), "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and len(raw) % 8 == 0
    # IEEE little-endian float64 / int32 datatype messages exactly as libhdf5 encodes H5T_NATIVE_DOUBLE / H5T_NATIVE_INT on x86-64
    assert bytes.fromhex("11203f00080000000000400034 0b003 4ff030000".replace(" ", "")) in raw
    assert bytes.fromhex("100800000400000000002000") in raw
    assert H5File(os.path.join(tmp_path, "nothing.h5")).keys() == []


def test_hdf5_file_structures_follow_the_format_specification(tmp_path, golden_dir):
    """Byte-level check of host/output_h5.h against the structures a libhdf5 / h5py reader parses first ("HDF5 File Format
    Specification", version-0 superblock, version-1 B-tree / symbol table node / local heap / object headers), field by field with
    struct offsets taken from the specification's tables -- independent of tests/h5mini.py.  The writer's output for a fixed small
    input is also committed as tests/golden/h5_writer_sample.h5 (so that a later environment WITH h5py can open it: datasets "psi"
    (1, 2, 3) = 1..6 with attributes steps / origin, "ex" (1, 2) with plane / plane_coordinate as well, empty "x") and must be
    reproduced byte for byte."""
    import struct
    build_driver()
    subprocess.check_call([os.path.join(ROOT, "host", "_build", "h5_selftest"), str(tmp_path)])
    raw = open(os.path.join(tmp_path, "sample.h5"), "rb").read()
    assert raw == open(os.path.join(golden_dir, "h5_writer_sample.h5"), "rb").read()
    UNDEF = 0xFFFFFFFFFFFFFFFF
    u8 = lambda o: raw[o]
    u16 = lambda o: struct.unpack_from("<H", raw, o)[0]
    u32 = lambda o: struct.unpack_from("<I", raw, o)[0]
    u64 = lambda o: struct.unpack_from("<Q", raw, o)[0]
    # ---- superblock, version 0 (spec II.A)
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert [u8(8), u8(9), u8(10), u8(11), u8(12)] == [0, 0, 0, 0, 0]      # superblock / free-space / root entry / reserved / shared header versions
    assert (u8(13), u8(14), u8(15)) == (8, 8, 0)                          # size of offsets, size of lengths, reserved
    leaf_k, internal_k = u16(16), u16(18)
    assert leaf_k >= 1 and internal_k >= 1 and u32(20) == 0               # group leaf / internal node K, file consistency flags
    assert u64(24) == 0 and u64(32) == UNDEF and u64(40) == len(raw) and u64(48) == UNDEF   # base, free-space info, EOF, driver info
    # root group symbol table entry (spec III.C): link name offset, object header address, cache type 1 + scratch = B-tree, heap addresses
    assert u64(56) == 0 and u32(72) == 1 and u32(76) == 0
    root_header, btree, heap = u64(64), u64(80), u64(88)
    assert all(a % 8 == 0 and 96 <= a < len(raw) for a in (root_header, btree, heap))
    # ---- local heap (spec III.D): signature, version 0, reserved, data segment size, free-list head, data segment address
    assert raw[heap:heap + 4] == b"HEAP" and raw[heap + 4:heap + 8] == b"\0\0\0\0"
    seg_size, free_head, seg = u64(heap + 8), u64(heap + 16), u64(heap + 24)
    assert seg % 8 == 0 and seg + seg_size <= len(raw) and raw[seg] == 0   # offset 0 of the segment holds the empty string
    assert free_head % 8 == 0 and free_head + 16 <= seg_size
    assert u64(seg + free_head) == 1 and u64(seg + free_head + 8) == seg_size - free_head  # one free block: next = 1 (none), size to the end
    # ---- group B-tree node (spec III.A.1): signature, node type 0, level 0, entries used, siblings, key / child / key
    assert raw[btree:btree + 4] == b"TREE" and (u8(btree + 4), u8(btree + 5), u16(btree + 6)) == (0, 0, 1)
    assert u64(btree + 8) == UNDEF and u64(btree + 16) == UNDEF
    key0, snod, key1 = u64(btree + 24), u64(btree + 32), u64(btree + 40)
    assert key0 == 0 and btree + 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8 <= len(raw)   # the node has room for 2K children, 2K+1 keys
    # ---- symbol table node (spec III.B): signature, version 1, reserved, number of symbols, 40-byte entries
    assert raw[snod:snod + 4] == b"SNOD" and (u8(snod + 4), u8(snod + 5)) == (1, 0)
    nsym = u16(snod + 6)
    assert nsym == 3 <= 2 * leaf_k

    def heap_string(off):
        end = raw.index(b"\0", seg + off)
        return raw[seg + off:end].decode()

    names, headers = [], []
    for i in range(nsym):
        e = snod + 8 + 40 * i
        names.append(heap_string(u64(e)))
        headers.append(u64(e + 8))
        assert u32(e + 16) == 0 and u32(e + 20) == 0 and raw[e + 24:e + 40] == bytes(16)   # cache type 0, reserved, empty scratch pad
    assert names == sorted(names) == ["ex", "psi", "x"] and heap_string(key1) == "x"           # strcmp order; key 1 = largest name of the child

    def messages(addr):
        """version-1 object header (spec IV.A.1.a): version, reserved, message count, reference count, header size, padding to 8"""
        assert (u8(addr), u8(addr + 1)) == (1, 0) and u32(addr + 4) == 1
        count, size = u16(addr + 2), u32(addr + 8)
        assert u32(addr + 12) == 0
        out, o = [], addr + 16
        for _ in range(count):
            mtype, msize, flags = u16(o), u16(o + 2), u8(o + 4)
            assert msize % 8 == 0 and raw[o + 5:o + 8] == b"\0\0\0"
            out.append((mtype, raw[o + 8:o + 8 + msize], flags))
            o += 8 + msize
        assert o == addr + 16 + size
        return out

    # root group: exactly one symbol table message (0x0011) naming the same B-tree and heap
    (mtype, body, _), = messages(root_header)
    assert mtype == 0x0011 and struct.unpack("<QQ", body) == (btree, heap)

    expect = {"psi": ((1, 2, 3), np.arange(1.0, 7.0)), "ex": ((1, 2), np.array([-0.5, 0.25])), "x": ((0,), np.zeros(0))}
    for name, addr in zip(names, headers):
        dims, data = expect[name]
        msgs = messages(addr)
        by_type = {}
        for mtype, body, flags in msgs:
            by_type.setdefault(mtype, []).append(body)
        # dataspace (0x0001, version 1): version, rank, flags, reserved (5 bytes), dimension sizes
        ds, = by_type[0x0001]
        assert ds[0] == 1 and ds[1] == len(dims) and ds[2] == 0 and ds[3:8] == bytes(5)
        assert struct.unpack_from(f"<{len(dims)}Q", ds, 8) == dims
        # datatype (0x0003): class 1 (floating point) version 1; bit field 0x20 = little endian, mantissa normalisation 2 (implied msb);
        # byte 1 = sign location 63; size 8; properties: bit offset 0, precision 64, exponent location 52 size 11, mantissa location 0
        # size 52, exponent bias 1023 -- the encoding of H5T_IEEE_F64LE
        dt, = by_type[0x0003]
        assert dt[0] == 0x11 and dt[1] == 0x20 and dt[2] == 63 and dt[3] == 0 and struct.unpack_from("<I", dt, 4)[0] == 8
        assert struct.unpack_from("<HHBBBBI", dt, 8) == (0, 64, 52, 11, 0, 52, 1023)
        # layout (0x0008, version 3): class 1 = contiguous, data address, size in bytes
        lay, = by_type[0x0008]
        assert lay[0] == 3 and lay[1] == 1
        data_addr, nbytes = struct.unpack_from("<QQ", lay, 2)
        assert nbytes == 8 * data.size
        if data.size:
            assert data_addr % 8 == 0 and np.array_equal(np.frombuffer(raw, "<f8", data.size, data_addr), data)
        else:
            assert data_addr == UNDEF
        # attributes (0x000C, version 1): version, reserved, name size (with NUL), datatype size, dataspace size, then the three parts,
        # each padded to a multiple of 8 bytes
        attrs = {}
        for a in by_type.get(0x000C, []):
            assert a[0] == 1 and a[1] == 0
            nsz, tsz, ssz = struct.unpack_from("<HHH", a, 2)
            pad = lambda v: (v + 7) // 8 * 8
            o = 8
            aname = a[o:o + nsz].rstrip(b"\0").decode()
            assert a[o + nsz - 1] == 0
            o += pad(nsz)
            atype = a[o:o + tsz]
            o += pad(tsz)
            aspace = a[o:o + ssz]
            o += pad(ssz)
            rank = aspace[1]
            adims = struct.unpack_from(f"<{rank}Q", aspace, 8) if rank else ()
            n = int(np.prod(adims)) if rank else 1
            if atype[0] == 0x11:
                attrs[aname] = np.frombuffer(a, "<f8", n, o)
            else:
                # class 0 (fixed point) version 1, little endian, signed (bit 3), size 4, bit offset 0, precision 32: H5T_STD_I32LE
                assert atype[0] == 0x10 and atype[1] == 0x08 and struct.unpack_from("<I", atype, 4)[0] == 4 and struct.unpack_from("<HH", atype, 8) == (0, 32)
                attrs[aname] = np.frombuffer(a, "<i4", n, o)
        if name in ("psi", "ex"):
            assert np.array_equal(attrs["steps"], [0.1, 0.2, 0.3]) and np.array_equal(attrs["origin"], [-0.05, 0.0, 0.15])
        if name == "ex":
            assert attrs["plane"][0] == 1 and attrs["plane_coordinate"][0] == 2.5
        assert set(by_type) <= {0x0001, 0x0003, 0x0005, 0x0008, 0x000C, 0x0000}
