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
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")], stdout=subprocess.DEVNULL)


def run_driver(name, workdir, golden_dir):
    build_driver()
    toml = os.path.join(golden_dir, "inputs", name + ".toml")
    subprocess.check_call(["cp", toml, workdir])
    return subprocess.run([DRIVER, name + ".toml"], cwd=workdir, capture_output=True, text=True)


def split_sections(text):
    return text.split("----------------------------------------\n")


@pytest.mark.parametrize("name", ["tiny_laser", "tiny_profile", "laser_test", "tiny_bunch", "tiny_bunch_stop", "tiny_beam"])
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
    r = subprocess.run([DRIVER, name + ".toml"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert not os.path.isdir(os.path.join(tmp_path, "Fields_000.h5.d"))
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
