"""Generates the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Only works in the build container (needs /root/reference and `make -C oracle ref`); the GPU box and the
test-suite only read the committed .npz files.  Usage:  python tests/golden/make_golden.py

The reference executable is oracle/_ref/quasistatic_pic_O2: /root/reference/src/*.cpp compiled unmodified with
-O2 (IEEE, no fast-math) against the shim headers in oracle/shims/, run with OMP_NUM_THREADS=1 so that its
atomic deposits land in index order and the output is reproducible to the last bit.

Fixtures:
  <name>.npz for the tiny inputs     every dataset of Fields_000.h5 (full 3-D) and Fields_xy_000.h5 + attributes
  tiny_beam.npz                      beam-driven run over three time iterations (two charged bunches + a photon beam): every
                                     dataset of Fields_NNN.h5 and Bunch_NNN.h5 per iteration   (python make_golden.py beam)
  tiny_qed.npz                       the same beams at gamma 4000 / 9000 with qed = true: Bunch_NNN and Photons_NNN    (... qed)
  laser_test_summary.npz             the shipped inputs/test/laser_test.toml (200^3): per-field min/max/L2 norm,
                                     per-slice L2 norms, on-axis lines [:, ny/2, nz/2], transverse lines through
                                     slice 100, and the value at [100,100,100]
  baseline512_6slices_summary.npz    BASELINE config #2's transverse problem (512 x 512, 2 x 2 ppc, M = 5, a0 = 1) on a 6-slice box:
                                     the same summary per field, transverse lines through slice 2, every 8th point of the
                                     xy cuts and of slice 1                                         (python make_golden.py baseline512)
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "quasistatic_pic_O2")
NAMES_3D = ["aSqr", "susceptibility", "rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz", "psi", "rho", "jx", "jy", "jz"]


def read_attrs(path):
    out = {}
    if not os.path.exists(path):
        return out
    for line in open(path):
        parts = line.split()
        name, dtype, count = parts[0], parts[1], int(parts[2])
        vals = parts[3:3 + count]
        out[name] = np.array([float(v) for v in vals]) if dtype == "<f8" else np.array([int(v) for v in vals], dtype=np.int32)
    return out


def run_reference(toml_path, workdir):
    name = os.path.basename(toml_path)
    shutil.copy(toml_path, os.path.join(workdir, name))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    log = subprocess.run([REF_EXE, name], cwd=workdir, env=env, check=True, capture_output=True, text=True).stdout
    return log


def collect(workdir, log):
    data = {}
    for kind, sub in (("f3", "Fields_000.h5.d"), ("xy", "Fields_xy_000.h5.d")):
        for nm in NAMES_3D:
            data[f"{kind}/{nm}"] = np.load(os.path.join(workdir, sub, nm + ".npy"))
            for an, av in read_attrs(os.path.join(workdir, sub, nm + ".attrs")).items():
                data[f"{kind}/{nm}@{an}"] = av
    data["stopped"] = np.array(int("Stop condition encountered" in log))
    data["stdout"] = np.array(log)
    return data


BUNCH_NAMES = ["x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi", "weight"]


def collect_time_loop(workdir, log, iterations):
    """multi-iteration run with beam output: Fields_NNN (3-D) and Bunch_NNN of every time iteration"""
    data = {"stdout": np.array(log), "iterations": np.array(iterations)}
    for t in range(iterations):
        for nm in NAMES_3D:
            data[f"it{t}/f3/{nm}"] = np.load(os.path.join(workdir, f"Fields_{t:03d}.h5.d", nm + ".npy"))
        for nm in BUNCH_NAMES:
            data[f"it{t}/bunch/{nm}"] = np.load(os.path.join(workdir, f"Bunch_{t:03d}.h5.d", nm + ".npy"))
        data[f"it{t}/photons"] = np.array(sorted(f for f in os.listdir(os.path.join(workdir, f"Photons_{t:03d}.h5.d")) if f.endswith(".npy")))
    return data


def collect_qed(workdir, log, iterations):
    """qed = true: Bunch_NNN and Photons_NNN of every time iteration (the reference run with one thread draws rand() in particle
    order from glibc's default seed, so the run is reproducible) + stdout with the "Max radiation probability" lines"""
    data = {"stdout": np.array(log), "iterations": np.array(iterations)}
    for t in range(iterations):
        for nm in BUNCH_NAMES:
            data[f"it{t}/bunch/{nm}"] = np.load(os.path.join(workdir, f"Bunch_{t:03d}.h5.d", nm + ".npy"))
            path = os.path.join(workdir, f"Photons_{t:03d}.h5.d", nm + ".npy")
            if os.path.exists(path):
                data[f"it{t}/photons/{nm}"] = np.load(path)
    return data


def main():
    if not os.path.exists(REF_EXE):
        sys.exit("build the reference first: make -C oracle ref")
    if len(sys.argv) > 1 and sys.argv[1] == "qed":
        with tempfile.TemporaryDirectory() as wd:
            log = run_reference(os.path.join(HERE, "inputs", "tiny_qed.toml"), wd)
            data = collect_qed(wd, log, 3)
        np.savez_compressed(os.path.join(HERE, "tiny_qed.npz"), **data)
        print("tiny_qed", data["it0/bunch/x"].shape, data["it0/photons/x"].shape, data["it1/photons/x"].shape)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "baseline512":
        with tempfile.TemporaryDirectory() as wd:
            log = run_reference(os.path.join(HERE, "inputs", "baseline512_6slices.toml"), wd)
            data = collect(wd, log)
        summary = {"stdout": data["stdout"], "stopped": data["stopped"]}
        for nm in NAMES_3D:
            a = data["f3/" + nm]
            nx, ny, nz = a.shape
            summary[nm + "/min"] = a.min()
            summary[nm + "/max"] = a.max()
            summary[nm + "/l2"] = np.linalg.norm(a)
            summary[nm + "/slice_l2"] = np.sqrt((a ** 2).sum(axis=(1, 2)))
            summary[nm + "/axis"] = a[:, ny // 2, nz // 2].copy()
            summary[nm + "/line_y"] = a[2, :, nz // 2].copy()
            summary[nm + "/line_z"] = a[2, ny // 2, :].copy()
            summary[nm + "/line_y_off"] = a[1, :, nz // 2 + 37].copy()   # off-axis lines: the on-axis ones of by / bz vanish by symmetry
            summary[nm + "/line_z_off"] = a[1, ny // 2 - 53, :].copy()
            summary[nm + "/xy"] = data["xy/" + nm][:, ::8].copy()
            summary[nm + "/plane"] = a[1, ::8, ::8].copy()
        np.savez_compressed(os.path.join(HERE, "baseline512_6slices_summary.npz"), **summary)
        print("baseline512_6slices summary written", data["f3/psi"].shape)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "beam":  # only the time-loop fixture (added after the others)
        with tempfile.TemporaryDirectory() as wd:
            log = run_reference(os.path.join(HERE, "inputs", "tiny_beam.toml"), wd)
            data = collect_time_loop(wd, log, 3)
        np.savez_compressed(os.path.join(HERE, "tiny_beam.npz"), **data)
        print("tiny_beam", data["it0/f3/psi"].shape, data["it0/bunch/x"].shape)
        return
    for name in ["tiny_laser", "tiny_profile", "tiny_bunch", "tiny_bunch_stop"]:
        with tempfile.TemporaryDirectory() as wd:
            log = run_reference(os.path.join(HERE, "inputs", name + ".toml"), wd)
            data = collect(wd, log)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, data["f3/psi"].shape, "stopped" if data["stopped"] else "")

    with tempfile.TemporaryDirectory() as wd:
        log = run_reference(os.path.join(HERE, "inputs", "laser_test.toml"), wd)
        data = collect(wd, log)
    summary = {"stdout": data["stdout"], "stopped": data["stopped"]}
    for nm in NAMES_3D:
        a = data["f3/" + nm]
        nx, ny, nz = a.shape
        summary[nm + "/min"] = a.min()
        summary[nm + "/max"] = a.max()
        summary[nm + "/l2"] = np.linalg.norm(a)
        summary[nm + "/slice_l2"] = np.sqrt((a ** 2).sum(axis=(1, 2)))
        summary[nm + "/axis"] = a[:, ny // 2, nz // 2].copy()
        summary[nm + "/line_y"] = a[100, :, nz // 2].copy()
        summary[nm + "/line_z"] = a[100, ny // 2, :].copy()
        summary[nm + "/probe"] = a[100, 100, 100]
        summary[nm + "/xy"] = data["xy/" + nm][::4, ::4].copy()
    np.savez_compressed(os.path.join(HERE, "laser_test_summary.npz"), **summary)
    print("laser_test summary written")


if __name__ == "__main__":
    main()
