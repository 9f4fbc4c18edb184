"""cuFFT as a COMPARATOR for the hand-written transforms (BASELINE north_star: 'with cuFFT timed only as a comparator').

Times, with CUDA events on the current stream, what a cuFFT-based transverse solve costs for the shapes of the BASELINE configs:
  real solve     rfft2 -> multiply by 1/symbol -> irfft2            (psi, psi_middle;   reference src/System_3d.cpp:686-706)
  complex solve  fft2  -> multiply by 1/symbol -> ifft2             (By + i Bz as one complex field)
through torch.fft (cuFFT plans cached by torch), next to the library's own solve measured the same way: qpic_profile_slices gives
the per-kernel times of the three kernels of a solve (rows, columns with the fused division, rows back).  Nothing here is on the
product path.

Usage (GPU box): python tools/cufft_compare.py [--n 512 1024 4096]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, wake_particle_lattice  # noqa: E402


def time_us(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[512, 1024])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    for n in a.n:
        src = torch.randn(n, n, dtype=torch.float64, device=dev)
        inv_h = torch.rand(n, n // 2 + 1, dtype=torch.float64, device=dev)
        csrc = torch.randn(n, n, dtype=torch.complex128, device=dev)
        inv_f = torch.rand(n, n, dtype=torch.float64, device=dev)
        t_real = time_us(lambda: torch.fft.irfft2(torch.fft.rfft2(src) * inv_h, s=(n, n)))
        t_cplx = time_us(lambda: torch.fft.ifft2(torch.fft.fft2(csrc) * inv_f))
        # the library's own kernels, eager with one CUDA event per launch (includes ~2-3 us of launch overhead per kernel, as the cuFFT
        # numbers above include torch's per-op launch overhead)
        nx = 40
        s = WakefieldSolver(nx, n, n, nx * 0.015625, 32.0 * n / 512, 32.0 * n / 512, 5)
        s.set_laser_gaussian(0.5, *gaussian_laser_widths(4.0, 8.0 * n / 512, 8.0 * n / 512), 0.4, 16.0 * n / 512, 16.0 * n / 512)
        s.init_wake_particles(*wake_particle_lattice(n, n, 32.0 * n / 512, 32.0 * n / 512, 1, 1))
        s.begin()
        s.step(8, wait=True)
        prof = s.profile_slices(8)
        per = {k: 1e3 * v[1] / v[0] for k, v in prof.items()}
        mine_real = per.get("k_fft_rows_r2c", 0) + per.get("k_fft_rows_c2r", 0)
        mine_cplx = per.get("k_fft_rows_b_forward", 0) + per.get("k_fft_rows_b_backward", 0)
        cols = per.get("k_fft_cols_solve", 0)  # mean over the half-spectrum and full-spectrum launches
        s.close()
        print(f"n = {n:5d}: cuFFT (torch.fft) real solve {t_real:8.1f} us, complex solve {t_cplx:8.1f} us | "
              f"qpic rows r2c + c2r {mine_real:6.1f} us, rows B fwd + bwd {mine_cplx:6.1f} us, columns (mean of both kinds) {cols:6.1f} us "
              f"=> real ~{mine_real + cols:6.1f} us, complex ~{mine_cplx + cols:6.1f} us (incl. fix-up / sources / stop test fused in)", flush=True)


if __name__ == "__main__":
    main()
