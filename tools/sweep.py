#!/usr/bin/env python3
"""Measure the BASELINE.json configurations beyond the headline one (device-resident, CUDA events on
the plan's stream) for our kernels and for the reference's own CUDA build (oracle/_ref), and write a
markdown table.  Usage (on the GPU box): python tools/sweep.py gpurun_out/sweep.md [quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import pycudwt  # noqa: E402

try:
    import pycudwt_ref
except Exception:  # noqa: BLE001
    pycudwt_ref = None

PEAK = 6549.4e9


def synth(shape, seed=1):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(shape, dtype=np.float32) * 50 + 128)


def time_ours(W, fn, reps=20, warm=3):
    for _ in range(warm):
        fn(W)
    W.sync()
    W.timer_start()
    for _ in range(reps):
        fn(W)
    return W.timer_stop() / reps


def time_ref(R, fn, reps=5, warm=2):
    def sync():
        R.norm1()
    for _ in range(warm):
        fn(R)
    sync()
    t0 = time.perf_counter(); sync(); ts = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(R)
    sync()
    return ((time.perf_counter() - t0) - ts) / reps * 1e3


def fwd_inv(W):
    W.forward(); W.inverse()


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "sweep.md")
    quick = len(sys.argv) > 2
    rows = []

    def add(name, shape, wname, levels, bytes_per_px, fn=fwd_inv, ref=True, **kw):
        img = synth(shape)
        W = pycudwt.Wavelets(img, wname, levels, **kw)
        ms = time_ours(W, fn)
        pix = img.size
        launches0 = W.launch_count
        fn(W)
        nl = W.launch_count - launches0
        del W
        ms_ref = None
        if ref and pycudwt_ref is not None and img.ndim == 2:
            try:
                R = pycudwt_ref.Wavelets(img, wname, levels, **kw)
                ms_ref = time_ref(R, fn)
                del R
            except Exception as e:  # noqa: BLE001
                ms_ref = None
        gbs = bytes_per_px * pix / (ms * 1e-3) / 1e9
        rows.append((name, "x".join(map(str, shape)), wname, levels, ms, pix / ms / 1e3, gbs, gbs * 1e9 / PEAK, nl,
                     ms_ref, (ms_ref / ms) if ms_ref else None))
        print(rows[-1], flush=True)

    # C2: 4096^2 haar and db2, 3 levels
    for w in ("haar", "db2"):
        add("C2 fwd+inv", (4096, 4096), w, 3, 16)
    # headline size, haar
    add("M fwd+inv", (8192, 8192), "haar", 3, 16)
    add("M fwd+inv", (8192, 8192), "db2", 3, 16)
    # C3: per-GPU shard of the 512x2048x2048 sym8 stack (64 slices), per-slice 2D DWT + global norms
    nsl = 8 if quick else 64
    add("C3 stack fwd+inv", (nsl, 2048, 2048), "sym8", 3, 16, ref=False)
    add("C3 one slice (ref loops over slices)", (2048, 2048), "sym8", 3, 16)

    def fwd_norms(W):
        W.forward(); W.norm1()
    add("C3 stack fwd + norm1", (nsl, 2048, 2048), "sym8", 3, 12, fn=fwd_norms, ref=False)
    # C4: SWT db4 4 levels 8192^2 + cycle spinning + hard threshold loop

    def denoise(W):
        W.forward(); W.hard_threshold(20.0); W.inverse()
    side = 4096 if quick else 8192
    add("C4 swt fwd+hard+inv (cycle spinning)", (side, side), "db4", 4, 2 * (3 * 4 + 2) * 4, fn=denoise, do_swt=1, do_cycle_spinning=1)
    add("C4 swt fwd+inv", (side, side), "db4", 4, 2 * (3 * 4 + 2) * 4, do_swt=1)
    # C1-like denoising step on the headline size

    def denoise_dwt(W):
        W.forward(); W.soft_threshold(10.0); W.inverse()
    add("denoise fwd+soft+inv", (8192, 8192), "db2", 3, 16, fn=denoise_dwt)
    # C5: filter-length sweep, 5 levels, separable vs non-separable
    wl = ["haar", "db2", "db3", "db4", "db6", "db8", "db10", "db12", "db16", "db20", "coif5", "sym8"]
    if quick:
        wl = ["haar", "db2", "db4", "db8", "db20"]
    for w in wl:
        add("C5 sep fwd+inv", (8192, 8192), w, 5, 16)
    for w in (["haar", "db2", "db4"] if quick else ["haar", "db2", "db3", "db4", "db6", "db8"]):
        add("C5 nonsep fwd+inv", (4096, 4096), w, 5, 16, do_separable=0)
    # batched 1D (ndim=1: every row an independent signal), DWT and SWT
    for w in (["db2"] if quick else ["haar", "db2", "sym8"]):
        add("1D batched fwd+inv", (8192, 8192), w, 3, 16, ndim=1)
        add("1D batched swt fwd+inv", (8192, 8192), w, 3, 2 * (3 + 2) * 4, ndim=1, do_swt=1)
    with open(out, "w") as f:
        f.write("| config | shape | wavelet | L | ms | Mpixel/s | GB/s (algorithmic) | frac of 6549 GB/s | launches | PDWT CUDA ms | speed-up vs PDWT |\n")
        f.write("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            f.write("| %s | %s | %s | %d | %.4f | %.0f | %.0f | %.3f | %d | %s | %s |\n" % (
                r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8],
                "%.3f" % r[9] if r[9] else "-", "%.1fx" % r[10] if r[10] else "-"))
    print("wrote", out)


if __name__ == "__main__":
    main()
