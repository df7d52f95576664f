#!/usr/bin/env python3
"""Generate tests/golden/pdwt_golden.npz from the REFERENCE ITSELF.

Runs the unmodified reference (PDWT + its Cython wrapper, built for sm_100a by `make -C oracle ref`
into oracle/_ref/) on a GPU and stores its outputs for small seeded inputs.  These vectors pin the
CPU oracle (tests/test_oracle.py) to the reference's actual CUDA arithmetic.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/pdwt_golden.npz'
    cp gpurun_out/pdwt_golden.npz tests/golden/

One reference instance is alive at a time (its filters are process-global __constant__ state,
SURVEY quirk Q4) and non-separable instances are not reused after inverse().
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))


def inputs():
    rng = np.random.default_rng(20261017)
    return {
        "even": rng.integers(0, 256, size=(32, 48)).astype(np.float32),
        "odd": rng.integers(0, 256, size=(31, 45)).astype(np.float32),
        "smooth": (rng.standard_normal((40, 40)) * 50 + 128).astype(np.float32),
    }


WAVELETS = ["haar", "db2", "db3", "db4", "db5", "sym4", "sym8", "coif1", "coif2",
            "bior1.3", "bior2.2", "bior3.1", "bior4.4", "rbio2.4", "rbio3.1", "db10"]

# (tag, ctor kwargs, levels)
MODES = [
    ("dwt2", dict(), 2),
    ("swt2", dict(do_swt=1), 2),
    ("dwt1", dict(ndim=1), 2),
    ("swt1", dict(do_swt=1, ndim=1), 2),
    ("ns_dwt2", dict(do_separable=0), 2),
    ("ns_swt2", dict(do_separable=0, do_swt=1), 2),
]


def flat_coeffs(c):
    out = []
    for b in c:
        if isinstance(b, list):
            out += [np.array(x) for x in b]
        else:
            out.append(np.array(b))
    return out


def main(path):
    import pycudwt_ref as ref
    store = {}
    for iname, img in inputs().items():
        store["in/" + iname] = img
        for wname in WAVELETS:
            for tag, kw, lev in MODES:
                if iname == "smooth" and (tag not in ("dwt2", "swt2", "dwt1") or wname not in ("haar", "db2", "sym4", "bior2.2")):
                    continue
                if iname == "odd" and tag in ("swt1", "ns_swt2") and wname not in ("haar", "db2", "db3", "bior2.2"):
                    continue
                key = "%s/%s/%s" % (iname, wname, tag)
                try:
                    W = ref.Wavelets(img, wname, lev, **kw)
                except IndexError:
                    # image too small for this filter: the reference clips the level count to 0
                    # and its wrapper trips over the empty size list (pypwt.pyx:193)
                    continue
                store[key + "/levels"] = np.array([W.levels], np.int32)
                W.forward()
                for i, b in enumerate(flat_coeffs(W.coeffs)):
                    store[key + "/c%d" % i] = b
                if iname == "smooth":
                    store[key + "/norm1"] = np.array([W.norm1()], np.float64)
                    if "1" not in tag:   # the 1D norm2sq of the reference is known-wrong (quirk Q3)
                        store[key + "/norm2sq"] = np.array([W.norm2sq()], np.float64)
                    W.soft_threshold(10.0, 1, 1)
                    for i, b in enumerate(flat_coeffs(W.coeffs)):
                        store[key + "/soft%d" % i] = b
                    W.forward()
                    W.hard_threshold(10.0, 1, 1)
                    for i, b in enumerate(flat_coeffs(W.coeffs)):
                        store[key + "/hard%d" % i] = b
                    W.forward()
                    W.shrink(0.5, 1)
                    store[key + "/shrinkA"] = np.array(W.coeff_only(0))
                    W.forward()
                W.inverse()
                store[key + "/inv"] = np.array(W.image)
                del W
    # cycle spinning: first two shifts of a fresh process + shifted image
    img = inputs()["even"]
    W = ref.Wavelets(img, "db2", 2, do_cycle_spinning=1)
    W.forward()
    store["cs/shifted_image"] = np.array(W.image)
    store["cs/A"] = np.array(W.coeff_only(0))
    W.inverse()
    store["cs/inv"] = np.array(W.image)
    del W
    np.savez_compressed(path, **store)
    print("wrote %s: %d arrays" % (path, len(store)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "pdwt_golden.npz"))
