import sys, os; sys.path.insert(0, "."); sys.path.insert(0, "oracle/_ref")
import numpy as np, pycudwt
import pycudwt_ref
rng = np.random.default_rng(1)
for shape in ((96, 96), (33, 70), (32, 64)):
    img = (rng.standard_normal(shape) * 50 + 100).astype(np.float32)
    R = pycudwt_ref.Wavelets(img, "haar", 1); R.forward(); rc = [R.coeffs[0]] + list(R.coeffs[1]); R.inverse(); ri = np.array(R.image)
    for mode in (0, 3):
        W = pycudwt.Wavelets(img, "haar", 1); W.set_kernel_mode(mode); W.forward(); c = [W.coeffs[0]] + list(W.coeffs[1]); W.inverse(); im = W.image
        print(shape, "mode", mode, "fwd bit-identical to PDWT:", [bool(np.array_equal(a, b)) for a, b in zip(c, rc)], "inv:", bool(np.array_equal(im, ri)), "max diff", float(np.abs(im - ri).max()))
for shape in ((33, 70), (32, 64), (7, 9)):
    img = (rng.standard_normal(shape) * 50 + 100).astype(np.float32)
    R = pycudwt_ref.Wavelets(img, "haar", 2); R.forward(); rc = [R.coeffs[0]] + [b for l in R.coeffs[1:] for b in l]
    W = pycudwt.Wavelets(img, "haar", 2); W.set_kernel_mode(1); W.forward(); c = [W.coeffs[0]] + [b for l in W.coeffs[1:] for b in l]
    print(shape, "generic kernels (mode 1) fwd bit-identical to PDWT:", all(np.array_equal(a, b) for a, b in zip(c, rc)))
