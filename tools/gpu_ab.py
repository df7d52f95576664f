"""Scratch A/B: time fwd / fwd+inv for the current env, and compare with kernel mode 3 (bit-identical when mode 3 runs the
register kernels, i.e. widths that are a multiple of 128 -- or Haar; (1024, 520) db2 goes through the fast kernels: False)."""
import sys, os, numpy as np
sys.path.insert(0, ".")
import pycudwt
wn = sys.argv[1] if len(sys.argv) > 1 else "db2"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
rng = np.random.default_rng(0)
for shape in ((512, 1024), (1024, 520), (264, 4096)):
    im = rng.standard_normal(shape).astype(np.float32)
    A = pycudwt.Wavelets(im, wn, 3); B = pycudwt.Wavelets(im, wn, 3); B.set_kernel_mode(3)
    A.forward(); B.forward()
    ca, cb = A.coeffs, B.coeffs
    ok = np.array_equal(ca[0], cb[0]) and all(np.array_equal(ca[i][j], cb[i][j]) for i in (1, 2, 3) for j in range(3))
    print("bit-identical", shape, ok, flush=True)
img = rng.standard_normal((N, N)).astype(np.float32)
W = pycudwt.Wavelets(img, wn, 3)
for what in ("fwd", "fwd+inv"):
    def step():
        W.forward()
        if what == "fwd+inv":
            W.inverse()
    for _ in range(10): step()
    best = 1e9
    for rep in range(3):
        W.timer_start()
        for _ in range(300): step()
        best = min(best, W.timer_stop() / 300)
    print(f"{wn} {N} {what}: {best:.4f} ms  variant={os.environ.get('PWT_FUSED_VARIANT','0')}", flush=True)
