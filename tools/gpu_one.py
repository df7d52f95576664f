"""Scratch driver for ncu: forward + inverse at 8192^2 (or a given shape), one wavelet, L levels, kernel mode.
usage: python tools/gpu_one.py wname [levels] [mode] [rows cols [batch]]"""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
wn = sys.argv[1] if len(sys.argv) > 1 else "sym8"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
shape = (8192, 8192)
if len(sys.argv) > 5: shape = (int(sys.argv[4]), int(sys.argv[5]))
if len(sys.argv) > 6: shape = (int(sys.argv[6]),) + shape
img = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
W = pycudwt.Wavelets(img, wn, L)
W.set_kernel_mode(mode)
for _ in range(2):
    W.forward(); W.inverse()
W.timer_start()
for _ in range(10):
    W.forward(); W.inverse()
print(wn, shape, L, mode, "fwd+inv ms", W.timer_stop() / 10)
