import ctypes, faulthandler, os, sys
import numpy as np
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = ctypes.CDLL(os.path.join(ROOT, "pypwt_b200", "libpwt_b200.so"))
lib.pwt_last_error.restype = ctypes.c_char_p
print("devices", lib.pwt_device_count(), flush=True)
img = np.random.default_rng(0).integers(0, 256, (256, 256)).astype(np.float32)
h = ctypes.c_void_p()
rc = lib.pwt_create(ctypes.byref(h), img.ctypes.data_as(ctypes.c_void_p), 256, 256, b"db2", 999, 1, 1, 0, 0, 2)
print("create rc", rc, lib.pwt_last_error(), flush=True)
print("forward", lib.pwt_forward(h), flush=True)
print("sync", lib.pwt_sync(h), lib.pwt_last_error(), flush=True)
buf = np.zeros((4, 4), np.float32)
print("get_coeff", lib.pwt_get_coeff(h, buf.ctypes.data_as(ctypes.c_void_p), 0), buf, flush=True)
print("inverse", lib.pwt_inverse(h), lib.pwt_sync(h), lib.pwt_last_error(), flush=True)
out = np.zeros_like(img)
print("get_image", lib.pwt_get_image(h, out.ctypes.data_as(ctypes.c_void_p)), np.abs(out - img).max(), flush=True)
lib.pwt_destroy(h)
print("destroyed", flush=True)
import pypwt_b200
print("pinned...", flush=True)
a = pypwt_b200.pinned_empty((4, 4))
print("pinned ok", a.shape, flush=True)
a[...] = 1
del a
print("pinned freed", flush=True)
W = pypwt_b200.Wavelets(img, "db2", 999)
print("W ok", W.levels, W.sizes, flush=True)
W.forward()
c = W.coeffs
print("coeffs ok", c[0], flush=True)
