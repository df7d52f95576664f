"""Volumetric DWT timings (CUDA events): forward + inverse against the 16 B/voxel roofline."""
import sys; sys.path.insert(0, ".")
import numpy as np, pypwt_b200
for wn, shp, L in (("haar", (512, 512, 512), 3), ("db2", (512, 512, 512), 3), ("db3", (512, 512, 512), 3), ("db4", (512, 512, 512), 3), ("db5", (512, 512, 512), 3), ("sym8", (512, 512, 512), 3), ("db2", (256, 1024, 1024), 3), ("db2", (1024, 256, 256), 3)):
    vol = np.random.default_rng(0).standard_normal(shp).astype(np.float32)
    W = pypwt_b200.Wavelets3D(vol, wn, L)
    for _ in range(3): W.forward(); W.inverse()
    W.sync(); ts = []
    for rep in range(3):
        W.timer_start()
        for _ in range(10): W.forward(); W.inverse()
        ts.append(W.timer_stop() / 10)
    t = sorted(ts)[1]; n = vol.size
    l0 = W.launch_count; W.forward(); W.inverse(); nl = W.launch_count - l0
    print("3D %-5s %s L%d fwd+inv %.4f ms  %.1f Gvox/s  frac of the 16 B/voxel roofline %.3f  launches %d" % (wn, shp, W.levels, t, n / t / 1e6, 16 * n / t / 1e6 / 6549.4, nl), flush=True)
