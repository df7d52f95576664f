"""End-to-end pipelining experiment: K host threads / plans process alternate frames (H2D + forward + inverse + D2H)."""
import sys, time, threading; sys.path.insert(0, ".")
import numpy as np, pycudwt, pypwt_b200
N = 8192
img_src = np.random.default_rng(0).standard_normal((N, N)).astype(np.float32)
img = pypwt_b200.pinned_empty((N, N)); img[:] = img_src
for K in (1, 2, 3, 4):
    plans = [pycudwt.Wavelets(img, "db2", 3) for _ in range(K)]
    outs = [pypwt_b200.pinned_empty((N, N)) for _ in range(K)]
    for k in range(K):
        plans[k].forward(img); plans[k].inverse(); plans[k].image_into(outs[k])
    for steps in (12, 48):
        def worker(k):
            P, o = plans[k], outs[k]
            for _ in range(steps // K):
                P.forward(img); P.inverse(); P.image_into(o)
        th = [threading.Thread(target=worker, args=(k,)) for k in range(K)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        n = (steps // K) * K
        print("threads %d steps %d: %.3f ms/frame  %.1f Mpx/s  %.1f GB/s per direction" % (K, n, dt / n * 1e3, N * N * n / dt / 1e6, img.nbytes * n / dt / 1e9), flush=True)
    del plans, outs
