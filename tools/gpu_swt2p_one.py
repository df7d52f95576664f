import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
img = np.random.default_rng(0).standard_normal((4096, 4096)).astype(np.float32)
W = pycudwt.Wavelets(img, "db10", 3, do_swt=1)
for _ in range(2): W.forward(); W.inverse()
W.sync()
