import sys; sys.path.insert(0, ".")
import importlib.util
spec = importlib.util.spec_from_file_location("g", "tools/gpu_row1d.py"); g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
tag = sys.argv[1] if len(sys.argv) > 1 else ""
cfgs = (((8192, 8192), 3), ((8192, 8192), 6), ((65536, 512), 3)) if "quick" not in sys.argv else (((8192, 8192), 3),)
for shape, L in cfgs:
    for wn in ("haar", "db2", "db4", "sym8", "db20"):
        f, fi = g.timeit(wn, shape, L)
        px = shape[0] * shape[1]
        print("%s 1D %-5s %5dx%-5d L%d fwd %.4f inv %.4f fwd+inv %.4f ms frac %.3f" % (tag, wn, shape[0], shape[1], L, f, fi - f, fi, 16 * px / fi / 1e6 / 6549.4), flush=True)
