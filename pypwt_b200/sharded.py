"""Multi-GPU front-ends: a stack of independent images sharded over the GPUs of one box.

The wavelet path shards only across independent images: every GPU owns a contiguous block of
slices and runs the ordinary kernels on it; nothing is exchanged on the data path.  The single collective is the
scalar all-reduce behind the GLOBAL `norm1` / `norm2sq` (and anything derived from them).  Two front-ends:

* `StackWavelets`  -- ONE process drives every GPU (`ncclCommInitAll`): `StackWavelets(stack3d, ...)`
  shards by itself, no launcher, no id exchange, no torch.  Transform calls are asynchronous, so one host thread keeps
  all GPUs busy; blocking calls (uploads, downloads) run on one worker thread per GPU.
* `ShardedWavelets` -- one process per GPU (torchrun / mpirun ...), the layout `bench.py --gpus N` uses.  The NCCL
  unique id travels through whatever collective layer the launcher provides (`dist`: anything with
  `broadcast_object_list` and `all_gather_object`, e.g. `torch.distributed`); this module itself never imports torch.

On GPUs the fused |c|, c^2 reduction kernel and `ncclAllReduce` are enqueued on the plan's stream by the C ABI
(`pwt_norms_allreduce`, `pwt_norms_allreduce_group`).  Without NCCL (CPU tests over `gloo`, injected engine) the two
local doubles are exchanged by `dist.all_gather_object`.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def partition(n_slices, world_size):
    """Contiguous blocks of ceil(S/G) slices: [(start, stop)] per rank (empty ranks get (S, S))."""
    per = -(-n_slices // world_size)
    return [(min(r * per, n_slices), min((r + 1) * per, n_slices)) for r in range(world_size)]


def _default_engine(shard, wname, levels, **kw):
    import pycudwt
    return pycudwt.Wavelets(shard, wname, levels, **kw)


class ShardedWavelets:
    """`Wavelets` over a 3D stack, sharded along axis 0 across the ranks of a one-process-per-GPU job.

    stack: the FULL stack (every rank passes the same array or a view; only the local block is used)
           or, with local_only=True, this rank's block only.
    Ranks whose block is empty (S < G, or ceil(S/G) * (G-1) >= S) still take part in every collective.
    """

    def __init__(self, stack, wname, levels, rank=0, world_size=1, dist=None, engine=None,
                 local_only=False, n_slices=None, **kw):
        self.rank, self.world_size, self.dist = rank, world_size, dist
        stack = np.asarray(stack)
        if stack.ndim != 3:
            raise ValueError("ShardedWavelets expects a 3D stack (slices, rows, cols)")
        total = n_slices if (local_only and n_slices is not None) else stack.shape[0]
        self.n_slices = total
        self.bounds = partition(total, world_size)
        lo, hi = self.bounds[rank]
        self.local_slices = (lo, hi)
        shard = stack if local_only else stack[lo:hi]
        if shard.shape[0] != hi - lo:
            raise ValueError("rank %d expects %d slices, got %d" % (rank, hi - lo, shard.shape[0]))
        self._engine = engine or _default_engine
        self.W = self._engine(shard, wname, levels, **kw) if hi > lo else None
        self._comm_plan = None     # the plan that owns this rank's NCCL communicator (self.W, or a stand-in when empty)
        self._nccl = False

    # -- communicator ----------------------------------------------------------------------------
    def init_nccl(self):
        """Create the NCCL communicator used by the fused norm all-reduce.  COLLECTIVE: every rank calls it, ranks
        with an empty block included (they join with a tiny all-zero plan, contributing 0 to the sums).  Rank 0
        decides whether NCCL is used (its engine must be the CUDA `Wavelets`) and says so in the broadcast."""
        if self.world_size == 1 or self.dist is None:
            return False
        uid = [None]
        if self.rank == 0:
            probe = self.W if self.W is not None else None
            if probe is None or hasattr(probe, "comm_init"):
                try:
                    import pypwt_b200
                    if pypwt_b200.device_count() > 0:
                        uid = [pypwt_b200.comm_unique_id()]
                except Exception:      # noqa: BLE001 -- no NCCL / no GPU: every rank falls back alike
                    uid = [None]
        self.dist.broadcast_object_list(uid, src=0)
        if uid[0] is None:
            return False
        plan = self.W
        if plan is None:
            import pycudwt
            plan = pycudwt.Wavelets(np.zeros((8, 8), np.float32), "haar", 1)     # zero coefficients: adds nothing
        plan.comm_init(self.world_size, self.rank, uid[0])
        self._comm_plan = plan
        self._nccl = True
        return True

    def close(self):
        if self._comm_plan is not None:
            self._comm_plan.comm_destroy()
            self._comm_plan = None
            self._nccl = False

    # -- local work: plain delegation ------------------------------------------------------------
    def forward(self, *a):
        if self.W is not None:
            self.W.forward(*a)

    def inverse(self):
        if self.W is not None:
            self.W.inverse()

    def soft_threshold(self, *a, **k):
        if self.W is not None:
            self.W.soft_threshold(*a, **k)

    def hard_threshold(self, *a, **k):
        if self.W is not None:
            self.W.hard_threshold(*a, **k)

    def shrink(self, *a, **k):
        if self.W is not None:
            self.W.shrink(*a, **k)

    @property
    def local_image(self):
        return None if self.W is None else self.W.image

    @property
    def local_coeffs(self):
        return None if self.W is None else self.W.coeffs

    # -- the collective ----------------------------------------------------------------------------
    def local_norms(self):
        if self.W is None:
            return 0.0, 0.0
        if hasattr(self.W, "norms"):
            return self.W.norms()
        return float(self.W.norm1()), float(self.W.norm2sq())

    def global_norms(self):
        """(norm1, norm2sq) over the whole stack, identical on every rank.  COLLECTIVE."""
        if self.world_size == 1:
            return self.local_norms()
        if self._nccl:
            return self._comm_plan.norms_allreduce()
        parts = [None] * self.world_size
        self.dist.all_gather_object(parts, tuple(float(x) for x in self.local_norms()))
        return float(sum(p[0] for p in parts)), float(sum(p[1] for p in parts))     # same order on every rank

    def norm1(self):
        return self.global_norms()[0]

    def norm2sq(self):
        return self.global_norms()[1]


class StackWavelets:
    """`Wavelets` over a 3D stack sharded across the GPUs of THIS process (default: all of them).

        S = StackWavelets(stack, "sym8", 3)          # slices split over the visible GPUs
        S.forward(); n1 = S.norm1(); S.soft_threshold(b); S.inverse(); out = S.image

    Same methods as `Wavelets` where they make sense for a stack; `coeffs` / `image` concatenate the shards along
    axis 0.  The global norms use one NCCL communicator created with `ncclCommInitAll` (skipped with one GPU).
    """

    def __init__(self, stack, wname, levels, devices=None, **kw):
        import pypwt_b200
        stack = np.asarray(stack)
        if stack.ndim != 3:
            raise ValueError("StackWavelets expects a 3D stack (slices, rows, cols)")
        ndev = pypwt_b200.device_count()
        if ndev < 1:
            raise RuntimeError("StackWavelets: no CUDA device available (there is no CPU fallback)")
        devices = list(range(ndev)) if devices is None else [int(d) for d in devices]
        if len(set(devices)) != len(devices) or not devices:
            raise ValueError("devices must be distinct")
        self.n_slices = stack.shape[0]
        bounds = [b for b in partition(self.n_slices, len(devices)) if b[1] > b[0]]
        self.devices = devices[:len(bounds)]
        self.bounds = bounds
        self._pool = ThreadPoolExecutor(max_workers=len(self.devices))

        def make(i):
            pypwt_b200.set_device(self.devices[i])     # cudaSetDevice is per host thread
            lo, hi = bounds[i]
            return pypwt_b200.Wavelets(stack[lo:hi], wname, levels, **kw)

        self.plans = list(self._pool.map(make, range(len(self.devices))))
        self.levels = self.plans[0].levels
        self._comm = False
        if len(self.plans) > 1:
            pypwt_b200.comm_init_all(self.plans)
            self._comm = True

    def _each(self, fn):
        """Run fn(plan, lo, hi) on every shard, blocking calls in parallel (the library releases the GIL)."""
        return list(self._pool.map(lambda a: fn(a[0], *a[1]), zip(self.plans, self.bounds)))

    # asynchronous calls: a single host thread queues them on every GPU
    def forward(self, stack=None):
        if stack is None:
            for W in self.plans:
                W.forward()
        else:
            stack = np.asarray(stack)
            self._each(lambda W, lo, hi: W.forward(stack[lo:hi]))

    def inverse(self):
        for W in self.plans:
            W.inverse()

    def soft_threshold(self, *a, **k):
        for W in self.plans:
            W.soft_threshold(*a, **k)

    def hard_threshold(self, *a, **k):
        for W in self.plans:
            W.hard_threshold(*a, **k)

    def shrink(self, *a, **k):
        for W in self.plans:
            W.shrink(*a, **k)

    def set_image(self, stack):
        stack = np.asarray(stack)
        self._each(lambda W, lo, hi: W.set_image(stack[lo:hi]))

    def sync(self):
        for W in self.plans:
            W.sync()

    def norms(self):
        """Global (norm1, norm2sq) of the whole stack: fused local reductions + grouped ncclAllReduce."""
        if not self._comm:
            return self.plans[0].norms()
        import pypwt_b200
        return pypwt_b200.norms_allreduce_group(self.plans)

    def norm1(self):
        return self.norms()[0]

    def norm2sq(self):
        return self.norms()[1]

    @property
    def image(self):
        return np.concatenate(self._each(lambda W, lo, hi: W.image), axis=0)

    def image_into(self, out):
        self._each(lambda W, lo, hi: W.image_into(out[lo:hi]))
        return out

    @property
    def coeffs(self):
        """[A, [H1, V1, D1], ...] with every band stacked over all slices."""
        per = self._each(lambda W, lo, hi: W.coeffs)
        out = [np.concatenate([c[0] for c in per], axis=0)]
        for l in range(1, self.levels + 1):
            out.append([np.concatenate([c[l][j] for c in per], axis=0) for j in range(3)])
        return out

    def close(self):
        for W in self.plans:
            W.comm_destroy()
        self._comm = False
        self._pool.shutdown(wait=True)

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass
