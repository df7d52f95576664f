"""Multi-GPU front-end: a stack of independent images sharded over the GPUs of one box.

The wavelet path shards only across independent images: every rank (one process per
GPU) owns a contiguous block of slices and runs the ordinary kernels on it; nothing is exchanged on
the data path.  The single collective is the scalar all-reduce behind the GLOBAL `norm1` /
`norm2sq` (and anything derived from them):

* on GPUs the fused |c|, c^2 reduction kernel and `ncclAllReduce` are enqueued on the same stream by
  `pwt_norms_allreduce` (C ABI), the NCCL unique id being distributed through `torch.distributed`;
* without NCCL (CPU tests, `gloo`) the two local doubles are all-reduced by `torch.distributed`.

`engine` is the per-rank transform object; the default builds a `pycudwt.Wavelets` on the local
shard.  Tests inject another engine with the same interface to exercise the host logic on CPU.
"""
import numpy as np


def partition(n_slices, world_size):
    """Contiguous blocks of ceil(S/G) slices: [(start, stop)] per rank (empty ranks get (S, S))."""
    per = -(-n_slices // world_size)
    return [(min(r * per, n_slices), min((r + 1) * per, n_slices)) for r in range(world_size)]


def _default_engine(shard, wname, levels, **kw):
    import pycudwt
    return pycudwt.Wavelets(shard, wname, levels, **kw)


class ShardedWavelets:
    """`Wavelets` over a 3D stack, sharded along axis 0 across the ranks of `torch.distributed`.

    stack: the FULL stack (every rank passes the same array or a view; only the local block is used)
           or, with local_only=True, this rank's block only.
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
        self.W = (engine or _default_engine)(shard, wname, levels, **kw) if hi > lo else None
        self._nccl = False

    # -- communicator ----------------------------------------------------------------------------
    def init_nccl(self):
        """Create the NCCL communicator used by the fused norm all-reduce (GPU ranks only)."""
        if self.world_size == 1 or self.dist is None or self.W is None or not hasattr(self.W, "comm_init"):
            return False
        import pypwt_b200
        uid = [pypwt_b200.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(uid, src=0)
        self.W.comm_init(self.world_size, self.rank, uid[0])
        self._nccl = True
        return True

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
        """(norm1, norm2sq) over the whole stack, identical on every rank."""
        if self.world_size == 1:
            return self.local_norms()
        if self._nccl:
            return self.W.norms_allreduce()
        import torch
        n1, n2 = self.local_norms()
        t = torch.tensor([n1, n2], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0]), float(t[1])

    def norm1(self):
        return self.global_norms()[0]

    def norm2sq(self):
        return self.global_norms()[1]
