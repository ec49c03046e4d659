"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  SURVEY.md 8(e):

* whole estimator  -> replica per GPU over independent frames (bench.py), no data-path collective;
* refiner          -> `pose_shard`: the pose batch splits with no cross-pose term, one all_gather of
                      the [qn,7] results;
* selector         -> `shard_selector`: references split across ranks.  EXACT, not the per-shard
                      approximation: every InstanceNorm statistic that spans references is all-reduced
                      as fp64 (sum, sum of squares) (<= 2x512 doubles per layer; latency-bound),
                      vp_norm gathers the 3*S scores, and the per-reference score features [rfn,512]
                      are all-gathered once before the replicated attention tail.
"""
import os

import torch
import torch.distributed as dist


class Comm:
    """Thin wrapper over a torch.distributed process group with the three calls the sharded
    networks need.  gloo groups stage CUDA tensors through the host (tests); NCCL works in place."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if dist.is_initialized() else 'none'
        # NCCL collectives are stream-ordered kernels: they can be captured into the stage's CUDA graph and
        # issued from branch streams (every rank issues them in the same host order); gloo stages through
        # the host and must run eagerly.  G6D_SHARD_GRAPHS=0 forces the eager path.
        self.capturable = self.backend == 'nccl' and os.environ.get('G6D_SHARD_GRAPHS', '1') != '0'
        self.calls = {'all_reduce': 0, 'all_gather': 0}

    def _stage(self, t):
        return t.cpu() if (self.backend == 'gloo' and t.is_cuda) else t

    def all_reduce_sum(self, t):
        if self.world == 1:
            return t
        self.calls['all_reduce'] += 1
        s = self._stage(t).contiguous()
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.group)
        return s.to(t.device)

    def all_gather_cat(self, t, dim=0):
        """Concatenate equally-shaped shards in rank order along `dim`."""
        if self.world == 1:
            return t
        self.calls['all_gather'] += 1
        s = self._stage(t).contiguous()
        if self.backend == 'nccl':          # one collective into one buffer (rank-major), no per-rank temporaries
            out = torch.empty((self.world,) + tuple(s.shape), device=s.device, dtype=s.dtype)
            dist.all_gather_into_tensor(out, s, group=self.group)
            if dim == 0:
                return out.reshape((self.world * s.shape[0],) + tuple(s.shape[1:]))
            return out.movedim(0, dim).reshape(tuple(s.shape[:dim]) + (self.world * s.shape[dim],) + tuple(s.shape[dim + 1:]))
        parts = [torch.empty_like(s) for _ in range(self.world)]
        dist.all_gather(parts, s, group=self.group)
        return torch.cat(parts, dim).to(t.device)

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)

    def shard_range(self, n):
        return shard_range(n, self.rank, self.world)


def shard_range(n, rank, world):
    """Contiguous [begin, end) of `n` units for `rank`; requires an even split so that gathers are
    rank-major == unit-major (the selector's slice order r*an + a depends on it)."""
    if n % world != 0:
        raise ValueError(f'{n} units do not split evenly over {world} ranks')
    per = n // world
    return rank * per, (rank + 1) * per


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); returns Comm."""
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get('LOCAL_RANK', 0))
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return Comm()


def merge_moments(sum_x, sum_x2, count, eps=1e-5):
    """(scale, shift) of an InstanceNorm from globally summed moments: the arithmetic
    g6d_instnorm_finalize performs after the all-reduce (kept here for the CPU tests)."""
    mean = sum_x / count
    var = (sum_x2 / count - mean * mean).clamp_min(0)
    rstd = 1.0 / torch.sqrt(var + eps)
    return rstd, -mean * rstd


def shard_selector(selector, comm):
    """Make `selector` hold only its rank's slice of the reference views.  Call before
    load_ref_imgs / extract_ref_feats (every rank passes the FULL reference set)."""
    selector.comm = comm
    return selector


def pose_shard(forward_fn, tensors, comm):
    """Refiner batch-of-poses over GPUs: `tensors` are batched on dim 0 over qn poses; every rank
    runs `forward_fn` on its contiguous slice and the [qn_local, 7] results are all-gathered."""
    qn = tensors[0].shape[0]
    b, e = comm.shard_range(qn)
    out = forward_fn(*[t[b:e].contiguous() for t in tensors])
    return comm.all_gather_cat(out, dim=0)
