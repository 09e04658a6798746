"""Environment sharding across the GPUs of one box.

Environments are independent, so the step path needs no collective: rank r of W owns the
contiguous block of global environment ids returned by :func:`shard`, and seeds its Philox
streams with the global id (``env_offset``) so results do not depend on W.  The only
communication is the optional episode-statistics reduction below (a handful of floats).
"""
import torch
import torch.distributed as dist


def shard(global_envs, world_size, rank):
    """(offset, count) of this rank's block; the first `global_envs % world_size` ranks get
    one extra environment."""
    base, extra = divmod(int(global_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


class EpisodeStats(object):
    """Running episode statistics of one shard: episodes, successes, crashes, truncations,
    sum of returns, sum of lengths, env-steps."""
    FIELDS = ('episodes', 'successes', 'crashes', 'truncated', 'return_sum', 'length_sum', 'steps')

    def __init__(self, num_envs, device):
        self.device = torch.device(device)
        self.acc = torch.zeros(len(self.FIELDS), dtype=torch.float64, device=self.device)
        self.ret = torch.zeros(num_envs, dtype=torch.float64, device=self.device)
        self.length = torch.zeros(num_envs, dtype=torch.float64, device=self.device)

    def update(self, reward, done, is_success, is_crash, truncated=None):
        d = done.to(torch.bool)
        self.ret += reward.to(torch.float64)
        self.length += 1
        z = torch.zeros((), dtype=torch.float64, device=self.device)
        tr = truncated.to(torch.float64).sum() if truncated is not None else z
        self.acc += torch.stack([d.sum().to(torch.float64), (is_success.to(torch.bool) & d).sum().to(torch.float64),
                                 (is_crash.to(torch.bool) & d).sum().to(torch.float64), tr,
                                 self.ret[d].sum(), self.length[d].sum(),
                                 torch.tensor(float(done.numel()), dtype=torch.float64, device=self.device)])
        self.ret[d] = 0
        self.length[d] = 0

    def reduce(self):
        """Sum over ranks (NCCL over NVLink on GPUs, gloo on CPU); a dict on every rank."""
        t = self.acc.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return dict(zip(self.FIELDS, t.tolist()))
