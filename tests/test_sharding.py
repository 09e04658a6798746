"""CPU: env sharding and the episode-stat reduction, world_size 2 over gloo."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nav_gym_b200.sharding import EpisodeStats, shard


def test_shard_partitions_exactly():
    for g, w in ((65536, 8), (4096, 1), (10, 4), (7, 8)):
        blocks = [shard(g, w, r) for r in range(w)]
        assert blocks[0][0] == 0
        assert sum(c for _, c in blocks) == g
        for (o0, c0), (o1, _) in zip(blocks, blocks[1:]):
            assert o0 + c0 == o1


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    off, cnt = shard(10, world, rank)
    st = EpisodeStats(cnt, 'cpu')
    g = torch.Generator().manual_seed(0)
    rew_all = torch.rand(6, 10, generator=g)
    done_all = torch.rand(6, 10, generator=g) < 0.3
    succ_all = torch.rand(6, 10, generator=g) < 0.5
    for t in range(6):
        sl = slice(off, off + cnt)
        st.update(rew_all[t, sl], done_all[t, sl], succ_all[t, sl], ~succ_all[t, sl])
    red = st.reduce()
    if rank == 0:
        # single-process reference over all 10 envs
        ref = EpisodeStats(10, 'cpu')
        for t in range(6):
            ref.update(rew_all[t], done_all[t], succ_all[t], ~succ_all[t])
        want = dict(zip(ref.FIELDS, ref.acc.tolist()))
        torch.save((red, want), out)
    dist.destroy_process_group()


def test_stats_reduce_world2(tmp_path):
    out = str(tmp_path / 'r.pt')
    mp.spawn(_worker, args=(2, 29653, out), nprocs=2, join=True)
    red, want = torch.load(out)
    for k in want:
        assert abs(red[k] - want[k]) < 1e-9, k
    assert red['episodes'] > 0 and red['steps'] == 60
