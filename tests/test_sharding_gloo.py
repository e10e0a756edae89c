"""Multi-GPU plumbing on CPU: env sharding by index and the rollout-statistics reduce over a 2-rank gloo group."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openballbot_rl_b200.training.utils import shard_envs


def test_shard_envs_partitions_exactly():
    for n in (1, 7, 10, 4096, 65536):
        for w in (1, 2, 4, 8):
            parts = [shard_envs(n, r, w) for r in range(w)]
            assert sum(c for c, _ in parts) == n
            offs = [o for _, o in parts]
            assert offs == sorted(offs) and offs[0] == 0
            for (c, o), (_, o2) in zip(parts[:-1], parts[1:]):
                assert o + c == o2


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openballbot_rl_b200.training.rollout import reduce_rollout_stats
    count, offset = shard_envs(10, rank, world)
    # synthetic finished-episode statistics of this rank's shard
    rng = np.random.default_rng(rank)
    ep_ret = torch.tensor(rng.uniform(1, 9, count), dtype=torch.float32)
    ep_len = torch.tensor(rng.integers(100, 500, count), dtype=torch.int32)
    done = torch.tensor(rng.integers(0, 2, count), dtype=torch.bool)
    stats = reduce_rollout_stats(ep_ret, ep_len, done, steps=count * 5)
    q.put((rank, offset, count, float(ep_ret[done].sum()), int(done.sum()), stats))
    dist.destroy_process_group()


def test_rollout_stats_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [0, 5] and [r[2] for r in res] == [5, 5]
    tot_ret, tot_n = sum(r[3] for r in res), sum(r[4] for r in res)
    for r in res:
        st = r[5]
        assert st["episodes"] == tot_n and st["env_steps"] == 50
        if tot_n:
            assert abs(st["ep_rew_mean"] - tot_ret / tot_n) < 1e-5
