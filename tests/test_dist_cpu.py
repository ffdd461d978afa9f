"""world_size-2 gloo test of the N > 1 host logic (clip ownership, max-over-ranks timing, final latent gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asva_b200 import dist_utils, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clips = [f"clip{i:02d}" for i in range(7)]
    mine = dist_utils.shard_items(clips, rank, world)
    lat = synth.synth_inputs(F=3, h=4, w=4, seed=123 + rank, k=1)[0]  # one clip per rank, seeds 123, 124
    allv = dist_utils.gather_latents(lat)
    slow = dist_utils.max_over_ranks(10.0 + rank, "cpu")
    q.put((rank, mine, allv, slow))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_clip_sharding_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = res[0][1] + res[1][1]
    assert sorted(owned) == [f"clip{i:02d}" for i in range(7)] and len(set(owned)) == 7
    want = torch.cat([synth.synth_inputs(F=3, h=4, w=4, seed=123 + r, k=1)[0] for r in range(world)])
    for r in range(world):
        assert torch.equal(res[r][2], want)
        assert res[r][3] == 11.0  # the slowest rank defines the step time
