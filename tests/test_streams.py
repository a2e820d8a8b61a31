"""CPU: stream sharding and whole-job timing over a 2-rank gloo group (the N>1 path of bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tdnet_b200.streams import clips_for_rank, partition_is_exact, whole_job_throughput


def test_round_robin_partition():
    assert clips_for_rank(0, 8, 8) == [0] and clips_for_rank(7, 8, 8) == [7]
    assert clips_for_rank(1, 2, 5) == [1, 3]
    for world in (1, 2, 4, 8):
        for n in (1, 7, 8, 33):
            assert partition_is_exact(world, n)
    with pytest.raises(ValueError):
        clips_for_rank(2, 2, 4)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0: 40 frames in 1000 ms, rank 1: 40 frames in 1250 ms -> 80 frames / 1.25 s = 64 fps
    frames, ms, fps = whole_job_throughput(40, 1000.0 if rank == 0 else 1250.0)
    # each rank runs its own clip with its own FIFO: oracle on two different clips gives different outputs
    from oracle.tdnet_oracle import TDOracle
    from common import make_weights
    from tdnet_b200.synth import synth_clip
    sd = make_weights("td2_psp50", "resnet18", 4, 4)
    o = TDOracle("td2_psp50", sd, "resnet18")
    clip = clips_for_rank(rank, world, world)[0]
    y = None
    for i, f in enumerate(synth_clip(2, 32, 32, clip_id=clip)):
        y = o(f, pos_id=i % 2)
    sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(sums, y.double().sum().reshape(1))
    out.put((rank, frames, ms, fps, [float(s) for s in sums]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_weak_scaling_accounting():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, frames, ms, fps, sums in res:
        assert frames == 80 and ms == 1250.0 and abs(fps - 64.0) < 1e-9
        assert sums[0] != sums[1]          # independent clips, independent FIFOs
    assert res[0][4] == res[1][4]          # both ranks agree on the gathered values
