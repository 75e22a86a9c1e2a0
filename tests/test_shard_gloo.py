"""CPU, world_size 2 over gloo: the multi-GPU host logic (file sharding, counter reduction,
max-over-ranks timing).  The per-rank work here is the oracle on the rank's files, so the test
also shows that the union of the shards equals the single-process result."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from rtl_433_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_files, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    from rtl_433_b200 import lib, synth
    devs = [d for d in lib.default_device_table() if d["protocol_num"] in (1, 2, 12, 19)]
    o = orc.Oracle(store_bitbuffers=False)
    o.add_devices(devs)
    mine = shard.files_for_rank(n_files, rank, world)
    samples = packages = events = 0
    per_file = {}
    for f in mine:
        x = synth.ook_stream(f, n_samples=1 << 17, n_bursts=1)
        r = o.run(x, 2)
        samples += len(x) // 2
        packages += len(r["packages"])
        events += len(r["events"])
        per_file[f] = (len(r["packages"]), len(r["events"]))
    tot = shard.reduce_report(samples, packages, events, 10.0 * (rank + 1))
    out[rank] = (mine, per_file, tot)
    dist.destroy_process_group()


def test_round_robin_shards_cover_every_file_once():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            mine = shard.files_for_rank(37, r, world)
            assert all(shard.owner_of(f, world) == r for f in mine)
            assert mine == sorted(mine)
            seen += mine
        assert sorted(seen) == list(range(37))


def test_two_ranks_gloo():
    world, n_files = 2, 6
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_files, out), nprocs=world, join=True)
    files0, per0, tot0 = out[0]
    files1, per1, tot1 = out[1]
    assert sorted(files0 + files1) == list(range(n_files))
    assert tot0 == tot1
    merged = {**per0, **per1}
    assert tot0[0] == n_files * (1 << 17)
    assert tot0[1] == sum(v[0] for v in merged.values()) > 0
    assert tot0[2] == sum(v[1] for v in merged.values()) > 0
    assert tot0[3] == 20.0  # max over ranks
