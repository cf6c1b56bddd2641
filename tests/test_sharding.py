"""CPU-only, world_size 2 over gloo: the multi-GPU path's host logic (contiguous shard
ranges, one all_gather of per-member compressed sizes, global concatenation offsets).
The per-shard "compress" is done by the oracle here; on the GPUs it is the CUDA path."""
import os
import socket

import numpy as np
import pytest

from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle as o
    from zippy_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = util.text_corpus(util.load_corpus())
    lo, hi = sharding.shard_range(n, rank, world)
    members = [o.compress(util.c2_block(T, i, 4096 + 37 * i), 1, o.dfGzip) for i in range(lo, hi)]
    sizes, offs = sharding.gather_sizes([len(m) for m in members], n)
    # the no-host-sync form (tensors stay where the collective put them) must agree
    t_sizes, t_offs = sharding.gather_sizes([len(m) for m in members], n, on_device=True)
    assert t_sizes.numpy().tolist() == list(sizes) and t_offs.numpy().tolist() == list(offs)
    # every rank writes its members at the global offsets into a shared file
    path = os.path.join(out_dir, "concat.bin")
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(int(offs[-1]))
    dist.barrier()
    with open(path, "r+b") as f:
        for k, m in enumerate(members):
            f.seek(int(offs[lo + k]))
            f.write(m)
    dist.barrier()
    np.save(os.path.join(out_dir, "offs_%d.npy" % rank), offs)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 10])
def test_two_rank_concat(tmp_path, n):
    import torch.multiprocessing as mp
    from oracle import oracle as o
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    T = util.text_corpus(util.load_corpus())
    blob = open(os.path.join(str(tmp_path), "concat.bin"), "rb").read()
    offs0 = np.load(os.path.join(str(tmp_path), "offs_0.npy"))
    offs1 = np.load(os.path.join(str(tmp_path), "offs_1.npy"))
    assert (offs0 == offs1).all() and len(offs0) == n + 1 and offs0[-1] == len(blob)
    for i in range(n):
        assert o.uncompress(blob[int(offs0[i]):int(offs0[i + 1])]) == util.c2_block(T, i, 4096 + 37 * i)


def test_shard_ranges():
    from zippy_b200 import sharding
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    lens = [1, 100, 1, 1, 50, 50, 1, 1]
    r = [sharding.shard_range_by_bytes(lens, k, 2) for k in range(2)]
    assert r[0][0] == 0 and r[1][1] == 8 and r[0][1] == r[1][0]
