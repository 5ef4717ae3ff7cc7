"""World-size-2 (and 3) gloo tests of the multi-GPU host logic: shard assignment and the per-round record
exchange (counts all-gather + padded record all-gather + rank-order append).  CPU only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clrrt_b200.exchange import RECORD_BYTES, concat_in_rank_order, gather_records
    results = []
    for rnd, cnt in enumerate(counts):
        n = cnt[rank]
        rng = np.random.default_rng(1000 * rnd + rank)
        local = torch.from_numpy(rng.integers(0, 256, max(n, 1) * RECORD_BYTES, dtype=np.uint8))
        gathered, c, stride = gather_records(local, n, world)
        assert c.tolist() == list(cnt)
        if stride == 0:
            results.append(np.zeros(0, np.uint8))
            continue
        results.append(concat_in_rank_order(gathered, c, stride).numpy().copy())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), *results)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_record_exchange_is_rank_ordered_and_identical_on_all_ranks(tmp_path, world):
    counts = [[5, 3, 4][:world], [0, 7, 1][:world], [0, 0, 0][:world], [129, 1, 64][:world]]
    mp.spawn(_worker, args=(world, _free_port(), counts, str(tmp_path)), nprocs=world, join=True)
    per_rank = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    for rnd, cnt in enumerate(counts):
        expect = np.concatenate([np.random.default_rng(1000 * rnd + r).integers(0, 256, max(cnt[r], 1) * 160, dtype=np.uint8)[:cnt[r] * 160]
                                 for r in range(world)])
        for r in range(world):
            got = per_rank[r][f"arr_{rnd}"]
            assert np.array_equal(got, expect), (rnd, r)


def test_shards_are_contiguous_and_cover_the_round():
    sys.path.insert(0, ROOT)
    from clrrt_b200.exchange import shard_range
    for n in (0, 1, 7, 4096, 65536, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
