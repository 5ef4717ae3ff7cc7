"""Run under torchrun (one rank per GPU): a round of 8192 samples sharded over the ranks with the NCCL record
all-gather must leave every rank with exactly the tree a single GPU builds from the whole round."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import clrrt_b200 as clrrt
from clrrt_b200.exchange import gather_records, shard_range
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = 8192
car, goal = (0, 0, 0, 0, 2, 0), (50, 0, 0, 0)
boxes = np.zeros((10, 7))
for i in range(10):
    boxes[i] = [8 + 4.7 * i, 3.0 if i % 2 == 0 else -3.0, 0.0, 4.0, 8.0, 0.0, 0.0]
pl = clrrt.Planner(device=local, tree_capacity=1 << 17, max_round=K)
pl.set_query(car, goal, 5.0); pl.set_obstacles(boxes); pl.tree_reset(clrrt.root_node(car))
ok = True
for rnd in range(3):
    s, h = clrrt.draw_samples(goal, K, seed=100 + rnd)     # every rank draws the same global round
    if rank == 0:                                           # single-GPU truth on a second context
        if rnd == 0:
            ref = clrrt.Planner(device=local, tree_capacity=1 << 17, max_round=K)
            ref.set_query(car, goal, 5.0); ref.set_obstacles(boxes); ref.tree_reset(clrrt.root_node(car))
        ref.expand_round(s, h)
    lo, hi = shard_range(K, rank, world)
    pl.set_defer_append(True)
    pl.expand_round(s[lo:hi], h[lo:hi])
    ptr, n = pl.round_records()
    src = bench._as_cuda_tensor(ptr, 2 * K * clrrt.RECORD_BYTES, local)
    out, counts, stride = gather_records(src, n, world)
    if stride:
        pl.append_records(out.data_ptr(), counts, stride)
    mine = pl.tree_download()
    # all ranks hold the same tree
    digest = torch.tensor([float(len(mine)), float(mine["state"].sum()), float(mine["parent"].sum())], dtype=torch.float64, device="cuda")
    allg = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(allg, digest)
    same = all(torch.equal(allg[0], a) for a in allg)
    if rank == 0:
        truth = ref.tree_download()
        eq = len(truth) == len(mine) and truth.tobytes() == mine.tobytes()
        print(f"round {rnd}: world {world} counts {counts.tolist()} tree {len(mine)} identical across ranks {same} equal to single-GPU tree {eq}")
        ok = ok and same and eq
if rank == 0:
    print("MULTIRANK_OK" if ok else "MULTIRANK_FAIL")
dist.destroy_process_group()
