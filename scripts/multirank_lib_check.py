"""Two (or more) ranks through the LIBRARY's own exchange (clrrt_comm_init + clrrt_expand_round, ncclAllGather inside
libclrrt_b200.so): no torch.distributed anywhere.  Each rank expands its contiguous shard of every round; the trees must
equal the single-GPU tree grown from all samples bit for bit (tree bytes and device digests).  Prints MULTIRANK_LIB_OK.

    python scripts/multirank_lib_check.py [world]      # spawns `world` processes, one per GPU
"""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CAR, GOAL = (0, 0, 0, 0, 3, 0), (50, 0, 0, 0)
K, ROUNDS = 4096, 4


def scene():
    from cpulib import scene_c1_boxes
    return scene_c1_boxes(moving=True)


def worker(rank, world, uid, q):
    import numpy as np
    import clrrt_b200 as clrrt
    pl = clrrt.Planner(device=rank, tree_capacity=1 << 16, max_round=K)
    pl.comm_init(uid, rank, world)
    pl.set_query(CAR, GOAL, 5.0)
    pl.set_obstacles(scene())
    pl.tree_reset(clrrt.root_node(CAR))
    s, h = clrrt.draw_samples(GOAL, K * ROUNDS, seed=7)
    sizes, ms_ex = [], []
    for r in range(ROUNDS):
        gs, gh = s[r * K:(r + 1) * K], h[r * K:(r + 1) * K]
        base, rem = divmod(K, world)
        lo = rank * base + min(rank, rem)
        hi = lo + base + (1 if rank < rem else 0)
        st = pl.expand_round(gs[lo:hi], gh[lo:hi])
        sizes.append((st.tree_size, st.nodes_added, st.nodes_local))
        ms_ex.append(st.ms_exchange)
    cg = pl.counters_global()
    q.put((rank, pl.tree_digest(), pl.tree_download().tobytes(), sizes, cg, ms_ex))
    pl.close()


def single(q):
    import clrrt_b200 as clrrt
    pl = clrrt.Planner(device=0, tree_capacity=1 << 16, max_round=K)
    pl.set_query(CAR, GOAL, 5.0)
    pl.set_obstacles(scene())
    pl.tree_reset(clrrt.root_node(CAR))
    s, h = clrrt.draw_samples(GOAL, K * ROUNDS, seed=7)
    for r in range(ROUNDS):
        pl.expand_round(s[r * K:(r + 1) * K], h[r * K:(r + 1) * K])
    q.put((-1, pl.tree_digest(), pl.tree_download().tobytes(), pl.tree_size(), pl.counters(), None))
    pl.close()


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    import clrrt_b200 as clrrt
    uid = clrrt.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, world, uid, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in ps:
        p.join()
    p1 = ctx.Process(target=single, args=(q,))
    p1.start()
    ref = q.get(timeout=600)
    p1.join()
    res.sort()
    for rank, dig, tree, sizes, cg, ms_ex in res:
        assert dig == ref[1], f"rank {rank}: digest {dig} != single-GPU digest {ref[1]}"
        assert tree == ref[2], f"rank {rank}: tree bytes differ from the single-GPU tree"
        assert sizes[-1][0] == ref[3]
        assert cg == ref[4], f"rank {rank}: counters summed over ranks {cg} != single-GPU counters {ref[4]}"
    print(f"world {world}: {ref[3]} nodes after {ROUNDS} rounds of {K} samples; digest {ref[1]} on every rank and on one GPU; "
          f"per-round (tree, added by all ranks, added by rank 0): {res[0][3]}; exchange ms per round on rank 0: "
          f"{[round(x, 3) for x in res[0][5]]}")
    print("MULTIRANK_LIB_OK")


if __name__ == "__main__":
    main()
