"""Per-round node exchange between ranks (one process per GPU, torch.distributed).

The path shards by samples: rank r expands its contiguous shard of the round's samples against the full,
replicated tree with append deferred, then every rank all-gathers the fixed-stride records of the nodes it
accepted and appends ALL ranks' chunks in rank order.  Contiguous shards + rank order = global sample order, so
the tree is identical for any world size.  Works with NCCL (device tensors) and gloo (CPU tensors);
the only collectives on the path are the two all-gathers below.
"""
import torch
import torch.distributed as dist

RECORD_BYTES = 160


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of n_total samples for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(local, n_local, world, counts_buf=None, gather_buf=None, sync=True):
    """local: uint8 tensor holding at least n_local*RECORD_BYTES bytes (this rank's records, packed).
    Returns (gathered, counts, stride): `gathered` is a uint8 tensor of world*stride*RECORD_BYTES bytes in which rank
    r's records start at r*stride*RECORD_BYTES; counts[r] = number of records of rank r (numpy int32).
    A NCCL collective is only ordered with torch's CURRENT stream: with sync=True (default) the host waits for it, so that
    the gathered buffer may be consumed on any stream (clrrt_append_records runs on the planner's own stream unless the
    planner was created on torch's); sync=False is for callers whose planner shares torch's current stream (bench.py)."""
    dev = local.device
    mine = torch.tensor([n_local], dtype=torch.int32, device=dev)
    counts_t = counts_buf if counts_buf is not None else torch.zeros(world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(counts_t, mine)
    counts = counts_t.cpu().numpy().copy()
    stride = int(counts.max())
    if stride == 0:
        return None, counts, 0
    nbytes = stride * RECORD_BYTES
    if local.numel() < nbytes:  # pad: every rank must contribute `stride` records
        padded = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        padded[:local.numel()] = local
        local = padded
    out = gather_buf[:world * nbytes] if gather_buf is not None else torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, local[:nbytes].contiguous())
    if sync and out.is_cuda:
        torch.cuda.current_stream(dev).synchronize()
    return out, counts, stride


def concat_in_rank_order(gathered, counts, stride):
    """The appended order: rank 0's records, then rank 1's, ... (what clrrt_append_records does on the device)."""
    parts = [gathered[r * stride * RECORD_BYTES: (r * stride + int(counts[r])) * RECORD_BYTES] for r in range(len(counts))]
    return torch.cat(parts) if parts else gathered[:0]
