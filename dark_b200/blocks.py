"""Independent-block sharding: the only multi-GPU axis of this path.

The reference encodes one block per file with one `Encoder` each (src/main.rs:95-113); a large
corpus is a list of independent blocks.  Blocks share nothing, so ranks (one process per GPU)
take disjoint subsets and never communicate on the data path — no NCCL, no P2P."""


def shard(num_blocks, rank, world_size):
    """Indices of the blocks rank `rank` of `world_size` processes: round-robin (blocks are
    equal-sized and near-equal cost, so static b mod G balances)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return list(range(rank, num_blocks, world_size))


def all_shards(num_blocks, world_size):
    return [shard(num_blocks, r, world_size) for r in range(world_size)]


def bwt_blocks(constructor, blocks):
    """Forward BWT of an iterable of host blocks on one context -> [(bwt, origin), ...], using the
    pipelined batch entry (copies overlap the transforms)."""
    return constructor.bwt_blocks(list(blocks))


def aggregate(ms_local, units_local, group=None):
    """(max over ranks of the device time, sum over ranks of the units processed) — the reduction
    behind the multi-GPU throughput.  Works on any torch.distributed backend (NCCL on the GPU box,
    gloo in the CPU tests); a single process returns its own numbers."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(ms_local), int(units_local)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=dev)
    u = torch.tensor([int(units_local)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), int(u.item())


def corpus_blocks(rank, world_size, per_rank):
    """Block numbers rank `rank` takes when every rank transforms `per_rank` blocks of the corpus: b = rank + i * world_size
    (the same round-robin as `shard` over per_rank * world_size blocks)."""
    return shard(per_rank * world_size, rank, world_size)


def gather_records(local_records, group=None):
    """All ranks' per-block records [(block, origin, crc32 hex), ...] merged and ordered by block number (on every rank).
    The only exchange of the multi-block path, and it is off the data path: a few bytes per block, after the timing."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(list(r) for r in local_records)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, [list(r) for r in local_records], group=group)
    return sorted(r for part in parts for r in part)


def corpus_digest(records):
    """CRC-32 over the per-block BWT CRCs in block order: equal for equal block sets however they were sharded."""
    import zlib
    return "%08x" % (zlib.crc32(",".join(r[2] for r in sorted(records)).encode()) & 0xFFFFFFFF)
