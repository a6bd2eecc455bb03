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
    """Run the forward BWT over an iterable of host blocks on one context; yields (bwt, origin)."""
    for blk in blocks:
        yield constructor.bwt(blk)
