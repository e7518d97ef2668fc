"""Host-side sharding of the scan across ranks (one process per GPU).

Units (files, or record-aligned chunks of one file) are independent; a unit's key list carries
`first = (unit ordinal << 40) | read ordinal inside the unit`, so that the merged table -- counts
add, `first` takes the minimum -- is the same for every rank count and equals the reference's
first-appearance order over the whole input (F:199-205).  On the device the merge is
frb_allmerge (NCCL all-gather of the per-rank lists + a merge kernel); `merge_lists` is the same
monoid on the host, used by the CPU tests and by callers that gather lists themselves.
frb_shardmerge sends every key to one owner rank instead (`key_owner`, the host twin of the device
function of the same name); `shard_lists` / `fold_share` are that exchange on the host.
"""

_M64 = (1 << 64) - 1


def assign(units, rank, world):
    """Round-robin ownership: unit i belongs to rank i % world."""
    return [u for i, u in enumerate(units) if i % world == rank]


def merge_lists(lists):
    """[(key, count, first), ...] per unit -> {key: count} ordered by the smallest `first`."""
    count, first = {}, {}
    for lst in lists:
        for key, n, pos in lst:
            count[key] = count.get(key, 0) + n
            first[key] = min(first.get(key, pos), pos)
    return {k: count[k] for k in sorted(count, key=first.__getitem__)}


def _hash64(k):
    """murmur3 finaliser, as hash64() in csrc/common.cuh."""
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & _M64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & _M64
    return k ^ (k >> 33)


def key_owner(packed_key, world):
    """Owner rank of a packed key in the sharded merge (csrc/table_kernels.cuh key_owner)."""
    return (_hash64(int(packed_key)) >> 40) % world


def shard_lists(entries, world, pack):
    """[(key, count, first), ...] of one rank -> per-owner parts (what the rank sends to every peer).
    `pack` maps a key string to its packed 64-bit form."""
    parts = [[] for _ in range(world)]
    for key, n, pos in entries:
        parts[key_owner(pack(key), world)].append((key, n, pos))
    return parts


def fold_share(received):
    """What an owner does with the parts it receives: one entry per key (count: +, first: min), ordered by
    first appearance.  Returns [(key, count, first), ...]."""
    count, first = {}, {}
    for part in received:
        for key, n, pos in part:
            count[key] = count.get(key, 0) + n
            first[key] = min(first.get(key, pos), pos)
    return [(k, count[k], first[k]) for k in sorted(count, key=first.__getitem__)]
