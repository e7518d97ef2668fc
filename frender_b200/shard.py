"""Host-side sharding of the scan across ranks (one process per GPU).

Units (files, or record-aligned chunks of one file) are independent; a unit's key list carries
`first = (unit ordinal << 40) | read ordinal inside the unit`, so that the merged table -- counts
add, `first` takes the minimum -- is the same for every rank count and equals the reference's
first-appearance order over the whole input (F:199-205).  On the device the merge is
frb_allmerge (NCCL all-gather of the per-rank lists + a merge kernel); `merge_lists` is the same
monoid on the host, used by the CPU tests and by callers that gather lists themselves.
"""


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
