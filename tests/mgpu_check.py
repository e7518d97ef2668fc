"""Run under torchrun with N ranks (one per GPU): the record chunks of ONE synthetic lane are dealt round-robin to
the ranks, every rank scans its chunks with their global line numbers, the per-rank tables are merged over NCCL
(frb_allmerge, then frb_shardmerge) and every rank must end with the oracle's tally of the whole lane, in the
oracle's order, and the oracle's classifications."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch.distributed as dist

    import frender_oracle as O
    from frender_b200 import _lib as L
    from frender_b200 import synth
    from frender_b200.engine import C, Context, unpack_keys
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = Context(int(os.environ["LOCAL_RANK"]), table_log2=18)
    spec = synth.make_spec("C2", n_samples=64)
    per = 30_000
    # CHUNK sharding of ONE file: rank r owns the chunks r, r + world, ... of the lane (three chunks each, so that a
    # rank's chunks are not contiguous); a chunk is scanned with its global line number, so `first` holds global
    # read ordinals and the file ordinal is the same (0) on every rank
    n_chunks = 3 * world
    per //= 3
    mine = [k for k in range(n_chunks) if k % world == rank]

    def scan_my_chunks():
        ctx._ck(L.lib.frb_scan_begin(ctx._h, 0, 0))
        for k in mine:
            piece = np.frombuffer(synth.generate(spec, k * per, (k + 1) * per), np.uint8)
            ctx._ck(L.lib.frb_scan_chunk_host(ctx._h, piece.ctypes.data_as(C.c_void_p), piece.size, 4 * k * per, L.RULE_SCAN))
        r, u = C.c_uint64(), C.c_uint64()
        ctx._ck(L.lib.frb_scan_end(ctx._h, C.byref(r), C.byref(u)))
        assert r.value == len(mine) * per
        ctx.file_names.append("lane")

    ident = (C.c_char * 128)()
    if rank == 0:
        ctx._ck(L.lib.frb_nccl_unique_id(ident))
    box = [bytes(ident)]
    dist.broadcast_object_list(box, src=0)
    ctx._ck(L.lib.frb_nccl_init(ctx._h, box[0], rank, world))
    ctx.reset()
    scan_my_chunks()
    n = C.c_uint64()
    ctx._ck(L.lib.frb_allmerge(ctx._h, C.byref(n)))
    keys, counts, _ = ctx.total_arrays()
    whole = synth.generate(spec, 0, n_chunks * per)
    want, visited = O.tally_text(whole.decode().splitlines(keepends=True))
    got = dict(zip(unpack_keys(keys), counts.tolist()))
    assert visited == n_chunks * per
    assert list(got.items()) == list(want.items()), f"rank {rank}: merged tally differs from the oracle"
    res, calls, _ = ctx.analyze(spec.indexes(), 1, True)
    want_res, want_calls, _ = O.scan_analysis(1, {"total": want}, spec.indexes(), 1, True)
    assert res == want_res and calls == want_calls, f"rank {rank}: matcher differs from the oracle"
    # sharded merge of the same per-rank tallies: disjoint shares, union == the oracle's tally (counts and
    # first-appearance order), every share classified like the oracle classifies those keys
    ctx.reset()
    scan_my_chunks()
    ctx.shardmerge()
    skeys, scounts, sfirst = ctx.total_arrays()
    sres, _, _ = ctx.analyze(spec.indexes(), 1, True)
    shares = [None] * world
    dist.all_gather_object(shares, (unpack_keys(skeys), scounts.tolist(), sfirst.tolist(), sres))
    union = [(f, k, n) for ks, ns, fs, _ in shares for k, n, f in zip(ks, ns, fs)]
    assert len({k for _, k, _ in union}) == len(union), "shares overlap"
    from frender_b200.shard import key_owner           # host twin of the device's owner function
    assert all(key_owner(int(pk), world) == rank for pk in skeys.tolist()), "a key sits on the wrong rank"
    assert sorted(fs for fs in shares[rank][2]) == shares[rank][2], "share not in first-appearance order"
    union.sort()
    assert [(k, n) for _, k, n in union] == list(want.items()), "union of the shares differs from the oracle"
    merged_res = {}
    for _, _, _, part in shares:
        merged_res.update(part["total"] if "total" in part else part)
    ref = want_res["total"] if "total" in want_res else want_res
    assert merged_res == ref or {k: merged_res[k] for k in ref} == ref, "sharded matcher differs from the oracle"
    dist.barrier()
    if rank == 0:
        print(f"mgpu ok: {world} ranks, {n_chunks} chunks of one file round-robin, {len(got)} unique keys, merged == oracle "
              f"on every rank (counts, first-appearance order, both matcher passes), shares {[len(s[0]) for s in shares]}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
