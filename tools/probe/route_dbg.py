import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from frender_b200.engine import Context, pack_keys, FrbError
import frender_oracle as O
from test_gpu_route import make_pair
ctx = Context(0, table_log2=16)
rng = random.Random(1)
t1, t2, table = make_pair(rng, 2000)
roles = O.sink_names(table)
names = sorted({n for n in roles.values() if n}); sid = {n: i for i, n in enumerate(names)}
role_of = {"index_hop": "#hop", "ambiguous": "#amb", "undetermined": "#und"}
keys = list(table)
routes = np.array([sid[roles[table[k][1]] if table[k][0] == "demuxable" else roles[role_of[table[k][0]]]] for k in keys], np.uint32)
ctx.route_load(pack_keys(keys), routes, len(names))
b1, b2 = t1.encode(), t2.encode()
def rec_end(b, k):   # offset after k records
    p = -1
    for _ in range(4 * k): p = b.index(b"\n", p + 1)
    return p + 1
for label, cuts in [("aligned cut", (rec_end(b1, 700), rec_end(b2, 700))), ("R2 cut mid seq", (rec_end(b1, 700), rec_end(b2, 700) + 80)),
                    ("R2 cut mid header", (rec_end(b1, 700), rec_end(b2, 700) + 10)), ("both mid", (rec_end(b1, 650) + 33, rec_end(b2, 700) + 100))]:
    ctx.route_reset()
    try:
        ctx.route_push(b1[:cuts[0]], b2[:cuts[1]], 0)
        r = ctx.route_pop(); print(label, "chunk0 pairs", r[4], "carry", r[5], r[6])
        ctx.route_push(b1[cuts[0]:], b2[cuts[1]:], 3)
        r = ctx.route_pop(); print(label, "chunk1 pairs", r[4], "carry", r[5], r[6])
    except FrbError as e:
        print(label, "ERROR", e)

def run(depth, cut1, cut2, seed):
    rng = random.Random(seed)
    ctx.route_reset()
    p1 = p2 = 0; q = []; k = 0
    try:
        while True:
            n1, n2 = cut1(rng), cut2(rng)
            d1, d2 = b1[p1:p1 + n1], b2[p2:p2 + n2]
            p1 += len(d1); p2 += len(d2)
            final = (1 if p1 >= len(b1) else 0) | (2 if p2 >= len(b2) else 0)
            ctx.route_push(d1, d2, final); q.append((k, len(d1), len(d2), final)); k += 1
            if len(q) == depth:
                r = ctx.route_pop(); print("  depth", depth, "chunk", q.pop(0), "pairs", r[4], "carry", r[5], r[6])
            if final == 3: break
        while q:
            r = ctx.route_pop(); print("  depth", depth, "chunk", q.pop(0), "pairs", r[4], "carry", r[5], r[6])
    except FrbError as e:
        print("  depth", depth, "ERROR at", q[0] if q else None, e)
for depth in (1, 2):
    print("random cuts, depth", depth)
    run(depth, lambda r: r.randint(100_000, 400_000), lambda r: r.randint(50_000, 500_000), 3)
