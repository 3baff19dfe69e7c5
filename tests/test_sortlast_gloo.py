"""The N > 1 host logic on CPU: two processes over torch.distributed / gloo run the sort-last plan
(lighthugger_b200/sortlast.py) with the oracle standing in for the device passes, exchange their key buffers with a
MIN all-reduce and their shadow layers likewise, and must reproduce the single-process frame bit-exactly
(SURVEY.md 8e "determinism": min is associative and commutative)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import torch
    import torch.distributed as dist
    from lighthugger_b200 import _structs as S, sortlast
    from lighthugger_b200.scene import Scene
    from helpers import make_oracle, vis_fields

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = Scene.preset("tiny")
        first, count = sortlast.instance_range(sc.num_instances, rank, world)
        caps = sortlast.capacities(sc, world, "cascades")
        o = make_oracle(sc, caps=caps, instance_range=(first, count))
        o.reset(); o.cull_instances(); o.cull_meshlets(); o.raster_visbuffer()
        n_local = sum(o.draw_counts())
        # keys carry GLOBAL slots: rank * slots_per_rank + local (what lh_comm_init sets up on the device)
        vis = o.vis.copy()
        covered = vis != S.LH_VIS_CLEAR
        vis[covered] += np.uint64(sortlast.slot_offset(caps, rank) << 7)
        t = torch.from_numpy(sortlast.to_ordered_int64(vis).copy())
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        composite = sortlast.from_ordered_int64(t.numpy())
        # reference tables of every rank, rank-major, as the all-gather inside lh_composite lays them out
        per_rank = caps.max_opaque_draws + caps.max_alpha_clip_draws
        mine = torch.zeros((per_rank, 2), dtype=torch.int64)
        refs = o.meshlet_references(0, o.draw_counts()[0])
        mine[:len(refs)] = torch.from_numpy(np.asarray(refs, np.int64))
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        table = torch.cat(gathered).numpy()
        # shadows: depth bounds and cascade matrices from the COMPOSITED frame (every rank derives the same ones), then
        # each rank draws its share of the (cascade, instance) pairs and the layers are merged with a MIN
        shadows = {}
        o.vis[:] = composite
        o.reduce_depth()
        for split in ("instances", "cascades"):
            o.generate_shadow_matrices()   # also resets the list counters (generate_shadow_matrices.glsl ends in reset_counters)
            mask, (sfirst, scount), _ = sortlast.shadow_plan(sc.num_instances, rank, world, split, "range")
            o.cfg.instance_first, o.cfg.instance_count = sfirst, scount
            o.shadow[:] = 1.0
            o.cull_instances_shadows()
            for c in range(4):
                if mask >> c & 1:
                    o.shadow_cascade(c)
            layers = torch.from_numpy(o.shadow.view(np.int32).copy())   # depths in [0, 1]: bits order like values
            dist.all_reduce(layers, op=dist.ReduceOp.MIN)
            shadows[split] = layers.numpy().view(np.float32)
        if rank == 0:
            np.savez(os.path.join(out_dir, "composite.npz"), vis=composite, table=table, per_rank=per_rank, n_local=n_local,
                     misc=np.frombuffer(bytes(o.misc()), np.uint8), shadow_instances=shadows["instances"],
                     shadow_cascades=shadows["cascades"])
    finally:
        dist.destroy_process_group()


def test_two_rank_composite_equals_single_process_frame(tmp_path):
    import torch.multiprocessing as mp
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from lighthugger_b200 import _structs as S, sortlast
    from lighthugger_b200.scene import Scene
    from helpers import make_oracle, vis_fields

    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "composite.npz"))

    sc = Scene.preset("tiny")
    whole = make_oracle(sc)
    whole.reset(); whole.cull_instances(); whole.cull_meshlets(); whole.raster_visbuffer()
    dw, sw, tw = vis_fields(whole.vis)
    dg, sg, tg = vis_fields(got["vis"])
    assert np.array_equal(dw, dg), "composited depth differs from the single-process frame"
    cov = dw != 0
    ref_w = whole.meshlet_references(0, whole.draw_counts()[0])
    want = np.stack([ref_w[sw[cov], 0], ref_w[sw[cov], 1], tw[cov]], 1)
    table = got["table"]
    have = np.stack([table[sg[cov], 0], table[sg[cov], 1], tg[cov]], 1)
    assert np.array_equal(have, want), "(instance, meshlet, triangle) under the composited keys differ"
    # depth bounds + cascade matrices from the composite, and the merged shadow layers under both shadow splits
    whole.reduce_depth(); whole.generate_shadow_matrices(); whole.cull_instances_shadows()
    for c in range(4):
        whole.shadow_cascade(c)
    assert bytes(got["misc"]) == bytes(whole.misc())
    for split in ("instances", "cascades"):
        assert np.array_equal(got["shadow_" + split].view(np.uint32), whole.shadow.view(np.uint32)), split


def test_plan_tiles_instances_and_cascades():
    sys.path[:0] = [ROOT]
    from lighthugger_b200 import sortlast
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 1000, 10000):
            ranges = [sortlast.instance_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
            assert all(ranges[r][0] + ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
            # every (cascade, instance) pair is drawn by exactly one rank, under both shadow splits
            for split in ("instances", "cascades"):
                cover = np.zeros((4, n), np.int32)
                for r in range(world):
                    mask, (first, count), stripes = sortlast.shadow_plan(n, r, world, split, "range")
                    assert stripes[0] == 0
                    for c in range(4):
                        if mask >> c & 1:
                            cover[c, first:first + count] += 1
                assert np.all(cover == 1), (world, n, split)
        if world == 4:
            assert [sortlast.shadow_plan(100, r, 4, "cascades")[:2] for r in range(4)] == [(1 << r, (0, 100)) for r in range(4)]
        # stripe ownership: every rank looks at all instances; the (index, count) pairs tile the stripes' residues
        for split in ("instances", "cascades"):
            seen = {}
            for r in range(world):
                (first, count), (stripes, index, members) = sortlast.main_plan(1000, r, world, "stripes")
                assert (first, count) == (0, 1000) and (world == 1 or (stripes % members == 0 and index == r and members == world))
                mask, rng, (s2, i2, m2) = sortlast.shadow_plan(1000, r, world, split, "stripes")
                assert rng == (0, 1000)
                for c in range(4):
                    if mask >> c & 1:
                        seen.setdefault(c, []).append((i2, m2))
            for c in range(4):
                members = seen[c][0][1]
                assert sorted(i for i, _ in seen[c]) == list(range(members)) and all(m == members for _, m in seen[c])
        for h in (1, 2, 1080, 2160, 4320, 37):
            rows = [sortlast.row_range(h, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == h and all(rows[r][1] == rows[r + 1][0] for r in range(world - 1))


def test_stripe_balancer_converges_on_a_skewed_cost_field():
    """The balancer sees only per-rank costs; with a cost density that is 8x higher on the left of the view the bounds
    must move until every rank's cost is within a few percent of the mean, and stay ascending inside [0, 1]."""
    from lighthugger_b200.sortlast import StripeBalancer
    import math
    density = lambda t: 8.0 if t < 0.25 else 1.0          # cost per unit width
    def cost(a, b, fixed=0.05):
        n = 2000
        return fixed + sum(density(a + (b - a) * (i + 0.5) / n) for i in range(n)) * (b - a) / n
    for world in (2, 4, 8):
        bal = StripeBalancer(world)
        for _ in range(12):
            c = [cost(bal.bounds[0][r], bal.bounds[0][r + 1]) for r in range(world)]
            bal.update(0, c)
        b = bal.bounds[0]
        assert b[0] == 0.0 and b[-1] == 1.0 and all(b[k] < b[k + 1] for k in range(world))
        c = [cost(b[r], b[r + 1]) for r in range(world)]
        assert bal.imbalance(c) < 1.08, (world, c)


def test_row_balancer_moves_rows_away_from_the_expensive_slices():
    """Rows of sky cost a fraction of rows of geometry: the slices must keep tiling [0, height) with integer bounds and
    end up with nearly equal costs."""
    from lighthugger_b200.sortlast import RowBalancer, row_range
    height = 2160
    per_row = lambda y: 0.1 if y < 700 else 1.0            # the top third of the frame is sky
    cost = lambda a, b: 5.0 + sum(per_row(y) for y in range(a, b))
    for world in (2, 4, 8):
        bal = RowBalancer(height, world)
        assert bal.bounds == [row_range(height, r, world)[0] for r in range(world)] + [height]
        for _ in range(12):
            bal.update([cost(bal.bounds[r], bal.bounds[r + 1]) for r in range(world)])
        b = bal.bounds
        assert b[0] == 0 and b[-1] == height and all(isinstance(x, int) for x in b) and all(b[k] + 8 <= b[k + 1] for k in range(world))
        c = [cost(b[r], b[r + 1]) for r in range(world)]
        assert max(c) / (sum(c) / world) < 1.1, (world, b, c)
