"""torchrun worker of tests/test_multigpu.py: renders one frame sort-last on WORLD_SIZE GPUs through the C ABI, under
both exchange modes (peer memory / whole-buffer NCCL) and both shadow splits, and checks on rank 0 that the composited
keys, the depth bounds + cascade matrices, the merged shadow layers and the lit colour are those of the same frame
rendered on one GPU (min is associative and commutative, so the composite must be bit-exact)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def decode(vis, tables, per_rank):
    """keys -> (depth bits, instance, meshlet, triangle) per pixel, -1 where nothing was drawn"""
    from helpers import vis_fields
    d, slot, tri = vis_fields(vis)
    out = np.full((vis.size, 3), -1, np.int64)
    cov = d != 0
    owner = slot[cov] // per_rank
    local = slot[cov] - owner * per_rank
    inst = np.zeros(owner.size, np.int64)
    meshlet = np.zeros(owner.size, np.int64)
    for r, t in enumerate(tables):
        m = owner == r
        inst[m], meshlet[m] = t[local[m], 0], t[local[m], 1]
    out[cov] = np.stack([inst, meshlet, tri[cov]], 1)
    return d, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="tiny")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from lighthugger_b200 import _structs as S, sortlast
    from lighthugger_b200.scene import Scene
    from helpers import Textures, make_context

    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sc = Scene.preset(args.scene)
    tex = Textures(sc, 32) if sc.num_texture_sets else None
    sky = sc.make_skybox(8)
    n = sc.num_instances
    W, H = sc.width, sc.height

    W2, H2 = (W * 3 // 4) & ~1, (H * 5 // 8) & ~1     # the size the resize case switches to
    want = want_resized = None
    if rank == 0:
        def single_gpu(width, height):
            one = make_context(sc, textures=tex, skybox=sky, device=local_rank, width=width, height=height)
            one.render_frame_host(one.uniforms())
            st = one.stats()
            table = one.meshlet_references(0, one.caps.max_opaque_draws + st.num_alpha_clip)
            d, triples = decode(one.target(S.LH_TARGET_VIS64), [table], one.caps.max_opaque_draws + one.caps.max_alpha_clip_draws)
            w = dict(depth=d, triples=triples, misc=bytes(one.misc()), shadow=one.target(S.LH_TARGET_SHADOW).view(np.uint32).copy(),
                     color=one.target(S.LH_TARGET_SCENE_RGBA16F).view(np.uint16).copy())
            one.close()
            return w
        want = single_gpu(W, H)
        want_resized = single_gpu(W2, H2)

    failures = []

    def check(ctx, tag, want, W, H, exchange, caps, per_rank, rows=None):
        """gathers every rank's part of the frame on rank 0 and compares it with the single-GPU frame `want`"""
        y0, y1 = ctx.comm_rows()
        assert (y0, y1) == (rows or (sortlast.row_range(H, rank, world) if exchange == S.LH_EXCHANGE_PEER else (0, H)))
        st = ctx.stats()
        mine = dict(rows=(y0, y1), vis=ctx.target(S.LH_TARGET_VIS64)[y0 * W:y1 * W].copy(),
                    table=ctx.meshlet_references(0, caps.max_opaque_draws + st.num_alpha_clip), misc=bytes(ctx.misc()),
                    shadow=ctx.target(S.LH_TARGET_SHADOW).view(np.uint32).copy(),
                    color=ctx.target(S.LH_TARGET_SCENE_RGBA16F).view(np.uint16).copy())
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank != 0:
            return
        tables = [p["table"] for p in parts]
        if exchange == S.LH_EXCHANGE_PEER:
            vis = np.concatenate([p["vis"] for p in parts])
        else:
            vis = parts[0]["vis"]
            if any(not np.array_equal(p["vis"], vis) for p in parts):
                failures.append(f"{tag}: ranks hold different composites")
        d, triples = decode(vis, tables, per_rank)
        if not np.array_equal(d, want["depth"]):
            failures.append(f"{tag}: depth bits differ on {(d != want['depth']).sum()} pixels")
        if not np.array_equal(triples, want["triples"]):
            failures.append(f"{tag}: (instance, meshlet, triangle) differ on {(triples != want['triples']).any(1).sum()} pixels")
        for r, p in enumerate(parts):
            if p["misc"] != want["misc"]:
                failures.append(f"{tag}: MiscStorage of rank {r} differs")
            if not np.array_equal(p["shadow"], want["shadow"]):
                failures.append(f"{tag}: shadow layers of rank {r} differ on {(p['shadow'] != want['shadow']).sum()} texels")
        # rank 0 holds the whole lit frame in both modes
        if not np.array_equal(parts[0]["color"], want["color"]):
            bad = (parts[0]["color"] != want["color"]).reshape(-1, 4).any(1)
            failures.append(f"{tag}: lit colour differs on {bad.sum()} pixels (rows {np.unique(np.nonzero(bad)[0] // W)[:8]})")
        print(f"[mgpu] {tag}: checked {vis.size} pixels, {int((d != 0).sum())} covered", flush=True)

    for exchange, ename, split, partition in ((S.LH_EXCHANGE_PEER, "peer", "instances", "stripes"),
                                              (S.LH_EXCHANGE_PEER, "peer", "cascades", "stripes"),
                                              (S.LH_EXCHANGE_PEER, "peer", "instances", "range"),
                                              (S.LH_EXCHANGE_NCCL, "nccl", "cascades", "range"),
                                              (S.LH_EXCHANGE_NCCL, "nccl", "instances", "stripes"),
                                              # stripes of unequal width (what the load balancer produces), then a collective
                                              # resize of every rank's targets and a frame at the new size
                                              (S.LH_EXCHANGE_PEER, "peer", "instances", "bounds")):
        caps = sortlast.capacities(sc, world, split)
        per_rank = caps.max_opaque_draws + caps.max_alpha_clip_draws
        ctx = make_context(sc, textures=tex, skybox=sky, caps=caps, device=local_rank)
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(rank, world, ids[0])
        ctx.comm_set_exchange(exchange)
        if partition == "bounds":
            sortlast.apply(ctx, n, rank, world, split, "stripes")
            bal = sortlast.StripeBalancer(world)
            bal.update(0, [1.0 + 3.0 * r for r in range(world)])      # pretend the right-hand ranks were slower
            bal.update(1, [4.0 - 3.0 * r / max(1, world - 1) for r in range(world)])
            bal.apply(ctx, rank)
            # ... and row slices of unequal height (lh_comm_set_rows; the balancer's rule on made-up costs)
            rows = sortlast.RowBalancer(H, world, min_rows=2)
            rows.update([1.0 + 2.0 * r for r in range(world)])
            rows.apply(ctx, rank)
            my_rows = (rows.bounds[rank], rows.bounds[rank + 1])
        else:
            my_rows = None
            sortlast.apply(ctx, n, rank, world, split, partition)
        u = ctx.uniforms()
        for _ in range(2):   # the second frame reuses the peer mappings and every self-cleaning counter
            ctx.render_frame_host(u)
        check(ctx, f"{ename}/{split}/{partition}", want, W, H, exchange, caps, per_rank, rows=my_rows)
        if partition == "bounds":
            ctx.resize(W2, H2)          # collective: unmap -> barrier -> reallocate; the row slices follow the new height
            u = ctx.uniforms()
            for _ in range(2):
                ctx.render_frame_host(u)
            check(ctx, f"{ename}/{split}/{partition} after resize to {W2}x{H2}", want_resized, W2, H2, exchange, caps, per_rank)
        ctx.close()
        dist.barrier()
    ok = [not failures]
    dist.broadcast_object_list(ok, src=0)
    if rank == 0:
        for f in failures:
            print("[mgpu] FAIL", f, flush=True)
        if not failures:
            print(f"MGPU OK scene={args.scene} world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok[0] else 1)


if __name__ == "__main__":
    main()
