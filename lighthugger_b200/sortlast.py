"""Host-side plan of the sort-last multi-GPU frame (the reference is single-GPU, so this has no
reference counterpart): which instances, key slots and shadow cascades a rank owns.  One process per GPU; the exchange
itself (uint64-min composite, reference-table all-gather, shadow-layer min) is done inside the C library over NCCL
(csrc/comm.cu).  Kept free of CUDA so that the plan is testable with the gloo backend on CPU."""
from . import _structs as S


def instance_range(num_instances, rank, world):
    """Contiguous instance range [first, first + count) of `rank`: the ranges tile [0, num_instances) exactly."""
    first, end = rank * num_instances // world, (rank + 1) * num_instances // world
    return first, end - first


def capacities(scene, world, split="instances"):
    """Per-rank list capacities.  Keys carry rank * (max_opaque + max_alpha) + local slot in a 25-bit field, so the
    per-rank share shrinks with the world size; every rank still culls against the full instance list capacity.
    With split == "cascades" a rank's cascade lists hold meshlets of ALL instances, so the opaque capacity is the whole
    scene's meshlet count where the slot field allows it (a cascade list that still overflows is reported as
    LH_ERR_CAPACITY at run time)."""
    if world == 1:
        return scene.default_capacities()
    per_rank = -(-scene.total_meshlets // world)
    alpha = max(1024, per_rank // 8)
    # twice the even share where the slot field allows it: stripes do not split the meshlets evenly
    opaque = max(per_rank + 1024, min(2 * per_rank + 1024, S.LH_VIS_MAX_SLOTS // world - alpha))
    if split == "cascades":
        opaque = max(opaque, min(scene.total_meshlets + 1024, S.LH_VIS_MAX_SLOTS // world - alpha))
    caps = S.LhCapacities(max_instances=scene.num_instances, max_opaque_draws=opaque, max_alpha_clip_draws=alpha)
    if (caps.max_opaque_draws + caps.max_alpha_clip_draws) * world > S.LH_VIS_MAX_SLOTS:
        raise ValueError("world * per-rank slots exceeds the 25-bit slot field of the visibility key")
    return caps


def slot_offset(caps, rank):
    """Global slot of the rank's local slot 0 (what lh_comm_init derives on the device side)."""
    return rank * (caps.max_opaque_draws + caps.max_alpha_clip_draws)


STRIPES_PER_RANK = 2   # interleaved screen stripes per GPU: wide stripes keep occluders together (measured on 2 x B200 at C2: 1 -> 299, 2 -> 321, 3 -> 318, 4 -> 310, 8 -> 294, 16 -> 276, 32 -> 266 frames/s)


def main_plan(num_instances, rank, world, partition="stripes"):
    """((first, count), (stripes, index, count)) of the main view of `rank`.
      "range":   contiguous instance ranges [r*N/G, (r+1)*N/G) -- every GPU gets a sparse copy of the whole scene;
      "stripes": all instances, filtered on the device by the screen stripe of their centre (lh_set_instance_stripes) --
                 neighbours share a GPU, so a GPU's own geometry keeps occluding itself."""
    if world == 1:
        return (0, num_instances), (0, 0, 1)
    if partition == "range":
        return instance_range(num_instances, rank, world), (0, 0, 1)
    if partition != "stripes":
        raise ValueError(f"unknown partition {partition!r}")
    return (0, num_instances), (STRIPES_PER_RANK * world, rank, world)


def shadow_plan(num_instances, rank, world, split="instances", partition="range"):
    """(cascade mask, (first instance, instance count), (stripes, index, count)) of the shadow work of `rank`.  Over the
    ranks the (cascade, instance) pairs are covered exactly once, so the uint32-min merge of the layers equals the
    single-GPU layers.
      "instances": every rank draws all four cascades of its share of the instances (even load whatever the cascade sizes);
      "cascades":  one cascade per GPU over ALL instances (BASELINE config 3: 4 GPUs); 2 GPUs take two cascades each, and
                   with more than 4 GPUs the ranks c, c + 4, ... split cascade c's instances between them.
    A share of the instances is an index range or a set of light-space stripes (see main_plan)."""
    if world == 1:
        return 0xF, (0, num_instances), (0, 0, 1)
    if split == "instances":
        mask, index, members = 0xF, rank, world
    elif split == "cascades":
        groups = min(world, 4)                   # ranks with equal rank % groups share their cascades
        mask = sum(1 << c for c in range(4) if c % groups == rank % groups)
        group = [r for r in range(world) if r % groups == rank % groups]
        index, members = group.index(rank), len(group)
    else:
        raise ValueError(f"unknown shadow split {split!r}")
    if members == 1:
        return mask, (0, num_instances), (0, 0, 1)
    if partition == "range":
        return mask, instance_range(num_instances, index, members), (0, 0, 1)
    if partition != "stripes":
        raise ValueError(f"unknown partition {partition!r}")
    return mask, (0, num_instances), (STRIPES_PER_RANK * members, index, members)


def apply(ctx, num_instances, rank, world, split="instances", partition="stripes"):
    """Configures a renderer.Context with the plan of `rank`."""
    (first, count), (stripes, index, members) = main_plan(num_instances, rank, world, partition)
    ctx.set_instance_range(first, count)
    ctx.set_instance_stripes(0, stripes, index, members)
    mask, (sfirst, scount), (sstripes, sindex, smembers) = shadow_plan(num_instances, rank, world, split, partition)
    ctx.set_shadow_cascade_mask(mask)
    ctx.set_shadow_instance_range(sfirst, scount)
    ctx.set_instance_stripes(1, sstripes, sindex, smembers)


class StripeBalancer:
    """Load balancing of the stripe partition from measured pass times (frame-to-frame coherence: the bounds found on one
    frame are used for the next).  One contiguous stripe per rank and view; every rank runs the same update on the same
    gathered costs, so all ranks hold identical bounds (which lh_set_instance_stripe_bounds requires).

    update(view, costs): costs[r] is what rank r measured for the view's passes with the current bounds.  The cost is
    taken as uniformly spread inside each stripe; the new bounds cut the cumulative cost into equal parts, and the move
    towards them is damped (the measured times also hold per-pass fixed costs that do not move with the bounds)."""

    def __init__(self, world, damping=0.6, min_width=0.005):
        self.world = world
        self.damping = damping
        self.min_width = min_width
        self.bounds = [[i / world for i in range(world + 1)] for _ in range(2)]

    def update(self, view, costs):
        b, n = self.bounds[view], self.world
        costs = [max(float(c), 1e-9) for c in costs]
        target = sum(costs) / n
        ideal = [0.0]
        r, acc = 0, 0.0
        for k in range(1, n):
            need = k * target
            while r < n - 1 and acc + costs[r] < need:
                acc += costs[r]
                r += 1
            ideal.append(b[r] + (b[r + 1] - b[r]) * min(1.0, max(0.0, (need - acc) / costs[r])))
        ideal.append(1.0)
        new = [bo + self.damping * (nw - bo) for bo, nw in zip(b, ideal)]
        new[0], new[-1] = 0.0, 1.0
        for k in range(1, n):                      # ascending, no stripe thinner than min_width
            new[k] = max(new[k], new[k - 1] + self.min_width)
        for k in range(n - 1, 0, -1):
            new[k] = min(new[k], new[k + 1] - self.min_width)
        self.bounds[view] = new
        return new

    def imbalance(self, costs):
        """max / mean of the measured costs (1.0 = perfectly even)"""
        return max(costs) / (sum(costs) / len(costs)) if costs else 1.0

    def apply(self, ctx, rank, views=(0, 1)):
        for v in views:
            ctx.set_instance_stripe_bounds(v, self.bounds[v], rank, self.world)


class RowBalancer:
    """Load balancing of the row slices the ranks composite and resolve (lh_comm_set_rows), from measured pass times: a
    slice of sky resolves in a fraction of the time of a slice of geometry, and the frame's closing barrier waits for the
    slowest rank.  Same rule as StripeBalancer: the cost is taken as uniform inside a slice, the new boundaries cut the
    cumulative cost into equal parts, the move is damped; every rank runs the same update on the same gathered costs."""

    def __init__(self, height, world, damping=0.6, min_rows=8):
        self.height, self.world, self.damping, self.min_rows = height, world, damping, min_rows
        self.bounds = [row_range(height, r, world)[0] for r in range(world)] + [height]

    def update(self, costs):
        b, n = [float(x) for x in self.bounds], self.world
        costs = [max(float(c), 1e-9) for c in costs]
        target = sum(costs) / n
        ideal = [0.0]
        r, acc = 0, 0.0
        for k in range(1, n):
            need = k * target
            while r < n - 1 and acc + costs[r] < need:
                acc += costs[r]
                r += 1
            ideal.append(b[r] + (b[r + 1] - b[r]) * min(1.0, max(0.0, (need - acc) / costs[r])))
        ideal.append(float(self.height))
        new = [int(round(bo + self.damping * (nw - bo))) for bo, nw in zip(b, ideal)]
        new[0], new[-1] = 0, self.height
        for k in range(1, n):
            new[k] = max(new[k], new[k - 1] + self.min_rows)
        for k in range(n - 1, 0, -1):
            new[k] = min(new[k], new[k + 1] - self.min_rows)
        self.bounds = new
        return new

    def apply(self, ctx, rank):
        ctx.comm_set_rows(self.bounds[rank], self.bounds[rank + 1])


def row_range(height, rank, world):
    """Screen rows [y0, y1) a rank composites and resolves (what lh_comm_init derives): even-sized slices."""
    per = ((height + world - 1) // world + 1) & ~1
    y0 = min(per * rank, height)
    return y0, min(y0 + per, height)


def to_ordered_int64(keys_u64):
    """uint64 keys -> int64 with the same order (flip the top bit): lets a backend without uint64 reductions (gloo)
    run the min-composite."""
    import numpy as np
    return (np.asarray(keys_u64, np.uint64) ^ np.uint64(1 << 63)).view(np.int64)


def from_ordered_int64(keys_i64):
    import numpy as np
    return np.asarray(keys_i64, np.int64).view(np.uint64) ^ np.uint64(1 << 63)
