"""Host-side plan of the sort-last multi-GPU frame (the reference is single-GPU, so this has no
reference counterpart): which instances, key slots and shadow cascades a rank owns.  One process per GPU; the exchange
itself (uint64-min composite, reference-table all-gather, shadow-layer min) is done inside the C library over NCCL
(csrc/comm.cu).  Kept free of CUDA so that the plan is testable with the gloo backend on CPU."""
from . import _structs as S


def instance_range(num_instances, rank, world):
    """Contiguous instance range [first, first + count) of `rank`: the ranges tile [0, num_instances) exactly."""
    first, end = rank * num_instances // world, (rank + 1) * num_instances // world
    return first, end - first


def capacities(scene, world):
    """Per-rank list capacities.  Keys carry rank * (max_opaque + max_alpha) + local slot in a 25-bit field, so the
    per-rank share shrinks with the world size; every rank still culls against the full instance list capacity."""
    if world == 1:
        return scene.default_capacities()
    per_rank = -(-scene.total_meshlets // world)
    caps = S.LhCapacities(max_instances=scene.num_instances, max_opaque_draws=per_rank + 1024,
                          max_alpha_clip_draws=max(1024, per_rank // 8))
    if (caps.max_opaque_draws + caps.max_alpha_clip_draws) * world > S.LH_VIS_MAX_SLOTS:
        raise ValueError("world * per-rank slots exceeds the 25-bit slot field of the visibility key")
    return caps


def slot_offset(caps, rank):
    """Global slot of the rank's local slot 0 (what lh_comm_init derives on the device side)."""
    return rank * (caps.max_opaque_draws + caps.max_alpha_clip_draws)


def cascade_mask(rank, world):
    """Shadow cascades rendered by `rank`: one cascade per GPU at 4 GPUs, two each at 2; at 8 GPUs a cascade is shared by
    the two ranks r and r + 4, which split its instances (each rank rasterises its own instance range) and the layers
    are merged by the uint32-min exchange."""
    g = min(world, 4)
    return sum(1 << c for c in range(4) if c % g == rank % g)


def to_ordered_int64(keys_u64):
    """uint64 keys -> int64 with the same order (flip the top bit): lets a backend without uint64 reductions (gloo)
    run the min-composite."""
    import numpy as np
    return (np.asarray(keys_u64, np.uint64) ^ np.uint64(1 << 63)).view(np.int64)


def from_ordered_int64(keys_i64):
    import numpy as np
    return np.asarray(keys_i64, np.int64).view(np.uint64) ^ np.uint64(1 << 63)
