"""Ray sharding across the GPUs of one box (one process per GPU, torch.distributed).

Every ray is independent given the feature maps and cameras (no cross-ray op anywhere in
reference models.py:206-621), so the flattened scene-major ray list ``[0, b*R)`` is cut
into one contiguous range per rank; each rank renders only its range
(``car_render_args.ray_begin/ray_end``) and the rendered tiles are gathered at the end.
The reference has no ray-sharded inference (its eval scripts make every rank render the
same data, experiment_scripts/eval_realestate10k.py:95-105); this is the B200 design of
SURVEY.md §8(e).

Collectives (NCCL on GPUs, gloo in the CPU tests):
  * ``broadcast_features``  encoder maps from the rank that encoded them, once per scene batch;
  * ``gather_tiles``        all_gather of each rank's rgb / valid_mask / depth_ray range.
"""
import torch
import torch.distributed as dist


def ray_range(total, rank, world):
    """Contiguous, balanced range of rank ``rank``: sizes differ by at most one ray."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_ranges(total, world):
    return [ray_range(total, r, world) for r in range(world)]


def scenes_touched(begin, end, R):
    """Scene indices a flattened ray range touches (for feature residency planning)."""
    if end <= begin:
        return range(0)
    return range(begin // R, (end - 1) // R + 1)


def broadcast_features(z, src=0, group=None):
    """Broadcast the feature-map list from ``src`` (in place on the other ranks)."""
    for t in z:
        dist.broadcast(t, src=src, group=group)
    return z


def gather_tiles(out, total, rank, world, keys=("rgb", "valid_mask", "depth_ray"), group=None):
    """All-gather the per-rank ray ranges of full-size output tensors.

    ``out[k]`` is a full-size tensor of which only this rank's range is valid (what
    ``CrossAttentionRenderer.forward(..., ray_range=...)`` returns).  Ranges are padded to
    the largest range so a single all_gather per key suffices."""
    ranges = all_ranges(total, world)
    width = max(e - b for b, e in ranges)
    res = {}
    for k in keys:
        t = out[k]
        flat = t.reshape(total, -1)
        b, e = ranges[rank]
        send = flat.new_zeros(width, flat.shape[1])
        send[: e - b] = flat[b:e]
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send, group=group)
        full = torch.cat([recv[r][: ranges[r][1] - ranges[r][0]] for r in range(world)], dim=0)
        res[k] = full.reshape(t.shape)
    return res


def render_sharded(model, input, z, rank=None, world=None, gather=True, group=None):
    """Render this rank's share of the rays and (optionally) gather the tiles on all ranks."""
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    b = input["context"]["rgb"].shape[0]
    R = input["query"]["uv"].shape[2]
    total = b * R
    rng = ray_range(total, rank, world)
    out = model(input, z=z, ray_range=rng)
    if gather and world > 1:
        out.update(gather_tiles(out, total, rank, world, group=group))
    out["ray_range"] = rng
    return out


def feature_shapes(b, n_view, H, W):
    """Shapes of the encoder's maps for b scenes of n_view context images (reference models.py:178-188)."""
    return [(b * n_view, 256, H // 4, W // 4), (b * n_view, 256, H // 2, W // 2), (b * n_view, 64, H, W)]


def render_scenes_pipelined(model, scenes, src=0, group=None, gather=True, device=None):
    """Strong scaling over FEW scenes: every scene's rays are split across all ranks, and its feature maps -
    which exist only on rank ``src``, where the encoder ran - are broadcast once (north_star: "encoder features
    broadcast once via NCCL over NVLink"), asynchronously and one scene ahead: the broadcast of scene k+1
    overlaps the rendering of scene k.

    ``scenes``: list of ``(input, z)``; ``z`` is the feature list on rank ``src`` and ``None`` elsewhere.
    Returns (list of out dicts with gathered tiles, bytes broadcast per rank)."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if not scenes:
        return [], 0

    def start(k):
        inp, z = scenes[k]
        if z is None:
            b, n = inp["context"]["rgb"].shape[:2]
            H, W = inp["context"]["rgb"].shape[2:4]
            ref = inp["query"]["uv"]
            dev = device if device is not None else ref.device
            z = [torch.empty(shp, dtype=torch.float32, device=dev) for shp in feature_shapes(b, n, H, W)]
        works = [dist.broadcast(t, src=src, group=group, async_op=True) for t in z] if world > 1 else []
        return z, works

    outs, nbytes = [], 0
    nxt = start(0)
    for k in range(len(scenes)):
        z, works = nxt
        if k + 1 < len(scenes):
            nxt = start(k + 1)                      # in flight while scene k renders
        for w in works:
            w.wait()                                # the compute stream waits for scene k's maps (no host block on NCCL)
        nbytes += sum(t.numel() * t.element_size() for t in z) if world > 1 else 0
        outs.append(render_sharded(model, scenes[k][0], z, rank=rank, world=world, gather=gather, group=group))
    return outs, nbytes


def average_gradients(model, group=None, average=True):
    """Gradient exchange of the data-parallel training step (reference ``average_gradients``,
    training.py:21-28: one all_reduce per parameter, divided by the world size).  Here every
    gradient that exists is packed into ONE flat buffer and reduced with a single collective
    (NVSwitch reduces in-network; one launch instead of ~60), then scattered back.

    With ray-sharded rendering (each rank back-propagated a different ray range of the SAME
    batch) pass ``average=False``: the shard gradients add up to the full-batch gradient."""
    params = [p for p in model.parameters() if p.grad is not None]
    if not params:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= float(dist.get_world_size(group))
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()
