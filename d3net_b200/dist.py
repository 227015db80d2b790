"""Scene-per-GPU sharding of the proposal path and the one collective it needs.

Scenes are independent units of work: every op is confined to one scene, so a
batch is cut into contiguous blocks of scenes, each rank runs the whole chain on its block with no
exchange step, and only the per-scene proposal tensors that the captioning / grounding heads consume
(model/pointgroup.py:223-263, convert_stack_to_batch) are all-gathered -- as ONE packed fp32 buffer of
fixed shape [scenes_per_rank, P, 46], a few MB, i.e. a single latency-bound NCCL all-gather over
NVLink/NVSwitch.  No reduction is involved, so there is nothing to fuse a kernel with.
"""
import torch
import torch.distributed as dist

PACK_WIDTH = 46   # feats 16 | bbox corners 8x3 | center 3 | sem_cls | score | mask


def scene_shard(n_scenes, rank, world_size):
    """Contiguous block [lo, hi) of scene indices owned by `rank` (uneven remainders go to low ranks)."""
    base, rem = divmod(n_scenes, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_proposals(out, batch, max_num_proposal=256):
    """convert_stack_to_batch's padding (without its randperm): per scene the first `max_num_proposal`
    proposals -> [B, P, 46] fp32.  CUDA tensors go through the library's two pack kernels; the torch
    restatement below is the CPU path (gloo tests) and what the kernels are tested against."""
    if out["proposals_score_feats"].is_cuda:
        return _pack_proposals_cuda(out, batch, max_num_proposal)
    return pack_proposals_torch(out, batch, max_num_proposal)


def _pack_proposals_cuda(out, batch, P):
    import ctypes
    from . import _native
    B = int(batch["n_scenes"])
    feats = out["proposals_score_feats"].contiguous()
    dev = feats.device
    n_prop, C = feats.shape
    packed = torch.empty((B, P, PACK_WIDTH), dtype=torch.float32, device=dev)
    score = torch.sigmoid(feats[:, 0]).contiguous() if n_prop else feats.new_zeros(0)      # stand-in objectness score
    ws = torch.empty(max(2 * n_prop, 1), dtype=torch.int32, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        _native.check(_native.lib().pg_pack_proposals(
            p(out["proposals_idx"]), p(out["proposals_offset"]), p(batch["locs_scaled"]), p(batch["semantic_preds"]),
            p(out["proposals_center"].contiguous()), p(out["proposals_size"].contiguous()), p(feats), p(score), n_prop, C, B, P,
            p(ws), p(packed), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "pack_proposals")
    return packed


def pack_proposals_torch(out, batch, max_num_proposal=256):
    """Loop-free torch restatement of the padding."""
    B = int(batch["n_scenes"])
    P = max_num_proposal
    dev = out["proposals_score_feats"].device
    offs = out["proposals_offset"].long()
    n_prop = offs.numel() - 1
    packed = torch.zeros((B, P, PACK_WIDTH), dtype=torch.float32, device=dev)
    if n_prop == 0:
        return packed
    first_pt = out["proposals_idx"][offs[:-1], 1].long()
    scene = batch["locs_scaled"][first_pt, 0]                                  # proposals_batchId (:349)
    order = torch.argsort(scene, stable=True)
    scene_sorted = scene[order]
    start = torch.searchsorted(scene_sorted, torch.arange(B, device=dev))
    slot = torch.arange(n_prop, device=dev) - start[scene_sorted]
    keep = slot < P
    src = order[keep]
    b, s = scene_sorted[keep], slot[keep]
    center, size = out["proposals_center"][src], out["proposals_size"][src]
    signs = torch.tensor([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)],
                         dtype=torch.float32, device=dev)
    corners = center[:, None, :] + 0.5 * size[:, None, :] * signs[None]       # axis-aligned 8 corners
    feats = out["proposals_score_feats"][src]
    row = torch.cat([feats[:, :16], corners.reshape(-1, 24), center,
                     batch["semantic_preds"][first_pt[src]].float()[:, None],   # sem_cls (:359)
                     torch.sigmoid(feats[:, :1]),                               # stand-in objectness score
                     torch.ones((src.numel(), 1), device=dev)], 1)
    packed[b, s] = row
    return packed


def all_gather_proposals(packed, group=None, n_scenes_total=None):
    """[b, P, 46] per rank -> [sum of b, P, 46] on every rank with one all_gather_into_tensor.

    ``all_gather_into_tensor`` needs the same shape on every rank.  With ``n_scenes_total`` given, the ranks are
    taken to hold the blocks of ``scene_shard(n_scenes_total, rank, world)`` -- uneven when the batch does not divide
    -- and every rank pads its block to ceil(n / world) scenes before the collective; the padding is cut out of the
    result.  Without it the blocks must be equal (checked against the shard rule when it can be)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b = packed.size(0)
    if n_scenes_total is None:
        per = b
        counts = None
    else:
        per = -(-int(n_scenes_total) // world)
        counts = [hi - lo for lo, hi in (scene_shard(int(n_scenes_total), r, world) for r in range(world))]
        if counts[rank] != b:
            raise ValueError("rank %d holds %d scenes, scene_shard(%d, %d, %d) says %d"
                             % (rank, b, n_scenes_total, rank, world, counts[rank]))
    send = packed.contiguous()
    if b < per:
        send = torch.cat([send, send.new_zeros((per - b,) + tuple(send.shape[1:]))], 0)
    gathered = torch.empty((world * per,) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, send, group=group)
    if counts is not None and any(c != per for c in counts):
        gathered = torch.cat([gathered[r * per:r * per + c] for r, c in enumerate(counts)], 0)
    return gathered
