"""Image-space tile sharding of ONE render across the GPUs of a node (no counterpart in the reference,
which has no distributed code at all -- SURVEY.md section 2a / 8e).

One process per GPU (torchrun), parameters replicated.  Tile ``t`` is owned by rank ``t % world``
(interleaved for load balance against non-uniform triangle density).  Per frame:

  forward   every rank runs the per-triangle preprocess and the depth sort (cheap, keeps ``radii`` identical everywhere),
            emits / sorts / composites only its own tiles;
  backward  every rank walks its own tiles, producing partial per-triangle accumulators (64 B per triangle, themselves
            bit-reproducible); their sum over the ranks feeds the replicated per-triangle backward.

Two exchange paths, same results on every rank either way:
  fabric    (default when torch symmetric memory gives a multicast mapping) the outputs live in a symmetric buffer; exchange kernels
            launched behind the composite kernels (include/ts2d.h: ts2d_exchange_tiles, ts2d_exchange_allreduce) copy the owned
            tiles' rows into every replica with multimem.st and combine home slices of the contrib statistics / accumulators inside
            the NVSwitch (multimem.ld_reduce); two signal-pad rendezvous per pass, no collective library call;
  NCCL      (TS2D_FABRIC=0, or no multicast) all-reduce(sum) of the zero-filled frame planes + contrib_sum, all-reduce(max) of
            contrib_max, all-reduce(sum) of the accumulators, on the current CUDA stream through torch.distributed
            (backend "nccl" on GPUs, "gloo" in the CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

_STATE = {"enabled": False, "group": None, "rank": 0, "world": 1, "fabric": None, "fabric_mode": "auto"}


class Fabric:
    """Symmetric buffers over NVLink peer memory with NVSwitch multicast aliases (torch symmetric memory: cuMem + cuMulticast), the
    storage the exchange kernels of include/ts2d.h work on.  Buffers are persistent (allocation + rendezvous is a collective): one per
    role, grown on demand."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm

        self.symm, self.group, self.device = symm, (group if group is not None else dist.group.WORLD), device
        self.buffers = {}

    def buffer(self, role: str, numel: int):
        """-> (local fp32 tensor of `numel` elements, multicast base address, handle).  Collective when it has to (re)allocate:
        every rank asks for the same roles and sizes in the same order (parameters are replicated)."""
        ent = self.buffers.get(role)
        if ent is None or ent[0].numel() < numel:
            cap = max(int(numel * 1.25), 1 << 16) if ent is not None else max(int(numel), 1 << 16)
            t = self.symm.empty(cap, dtype=torch.float32, device=self.device)
            h = self.symm.rendezvous(t, self.group)
            if not int(h.multicast_ptr):
                raise RuntimeError("symmetric memory has no multicast mapping on this system")
            ent = (t, h)
            self.buffers[role] = ent
        t, h = ent
        return t[:numel], int(h.multicast_ptr), h


def fabric(device) -> Optional["Fabric"]:
    """The peer-memory fabric of the current sharding group, or None (then the NCCL collectives below carry the exchange).
    TS2D_FABRIC=0 disables it, TS2D_FABRIC=1 makes a failure to set it up an error instead of a silent NCCL path."""
    import os

    if not _STATE["enabled"] or _STATE["world"] == 1 or device.type != "cuda":
        return None
    mode = os.environ.get("TS2D_FABRIC", "auto")
    if mode == "0":
        return None
    from . import _lib

    if _STATE["world"] > _lib.MAX_RANKS:  # the exchange kernels address at most TS2D_MAX_RANKS replicas: larger groups take the NCCL path
        return None
    f = _STATE["fabric"]
    if f is None:
        try:
            f = Fabric(_STATE["group"], device)
            f.buffer("probe", 1 << 16)  # allocation + rendezvous + multicast mapping must all work (collective: same on all ranks)
        except Exception:  # noqa: BLE001
            if mode == "1":
                raise
            f = False
        _STATE["fabric"] = f
    return f or None


def enable_tile_sharding(group: Optional["dist.ProcessGroup"] = None) -> Tuple[int, int]:
    """Turn on tile sharding over ``group`` (default: the WORLD group). Returns (rank, world)."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _STATE.update(enabled=True, group=group, rank=dist.get_rank(group), world=dist.get_world_size(group), fabric=None)
    return _STATE["rank"], _STATE["world"]


def disable_tile_sharding() -> None:
    _STATE.update(enabled=False, group=None, rank=0, world=1, fabric=None)


def current_shard() -> Tuple[int, int]:
    if not _STATE["enabled"] or _STATE["world"] == 1:
        return (0, 1)
    return (_STATE["rank"], _STATE["world"])


def owned_tiles(n_tiles: int, rank: int, world: int) -> range:
    """Tile ids owned by ``rank`` (the rule the CUDA kernels use: tile % world == rank)."""
    return range(rank, n_tiles, world)


def _bucket_all_reduce(tensors, op) -> None:
    """One collective for several tensors: pack -> all_reduce -> unpack (launch-latency bound otherwise)."""
    tensors = [t for t in tensors if t is not None and t.numel() > 0]
    if not tensors:
        return
    if len(tensors) == 1:
        dist.all_reduce(tensors[0], op=op, group=_STATE["group"])
        return
    # tensors carved back to back out of one buffer (the sharded forward allocates its image planes that way): reduce in place
    adjacent = all(t.is_contiguous() and t.dtype == tensors[0].dtype for t in tensors) and all(
        b.untyped_storage().data_ptr() == a.untyped_storage().data_ptr() and b.data_ptr() == a.data_ptr() + a.numel() * a.element_size()
        for a, b in zip(tensors, tensors[1:]))
    if adjacent:
        total = sum(t.numel() for t in tensors)
        dist.all_reduce(torch.as_strided(tensors[0], (total,), (1,)), op=op, group=_STATE["group"])
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=op, group=_STATE["group"])
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def assemble_forward(out_feature, depth=None, normal=None, contrib_sum=None, contrib_max=None) -> None:
    """Each rank holds its tiles in zero-filled full-size planes: sum them; contrib_max by max."""
    _bucket_all_reduce([out_feature, depth, normal, contrib_sum], dist.ReduceOp.SUM)
    if contrib_max is not None and contrib_max.numel() > 0:
        dist.all_reduce(contrib_max, op=dist.ReduceOp.MAX, group=_STATE["group"])


def reduce_accumulators(acc: torch.Tensor) -> None:
    """All-reduce(sum) of the per-triangle screen-space gradient accumulators (16 fp32 per triangle) between the composite
    backward and the per-triangle backward: 64 B/triangle on the wire instead of the 240+ B/triangle of the final gradients
    (vertex 36 + SH 192 + opacity 4 + center2D 8), and the per-triangle stage then runs replicated on identical data."""
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=_STATE["group"])


def reduce_gradients(*grads) -> None:
    """NCCL all-reduce(sum) of the per-triangle gradient tensors (partial sums over each rank's tiles)."""
    _bucket_all_reduce(list(grads), dist.ReduceOp.SUM)
