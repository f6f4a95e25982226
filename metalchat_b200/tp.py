"""Tensor-parallel wiring on the host: one process per GPU (torchrun), handles exchanged over torch.distributed.

The reference is single-device (nn/llama.h:86); this is the B200-side addition named by BASELINE.json: column-split
wq/wk/wv/w1/w3, row-split wo/w2, vocabulary-split head, all-reduce fused into the GEMV kernels over NVLink peer memory.
"""
from __future__ import annotations

import os

from . import capi


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def gather_blobs(blob: bytes, group=None) -> list[bytes]:
    """all-gather of one opaque byte string per rank, in rank order (works on gloo and nccl groups)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, blob, group=group)
    return out


def connect(model: capi.Llama, group=None):
    """Exchanges the IPC handles of every rank's exchange region and maps the peers' regions."""
    handles = gather_blobs(model.tp_export(), group)
    assert all(isinstance(h, (bytes, bytearray)) and len(h) == 64 for h in handles), "malformed IPC handle"
    model.tp_connect([bytes(h) for h in handles])


def use_nccl(model: capi.Llama, group=None):
    """Comparator (bench.py --tp-collective nccl): the all-reduces of a block go through ncclAllReduce instead of the fused exchange."""
    import torch.distributed as dist

    box = [capi.nccl_unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    model.tp_use_nccl(bytes(box[0]))


def create(dev: capi.Device, group=None, collective: str = "fused", **cfg) -> capi.Llama:
    """Creates this rank's shard of a tensor-parallel model and wires it to its peers."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    m = capi.Llama(dev, capi.llama_config(**cfg, tp_rank=rank, tp_world=world))
    if world > 1:
        connect(m, group)
        if collective == "nccl":
            use_nccl(m, group)
        else:
            assert collective == "fused", collective
    return m
