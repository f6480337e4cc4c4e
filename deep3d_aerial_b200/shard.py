"""Reference-view sharding across the GPUs of one node.

A scene block is a list of reference-view ids (`blocks.txt`, reference IO/params_io.py:430-444); every
reference view is an independent unit of work (mvs/mvs_cas/predict.py:126-188 loops over them serially
and writes per-view files), so the block is dealt out to the ranks -- one process per GPU -- and
NOTHING is exchanged on the data path.  The only communication is the host-side join of per-rank
results / timings at the end (`join_max`, `join_sum`, `gather_objects`), over whatever backend the
process group was created with (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import os
import sys
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def world() -> tuple:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def partition(view_ids: Sequence[int], world_size: int, rank: int, mode: str = "round_robin") -> List[int]:
    """The reference views rank `rank` owns.  round_robin keeps neighbouring views (which share source
    images) on different GPUs at the same time step; contiguous keeps them on the same GPU (better for a
    per-rank feature cache).  Every view is owned by exactly one rank."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank %d / world_size %d" % (rank, world_size))
    ids = list(view_ids)
    if mode == "round_robin":
        return ids[rank::world_size]
    if mode == "contiguous":
        base, extra = divmod(len(ids), world_size)
        start = rank * base + min(rank, extra)
        return ids[start:start + base + (1 if rank < extra else 0)]
    raise ValueError("unknown partition mode %r" % mode)


def bind_host_to_gpu(local_rank: int) -> bool:
    """Pin this process to the CPUs (and so, by first touch, the memory node) NVML reports as closest to its GPU.
    With one process per GPU every rank streams its views' features from pinned host memory at the same time
    (204 MB per view at the WHU-OMVS shape); ranks scheduled on the far socket halve that bandwidth.
    Best effort: returns False when NVML or the affinity call is unavailable."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        index = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                index = int(ids[local_rank])
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:  # noqa: BLE001 -- affinity is an optimisation, never a requirement
        return False


def init(backend: Optional[str] = None) -> tuple:
    """Create the process group when launched under torchrun (WORLD_SIZE > 1); returns world()."""
    rank, size, local = world()
    if size > 1 and torch.cuda.is_available() and os.environ.get("D3D_NO_AFFINITY") != "1":
        bind_host_to_gpu(local)
    if size > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        # NCCL prints its version banner on STDOUT when the first communicator is created -- next to the one JSON
        # line bench.py owes its caller.  Create the communicator here, with file descriptor 1 pointed at stderr.
        sys.stdout.flush()
        saved = os.dup(1)
        try:
            os.dup2(2, 1)
            dist.init_process_group(backend=backend, rank=rank, world_size=size)
            if backend == "nccl":
                dist.barrier()
                torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    return rank, size, local


def finalize() -> None:
    """Tear the process group down (call once, at the end of a rank's work)."""
    if dist.is_initialized():
        dist.destroy_process_group()


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def join_max(value: float) -> float:
    """max over ranks (timings: the job is as slow as its slowest rank)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def join_sum(value: float) -> float:
    """sum over ranks (units of work done)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_objects(obj) -> list:
    """Every rank's `obj` on every rank (per-view result records; small, host side)."""
    if not dist.is_initialized():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def run_block(view_ids: Sequence[int], fn, mode: str = "round_robin") -> dict:
    """Process this rank's share of a scene block: fn(view_id) -> result; returns {view_id: result} for
    the views owned here.  Restartable per view: a failed view raises with its id."""
    rank, size, _ = world()
    out = {}
    for vid in partition(view_ids, size, rank, mode):
        try:
            out[vid] = fn(vid)
        except Exception as exc:  # noqa: BLE001 -- re-raised with the view id attached
            raise RuntimeError("reference view %r failed on rank %d: %s" % (vid, rank, exc)) from exc
    return out
