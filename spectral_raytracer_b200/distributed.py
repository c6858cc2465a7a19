"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), frames sharded across
ranks, one sum-reduce of the spectral accumulation buffers onto the root.

A sample depends only on (pixel, frame_id, intended_frames) (shader.rs:280-281, :389-391) and the image is
the mean over frames (main.rs:1316), so ranks need no exchange while rendering; the only collective is the
final reduce of W*H*n_lambda f32 (265 MB at 1080p, 1.06 GB at 4K).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def frame_shard(first_frame: int, n_frames: int, rank: int, world: int):
    """Contiguous block of frame ids for `rank`: (first, count); the remainder goes to the low ranks."""
    base, rem = divmod(n_frames, world)
    count = base + (1 if rank < rem else 0)
    start = first_frame + rank * base + min(rank, rem)
    return start, count


class _DevicePtr:
    """Expose a raw device allocation through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


def accum_as_tensor(renderer) -> torch.Tensor:
    """The context's spectral accumulation buffer as a torch tensor that aliases the device memory
    (srt_accum_device_ptr), for use with torch.distributed collectives."""
    if hasattr(renderer, "accum_tensor"):  # (CPU stand-ins of the gloo tests)
        return renderer.accum_tensor()
    ptr, n = renderer.accum_device_ptr()
    dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.as_tensor(_DevicePtr(ptr, n), device=dev)
    assert t.data_ptr() == ptr
    return t


def reduce_sum_(buffer: torch.Tensor, frames_local: int, dst: int = 0, group=None) -> int:
    """Sum `buffer` (accumulated radiance) of all ranks into rank `dst`, and add up the frame counts.
    Works for CUDA tensors (NCCL) and CPU tensors (gloo, used by the CPU tests).  Returns the total number
    of frames the reduced buffer holds; blocks until the result is usable by other streams.  The buffers of
    the other ranks are left as they are -- see reduce_contexts_ for the version that keeps contexts consistent."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(buffer, dst=dst, op=dist.ReduceOp.SUM, group=group)
        n = torch.tensor([frames_local], dtype=torch.int64, device=buffer.device)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
        if buffer.is_cuda:
            torch.cuda.current_stream(buffer.device).synchronize()
        return int(n.item())
    return frames_local


def reduce_contexts_(renderer, dst: int = 0, group=None) -> int:
    """The exchange step of a frame-sharded render: the accumulation buffers of all ranks are summed
    onto rank `dst`, whose context then holds the whole render (frame count = the sum of all ranks' counts); every
    other rank's context is cleared -- empty image, frame count 0 -- so that further rounds (progressive rendering,
    checkpoint resume) do not add its old radiance a second time.  Returns the total frame count."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = reduce_sum_(accum_as_tensor(renderer), renderer.frames_accumulated, dst=dst, group=group)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if rank == dst:
            renderer.frames_accumulated = total
        else:
            renderer.clear()
    return total


def render_sharded(renderer, first_frame: int, n_frames: int, *, frames_per_call: int = 64, dst: int = 0,
                   group=None, progress=None):
    """Render frames [first_frame, first_frame + n_frames) cooperatively: every rank renders its shard into its
    own accumulation buffer, then the buffers are reduced onto `dst`, whose context ends up holding the whole
    render (resolve it there) including whatever it held before the call; the other contexts end up empty.
    Returns the total frame count on `dst`."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    start, count = frame_shard(first_frame, n_frames, rank, world)
    done = 0
    while done < count:
        n = min(frames_per_call, count - done)
        renderer.render_frames(start + done, n)
        done += n
        if progress is not None and progress(done / max(1, count)) is False:
            break
    return reduce_contexts_(renderer, dst=dst, group=group)
