"""Swarm-wide keyframe exchange: ONE all-gather of the fixed-size keyframe record per keyframe round.

Replaces LoopNet::broadcast_fisheye_desc / image_desc_callback (/root/reference/swarm_loop/src/loop_net.cpp:20-120,
142-172): the reference multicasts one LCM header message (global descriptor) plus one message per landmark over UDP
and reassembles them with timeouts; on the 8-GPU box the 8 drones are 8 ranks and the same information moves as one
`all_gather_into_tensor` of `lib.RECORD_BYTES` bytes per rank over NVLink (NCCL).  The reference's "skip my own
messages" rule (loop_net.cpp:133-136) becomes "my own slot goes to the local database" -- handled by
osb_frontend_ingest, which routes each record by its drone_id.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import lib as _l


def exchange_records(record: torch.Tensor, gathered: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """record: uint8 tensor [RECORD_BYTES] (this rank's osb_keyframe_record, device memory under NCCL, host memory
    under gloo) -> uint8 tensor [world * RECORD_BYTES] holding every rank's record in rank order."""
    assert record.dtype == torch.uint8 and record.numel() == _l.RECORD_BYTES
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if gathered is None:
        gathered = torch.empty(world * _l.RECORD_BYTES, dtype=torch.uint8, device=record.device)
    if world == 1:
        gathered.copy_(record)
        return gathered
    dist.all_gather_into_tensor(gathered, record, group=group)
    return gathered


def record_view(gathered: torch.Tensor, rank: int) -> torch.Tensor:
    return gathered[rank * _l.RECORD_BYTES:(rank + 1) * _l.RECORD_BYTES]


def routing(gathered_host: torch.Tensor, self_id: int):
    """Host-side statement of what osb_frontend_ingest does with a gathered buffer (used by the CPU tests):
    -> list of (slot, drone_id, msg_id, 'local' | 'remote', [directions with landmark_num > 0])."""
    out = []
    n = gathered_host.numel() // _l.RECORD_BYTES
    raw = gathered_host.cpu().numpy().tobytes()
    for r in range(n):
        rec = _l.KeyframeRecord.from_buffer_copy(raw[r * _l.RECORD_BYTES:(r + 1) * _l.RECORD_BYTES])
        dirs = [d for d in range(min(rec.n_dirs, _l.MAX_DIRS)) if rec.n_kpts[d] > 0]      # loop_detector.cpp:153
        out.append((r, rec.drone_id, rec.msg_id, "local" if rec.drone_id == self_id else "remote", dirs))
    return out
