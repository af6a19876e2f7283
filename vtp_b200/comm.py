"""Peer-memory plumbing of the contrastive exchange (collective C2; csrc/clip.cu).

Each rank owns ONE cudaMalloc'ed buffer  [signal pad: uint64 x 32 | image features B x E bf16 | text features B x E bf16]
whose CUDA IPC handle is exchanged once over the process group (64 bytes per rank, set-up time only).  After that the
per-step exchange involves no host and no NCCL call: ranks write their normalised features into their own buffer, cross a
flag barrier through the peers' signal pads, and the gather+logits kernel pulls every rank's rows over NVLink.

The buffer is aliased as torch tensors through `__cuda_array_interface__` so the feature-producing kernels write into
it directly (no staging copy).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import lib

_PAD_BYTES = 256  # uint64[32] signal pad, 256-byte aligned feature area behind it


class _RawCuda:
    """Minimal `__cuda_array_interface__` carrier so torch can alias memory that the C-ABI library allocated."""

    def __init__(self, ptr: int, nelem: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False), "version": 2}


def alias_bf16(ptr: int, shape, device) -> torch.Tensor:
    n = 1
    for s in shape:
        n *= s
    t = torch.as_tensor(_RawCuda(ptr, n, "<i2"), device=device)
    return t.view(torch.bfloat16).view(*shape)


class PeerFeatures:
    """Feature exchange buffers of one rank + the mapped addresses of every peer's."""

    def __init__(self, B: int, E: int, device, process_group=None, world: Optional[int] = None, rank: Optional[int] = None):
        """world / rank default to the process group's; a trainer that runs single-rank inside an initialised job (the
        global-batch reference of the 2-GPU test) passes world = 1 and takes no part in any collective."""
        import torch.distributed as dist

        self.B, self.E = B, E
        self.device = torch.device(device)
        self.pg = process_group
        live = dist.is_available() and dist.is_initialized()
        self.world = world if world is not None else (dist.get_world_size(process_group) if live else 1)
        self.rank = rank if rank is not None else (dist.get_rank(process_group) if live and self.world > 1 else 0)
        if self.world > 16:
            raise lib.VtpError("the peer-memory contrastive exchange supports at most 16 ranks (one NVSwitch domain)")
        self.feat_bytes = B * E * 2
        self.nbytes = _PAD_BYTES + 2 * self.feat_bytes
        with torch.cuda.device(self.device):
            self.base = lib.comm_alloc(self.nbytes)
            handle = lib.comm_get_handle(self.base)
        self._opened: List[int] = []
        if self.world > 1:
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, handle, group=process_group)
            self.bases = []
            with torch.cuda.device(self.device):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self.bases.append(self.base)
                    else:
                        p = lib.comm_open_handle(h)
                        self._opened.append(p)
                        self.bases.append(p)
            dist.barrier(group=process_group)  # everybody has mapped everybody before the first flag is written
        else:
            self.bases = [self.base]
        self.img = alias_bf16(self.base + _PAD_BYTES, (B, E), self.device)
        self.txt = alias_bf16(self.base + _PAD_BYTES + self.feat_bytes, (B, E), self.device)
        self.img_ptrs = [b + _PAD_BYTES for b in self.bases]
        self.txt_ptrs = [b + _PAD_BYTES + self.feat_bytes for b in self.bases]
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 0

    def barrier(self, poison: Optional[torch.Tensor] = None):
        """All ranks' preceding stream work (their feature writes / reads) is complete and visible after this.
        `poison`: fp32 device scalar (the step's loss slot) set to NaN if the bounded wait times out, so the failure is
        seen by whoever reads the step's result — no extra synchronisation on the hot path."""
        self.epoch += 1
        lib.comm_barrier(self.bases, self.rank, self.epoch, self.err, poison)

    def check(self):
        """Host-side check of the barrier time-out flag (synchronises; call it off the hot path)."""
        if int(self.err.item()) != 0:
            raise lib.VtpError("peer-memory barrier timed out: a rank did not reach the contrastive exchange")

    def close(self):
        if getattr(self, "base", None) is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.comm_close_handle(p)
            lib.comm_free(self.base)
        self._opened, self.base = [], None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
