"""Synthetic batches of the shapes SURVEY.md §8(d) fixes for the benchmark (there is no dataset offline):
images ~ N(0,1) (ImageNet-normalised statistics), captions = SOT + U{4..40} random ids + EOT + zero padding,
2 global crops 256² + 8 local crops 96² per image, iBOT masks = exactly 30 % of the patches on 50 % of the global
crops.  All tensors are created on the CPU (optionally pinned) so that the end-to-end benchmark can time the H2D copy."""
from __future__ import annotations

from typing import Dict

import torch


def make_batch(B: int, *, image_size: int = 256, local_size: int = 96, n_local: int = 8, vocab: int = 49408,
               context: int = 77, mask_ratio: float = 0.3, mask_prob: float = 0.5, seed: int = 1234,
               pin: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    HW = (image_size // 16) ** 2
    ids = torch.zeros(B, context, dtype=torch.long)
    lens = torch.randint(4, 41, (B,), generator=g)
    body = torch.randint(1, vocab - 2, (B, context), generator=g)
    pos = torch.arange(context)[None, :]
    ids = torch.where((pos >= 1) & (pos <= lens[:, None]), body, ids)
    ids[:, 0] = vocab - 2
    ids[torch.arange(B), lens + 1] = vocab - 1
    n_mask = int(round(mask_ratio * HW))
    masks = torch.zeros(2 * B, HW, dtype=torch.bool)
    chosen = torch.randperm(2 * B, generator=g)[: max(1, int(round(mask_prob * 2 * B)))]
    for i in chosen.tolist():
        masks[i, torch.randperm(HW, generator=g)[:n_mask]] = True
    mask_idx = masks.flatten().nonzero().flatten()
    mw = (1.0 / masks.sum(-1).clamp(min=1).float())[:, None].expand_as(masks)[masks].contiguous()
    batch = dict(
        image=torch.randn(B, 3, image_size, image_size, generator=g),
        text=ids,
        global_crops=torch.randn(2 * B, 3, image_size, image_size, generator=g),
        local_crops=torch.randn(n_local * B, 3, local_size, local_size, generator=g),
        mask_indices=mask_idx,
        masks_weight=mw,
        rec_image=torch.randn(B, 3, image_size, image_size, generator=g),
    )
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def to_device(batch: Dict[str, torch.Tensor], device, non_blocking: bool = True) -> Dict[str, torch.Tensor]:
    return {k: v.to(device, non_blocking=non_blocking) for k, v in batch.items()}


def batch_bytes(batch: Dict[str, torch.Tensor]) -> int:
    return sum(v.numel() * v.element_size() for v in batch.values())


class BatchPrefetcher:
    """Host -> device input pipeline of the training loop: the pinned host batch of step i+1 is copied on a side
    stream while step i computes (what a pinned-memory DataLoader + `non_blocking` prefetch does around the reference).

        pf = BatchPrefetcher(device); pf.put(host_batch)
        for ...:
            batch, slot = pf.get(); pf.put(next_host_batch)
            loss = trainer.train_step(batch); pf.release(slot)

    `depth` device buffers rotate; a buffer is overwritten only after the step that consumed it (event recorded by
    `release`) and is handed out only after its copy landed (event recorded by `put`)."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.buffers = [None] * depth
        self.copied = [None] * depth   # event: H2D copy of this slot finished
        self.freed = [None] * depth    # event: the step that used this slot finished
        self.queue = []
        self.next_slot = 0

    def put(self, host_batch: Dict[str, torch.Tensor]) -> None:
        slot = self.next_slot
        self.next_slot = (slot + 1) % self.depth
        assert slot not in self.queue, "prefetch depth exceeded: get() a batch before putting another"
        with torch.cuda.stream(self.stream):
            if self.freed[slot] is not None:
                self.stream.wait_event(self.freed[slot])
            buf = self.buffers[slot]
            if buf is None or any(buf[k].shape != v.shape or buf[k].dtype != v.dtype for k, v in host_batch.items()):
                buf = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host_batch.items()}
                self.buffers[slot] = buf
            for k, v in host_batch.items():
                buf[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.copied[slot] = ev
        self.queue.append(slot)

    def get(self):
        slot = self.queue.pop(0)
        torch.cuda.current_stream(self.device).wait_event(self.copied[slot])
        return self.buffers[slot], slot

    def release(self, slot: int) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.freed[slot] = ev
