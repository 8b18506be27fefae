"""torch.distributed plumbing for the one-process-per-GPU layout (NCCL over NVLink 5 / NVSwitch).

The reference's "distributed" layer (wdbx/core/distributed.py) is a TCP ping server that never
carries vectors or results; the only real exchange on the search path is the cross-shard merge
(wdbx/core/vector_store.py:323-330).  With shards striped over ranks that exchange becomes: every
rank all-gathers its [B, k] packed 8-byte keys and merges the G lists with kernel K3.
"""
from __future__ import annotations

import os
from typing import Optional


class DistContext:
    """rank / world / device of this process; ``world == 1`` means no collective at all."""

    def __init__(self, rank: int = 0, world: int = 1, device: Optional[int] = None, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.device = device if device is not None else int(os.environ.get("LOCAL_RANK", "0"))

    @classmethod
    def from_env(cls, device: Optional[int] = None) -> "DistContext":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return cls(0, 1, device)
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend == "nccl":
                local = device if device is not None else int(os.environ.get("LOCAL_RANK", "0"))
                torch.cuda.set_device(local)
                dist.init_process_group(backend, device_id=torch.device("cuda", local))
            else:
                dist.init_process_group(backend)
        return cls(dist.get_rank(), dist.get_world_size(), device)

    # ------------------------------------------------------------------ collectives
    def all_gather_keys(self, keys):
        """keys: [B, k] int64 tensor (packed u64 ranking keys) -> [world, B, k] on every rank."""
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return keys.unsqueeze(0)
        out = torch.empty((self.world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
        if keys.is_cuda:
            dist.all_gather_into_tensor(out, keys.contiguous(), group=self.group)
        else:  # gloo (CPU tests)
            parts = list(out.unbind(0))
            dist.all_gather(parts, keys.contiguous(), group=self.group)
        return out

    def all_gather_bytes(self, payload: bytes):
        """Gather one small bytes object per rank (rank order) -- IPC handles of the exchange buffers."""
        import torch.distributed as dist

        if self.world == 1:
            return [payload]
        out = [None] * self.world
        dist.all_gather_object(out, payload, group=self.group)
        return out

    def broadcast_array(self, arr, src: int):
        """Broadcast a numpy fp32 array from rank `src` (used by VectorStore.get)."""
        import numpy as np
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return arr
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda(self.device)
        dist.broadcast(t, src=src, group=self.group)
        return t.cpu().numpy()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)
