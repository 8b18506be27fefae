"""Numpy test double of ``wdbx_b200.Engine`` (TEST INFRASTRUCTURE, lives under tests/ only).

Lets the host-side logic (VectorStore bookkeeping, shard striping, id mapping, filter / threshold
semantics, the SPMD all-gather + merge flow over gloo) run in the CPU container.  Scores come
from the oracle; keys are packed exactly like csrc/common.cuh so merges behave identically.
The product never imports this file.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import exact_search as oracle


def mono_u32(s: np.ndarray) -> np.ndarray:
    s = np.array(s, dtype=np.float32, copy=True)
    s[np.isnan(s)] = -np.inf
    s = s + np.float32(0.0)
    b = s.view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def unmono_f32(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.uint32)
    b = np.where(m & 0x80000000, m & np.uint32(0x7FFFFFFF), ~m).astype(np.uint32)
    return b.view(np.float32)


def pack_keys(scores: np.ndarray, gids: np.ndarray) -> np.ndarray:
    return (mono_u32(scores).astype(np.uint64) << np.uint64(32)) | (~np.asarray(gids, dtype=np.uint32)).astype(np.uint64)


def unpack_keys(keys: np.ndarray):
    keys = np.asarray(keys, dtype=np.uint64)
    scores = unmono_f32((keys >> np.uint64(32)).astype(np.uint32))
    gids = (~(keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.int64)
    empty = keys == 0
    scores = np.where(empty, -np.inf, scores).astype(np.float32)
    gids = np.where(empty, -1, gids)
    return scores, gids


class FakeEngine:
    def __init__(self, device=0, dim=4, dtype="fp32", num_segments=1):
        self.device, self.dim, self.dtype, self.num_segments = device, dim, dtype, num_segments
        self.rows = [np.empty((0, dim), np.float32) for _ in range(num_segments)]
        self.gids = [np.empty(0, np.int64) for _ in range(num_segments)]
        self.dead = [np.empty(0, bool) for _ in range(num_segments)]
        self.next_gid = 0
        self.launches = 0

    # mutation
    def reserve(self, segment, rows):
        pass

    def append(self, segment, rows, gids=None):
        rows = rows.cpu().numpy() if isinstance(rows, torch.Tensor) else np.asarray(rows, np.float32)
        if rows.ndim == 1:
            rows = rows[None, :]
        if self.dtype == "bf16":
            rows = oracle.bf16_round(rows)
        n = rows.shape[0]
        first = self.rows[segment].shape[0]
        g = np.arange(self.next_gid, self.next_gid + n) if gids is None else np.asarray(gids, np.int64)
        self.rows[segment] = np.concatenate([self.rows[segment], rows])
        self.gids[segment] = np.concatenate([self.gids[segment], g])
        self.dead[segment] = np.concatenate([self.dead[segment], np.zeros(n, bool)])
        self.next_gid += n
        return first

    def overwrite(self, segment, row, vector):
        v = np.asarray(vector, np.float32)
        self.rows[segment][row] = oracle.bf16_round(v) if self.dtype == "bf16" else v
        self.dead[segment][row] = False

    def tombstone(self, segment, row, dead=True):
        self.dead[segment][row] = dead

    def clear(self, segment=-1):
        for s in (range(self.num_segments) if segment < 0 else [segment]):
            self.rows[s] = np.empty((0, self.dim), np.float32)
            self.gids[s] = np.empty(0, np.int64)
            self.dead[s] = np.empty(0, bool)

    def read_row(self, segment, row):
        return self.rows[segment][row].copy()

    def read_rows(self, segment, row0, n):
        return self.rows[segment][row0:row0 + n].copy()

    # search
    def _keys(self, Q, k, metric, segs):
        Q = np.asarray(Q, np.float32).reshape(-1, self.dim)
        X = np.concatenate([self.rows[s] for s in segs])
        g = np.concatenate([self.gids[s] for s in segs])
        dead = np.concatenate([self.dead[s] for s in segs])
        keys = np.zeros((Q.shape[0], k), np.uint64)
        for b in range(Q.shape[0]):
            if X.shape[0] == 0:
                continue
            s = oracle.scores_fp32(X, Q[b], metric)
            kk = pack_keys(s, g)
            kk = np.sort(kk[~dead])[::-1][:k]
            keys[b, : len(kk)] = kk
        self.launches += 1
        return keys

    def search_host(self, queries, k, metric="cosine", per_segment=False, want_keys=False, segment=-1):
        if per_segment:
            keys = np.stack([self._keys(queries, k, metric, [s]) for s in range(self.num_segments)])
        else:
            keys = self._keys(queries, k, metric, range(self.num_segments) if segment < 0 else [segment])
        scores, gids = unpack_keys(keys)
        counts = (keys != 0).sum(-1).astype(np.int32)
        return (scores, gids, counts, keys) if want_keys else (scores, gids, counts)

    def search_filtered_host(self, queries, k, metric="cosine", min_score=float("-inf"), allow=None):
        Q = np.asarray(queries, np.float32).reshape(-1, self.dim)
        keys = np.zeros((Q.shape[0], k), np.uint64)
        X = np.concatenate(self.rows)
        g = np.concatenate(self.gids)
        dead = np.concatenate(self.dead).copy()
        if allow is not None:
            ok = []
            for s in range(self.num_segments):
                n = self.rows[s].shape[0]
                if allow[s] is None:
                    ok.append(np.ones(n, bool))
                else:
                    bits = np.unpackbits(np.asarray(allow[s], np.uint32).view(np.uint8), bitorder="little")[:n]
                    ok.append(bits.astype(bool))
            dead |= ~np.concatenate(ok)
        for b in range(Q.shape[0]):
            if X.shape[0] == 0:
                continue
            s_ = oracle.scores_fp32(X, Q[b], metric)
            live = ~dead & ~(s_ < min_score)
            kk = np.sort(pack_keys(s_, g)[live])[::-1][:k]
            keys[b, : len(kk)] = kk
        self.launches += 1
        scores, gids = unpack_keys(keys)
        return scores, gids, (keys != 0).sum(-1).astype(np.int32)

    def upload(self, queries):
        q = np.ascontiguousarray(queries, np.float32)
        return torch.from_numpy(q[None, :] if q.ndim == 1 else q)

    def search(self, q_dev, k, metric="cosine", segment=-1, out=None, stream=None):
        keys = self._keys(q_dev.numpy(), k, metric, range(self.num_segments) if segment < 0 else [segment])
        scores, gids = unpack_keys(keys)
        return {"keys": torch.from_numpy(keys.view(np.int64)), "scores": torch.from_numpy(scores),
                "gids": torch.from_numpy(gids), "counts": torch.from_numpy((keys != 0).sum(-1).astype(np.int32))}

    def merge(self, keys, out=None, stream=None):
        kk = keys.numpy().view(np.uint64)  # [G, B, k]
        G, B, k = kk.shape
        flat = np.transpose(kk, (1, 0, 2)).reshape(B, G * k)
        merged = np.sort(flat, axis=1)[:, ::-1][:, :k].copy()
        scores, gids = unpack_keys(merged)
        self.launches += 1
        return {"keys": torch.from_numpy(merged.view(np.int64)), "scores": torch.from_numpy(scores),
                "gids": torch.from_numpy(gids), "counts": torch.from_numpy((merged != 0).sum(-1).astype(np.int32))}

    def set_tuning(self, *a, **kw):
        pass

    def set_kernel_timing(self, enable=True):
        pass

    def stats(self):
        return {"kernel_launches": self.launches, "rows_total": int(sum(r.shape[0] for r in self.rows)),
                "rows_live": int(sum((~d).sum() for d in self.dead))}

    def close(self):
        pass


class FakeGroup:
    """Numpy double of the C group (wdbx_b200_group_*): every fake engine answers for its stripe, the packed keys
    are merged on the host exactly as the merge kernel / the on-device exchange would."""

    NAMES = {0: "cosine", 1: "ip", 2: "l2"}

    def __init__(self, engines):
        self.engines = list(engines)
        self.calls = 0

    def close(self):
        pass

    @staticmethod
    def _merge(keys_per_engine, k):
        kk = np.concatenate(keys_per_engine, axis=-1)
        merged = np.sort(kk, axis=-1)[..., ::-1][..., :k].copy()
        scores, gids = unpack_keys(merged)
        return scores, gids, (merged != 0).sum(-1).astype(np.int32), merged

    def search_host(self, q, k, metric, segment, min_score, allow, want_keys):
        name = self.NAMES[int(metric)]
        self.calls += 1
        per = []
        for i, e in enumerate(self.engines):
            if allow is not None or min_score > float("-inf"):
                s, g, c = e.search_filtered_host(q, k, name, min_score, None if allow is None else allow[i])
                keys = np.where(np.arange(k)[None, :] < c[:, None], pack_keys(s, np.maximum(g, 0)), np.uint64(0))
            else:
                keys = e.search_host(q, k, name, per_segment=(segment == -2), want_keys=True,
                                     segment=(segment if segment >= 0 else -1))[3]
            per.append(keys)
        scores, gids, counts, merged = self._merge(per, k)
        return (scores, gids, counts, merged) if want_keys else (scores, gids, counts)

    def search_device(self, q_dev, k, metric, out, stream):
        scores, gids, counts, merged = self.search_host(q_dev.numpy(), k, metric, -1, float("-inf"), None, True)
        return {"keys": torch.from_numpy(merged.view(np.int64)), "scores": torch.from_numpy(scores),
                "gids": torch.from_numpy(gids), "counts": torch.from_numpy(counts)}


FakeEngine.group_factory = FakeGroup
