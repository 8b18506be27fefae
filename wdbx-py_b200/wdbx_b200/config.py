"""Minimal WDBXConfig mirror (reference: wdbx/core/config.py:14-314).

Only what the search path reads is kept: ``get / set / has / get_typed`` over
defaults < ``WDBX_*`` environment < runtime dict (config.py:64-80).  Engine keys are read both
bare and ``WDBX_``-prefixed, because the reference's core reads bare keys while its env/YAML
loaders only ever produce prefixed ones.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional

DEFAULTS: Dict[str, Any] = {
    "VECTOR_STORE_SAVE_IMMEDIATELY": False,
    "GPU_DEVICE": None,         # default: LOCAL_RANK or 0
    "GPU_DEVICES": None,        # single process, several GPUs: "all", "0-7", "0,1,2,3" or a list (multi_engine.py)
    "GPU_DTYPE": "fp32",        # fp32 | bf16 (storage)
    "GPU_METRIC": "cosine",     # cosine | ip | l2
    "GPU_CAPACITY_ROWS": 0,     # rows to reserve per shard up front
    "GPU_STRICT": False,        # raise instead of "log + []" on engine errors
    "GPU_PREFILTER": False,     # opt-in: filter_metadata becomes a device-side PRE-filter (full k among matches)
    "GPU_OVERLAP": False,        # device-resident searches (search_device) may overlap on the device; see wdbx_b200.h "overlap"
    "GPU_FUSED_EXCHANGE": True,  # multi-GPU: fuse the cross-GPU merge into the scan kernel (NVLink P2P)
    "GPU_BATCH_WINDOW_US": 200,  # micro-batching window of vector_search_async
    "GPU_BATCH_MAX": 64,         # queries coalesced into one pass by the async front-end (10M x 768: 64 queries
                                 # cost 2.46 ms on the filter path, one query 2.24 ms)
}


def _parse(value: str) -> Any:
    low = value.strip().lower()
    if low in ("true", "yes", "on"):
        return True
    if low in ("false", "no", "off"):
        return False
    for cast in (int, float):
        try:
            return cast(value)
        except ValueError:
            pass
    if value[:1] in "[{":
        try:
            return json.loads(value)
        except ValueError:
            pass
    return value


class WDBXConfig:
    def __init__(self, config_dict: Optional[Dict[str, Any]] = None, config_path: Optional[str] = None):
        self.config_dict: Dict[str, Any] = dict(DEFAULTS)
        if config_path and os.path.exists(config_path):
            try:
                with open(config_path) as f:
                    self.config_dict.update(json.load(f))
            except (OSError, ValueError):
                pass
        for k, v in os.environ.items():
            if k.startswith("WDBX_") and not k.startswith("WDBX_B200_"):
                self.config_dict[k] = _parse(v)
        if config_dict:
            self.config_dict.update(config_dict)

    def get(self, key: str, default: Any = None) -> Any:
        if key in self.config_dict and self.config_dict[key] is not None:
            return self.config_dict[key]
        alt = key[5:] if key.startswith("WDBX_") else "WDBX_" + key
        if alt in self.config_dict and self.config_dict[alt] is not None:
            return self.config_dict[alt]
        return default

    def set(self, key: str, value: Any) -> None:
        self.config_dict[key] = value

    def has(self, key: str) -> bool:
        return key in self.config_dict

    def get_typed(self, key: str, expected_type: type, default: Any = None) -> Any:
        v = self.get(key, default)
        if isinstance(v, expected_type):
            return v
        try:
            if expected_type is bool and isinstance(v, str):
                return bool(_parse(v))
            return expected_type(v)
        except (TypeError, ValueError):
            return default

    def get_all(self) -> Dict[str, Any]:
        return dict(self.config_dict)

    def __getitem__(self, key):
        return self.config_dict[key]

    def __setitem__(self, key, value):
        self.config_dict[key] = value

    def __contains__(self, key):
        return key in self.config_dict

    def __len__(self):
        return len(self.config_dict)
