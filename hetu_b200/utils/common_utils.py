"""(ref: python/hetu/utils/common_utils.py)"""
from __future__ import annotations

import os
import random
from typing import Any, Dict

import numpy as np


def set_random_seed(seed: int):
    import torch

    from ..core import set_seed
    random.seed(seed)
    np.random.seed(seed % (1 << 32))
    torch.manual_seed(seed)
    set_seed(seed)


def to_dict(obj) -> Dict[str, Any]:
    import dataclasses
    return dataclasses.asdict(obj) if dataclasses.is_dataclass(obj) else dict(vars(obj))


def env_flag(name: str, default: bool = False) -> bool:
    v = os.environ.get(name)
    return default if v is None else v.lower() not in ("0", "false", "off", "")
