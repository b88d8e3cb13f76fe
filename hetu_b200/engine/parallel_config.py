"""(ref: python/hetu/engine/parallel_config.py: generate_gpt_3d_config, config_spread_zero, read / save helpers)"""
from ..models.generate_ds_config import generate_gpt_3d_config, generate_gpt_4d_config, generate_llama_4d_config  # noqa: F401
from ..models.parallel_config import (generate_ds_parallel_config, generate_hetero_ds_parallel_config, read_ds_parallel_config,  # noqa: F401
                                      save_ds_parallel_config)


def config_spread_zero(ds_parallel_config: dict) -> dict:
    """copy the top-level `zero` flag onto every variable leaf"""
    zero = bool(ds_parallel_config.get("zero", False))

    def walk(node):
        if isinstance(node, dict):
            if node.get("type") == "variable":
                node["zero"] = zero
            for v in node.values():
                walk(v)
    walk(ds_parallel_config)
    return ds_parallel_config
