"""python -m hetu.models.llama.generate_llama_4d_config --config-path DIR --config-name NAME [ds_parallel.tp=4 ...]
writes the homogeneous (dp, cp, tp, pp, zero) ds_parallel_config JSON described by the YAML section `ds_parallel`
(num_layers, num_gpus, dp, cp, tp, pp, zero, recompute.{...}, ds_parallel_config_path, ds_parallel_config_name).
(ref: python/hetu/models/llama/generate_llama_4d_config.py)"""
import os

from ...utils import hydra_lite
from ..generate_ds_config import generate_llama_4d_config  # noqa: F401
from ..parallel_config import save_ds_parallel_config


def main(argv=None):
    c = hydra_lite.load(argv).ds_parallel
    assert c.dp * c.get("cp", 1) * c.tp * c.pp == c.num_gpus, f"dp * cp * tp * pp != num_gpus {c.num_gpus}"
    rc = c.get("recompute") or {}
    layers = [i for sub in (rc.get("recompute_layer_idxs") or rc.get("layer_idxs") or []) for i in (sub if isinstance(sub, list) else [sub])]
    cfg = generate_llama_4d_config(c.num_layers, c.num_gpus, c.dp, c.get("cp", 1), c.tp, c.pp, bool(c.get("zero", True)), layers)
    out = os.path.join(c.get("ds_parallel_config_path", "."), c.get("ds_parallel_config_name", "ds_parallel_config.json"))
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    save_ds_parallel_config(cfg, out)
    print("wrote", out)
    return out


if __name__ == "__main__":
    main()
