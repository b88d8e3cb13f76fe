"""python -m hetu.models.llama.generate_llama_hetero_4d_config --config-path DIR --config-name NAME [...]
heterogeneous (Malleus / Ampelos) strategy JSON from the YAML section `ds_parallel`: `hetero_layers` = layers per stage for
every pipeline, `tp` (or `hetero_tp` per pipeline), optional `rank_to_device_mapping` / `unused_rank`.
(ref: python/hetu/models/llama/generate_llama_hetero_4d_config.py)"""
import os

from ...utils import hydra_lite
from ..parallel_config import generate_hetero_ds_parallel_config, save_ds_parallel_config


def generate_llama_hetero_4d_config(hetero_layers, tp=1, hetero_tp=None, rank_to_device_mapping=None, unused_rank=(), zero=False):
    mapping = {int(k): int(v) for k, v in (rank_to_device_mapping or {}).items()}
    pipelines, rank = [], 0
    for p, stages in enumerate(hetero_layers):
        t = (hetero_tp[p] if hetero_tp else tp)
        lo, pl = 0, []
        for nl in stages:
            devs = [mapping.get(r, r) for r in range(rank, rank + t) if r not in unused_rank]
            pl.append({"devices": devs, "layers": [lo, lo + nl - 1]})
            lo += nl
            rank += t
        pipelines.append({"stages": pl})
    return generate_hetero_ds_parallel_config(sum(hetero_layers[0]), pipelines, zero=zero)


def main(argv=None):
    c = hydra_lite.load(argv).ds_parallel
    cfg = generate_llama_hetero_4d_config([list(s) for s in c.hetero_layers], c.get("tp", 1), c.get("hetero_tp"), c.get("rank_to_device_mapping"),
                                       tuple(c.get("unused_rank") or ()), bool(c.get("zero", False)))
    out = os.path.join(c.get("ds_parallel_config_path", "."), c.get("ds_parallel_config_name", "hetero_ds_parallel_config.json"))
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    save_ds_parallel_config(cfg, out)
    print("wrote", out)
    return out


if __name__ == "__main__":
    main()
