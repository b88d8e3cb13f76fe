"""CLI: write a ds_parallel_config JSON for a (dp, cp, tp, pp, zero) strategy -- the role of the reference's
python/hetu/models/{gpt,llama}/generate_*_4d_config.py and engine/parallel_config.py::generate_gpt_3d_config.

    python -m hetu_b200.models.generate_ds_config --model llama --num-layers 32 --num-gpus 8 --dp 2 --tp 2 --pp 2 --zero \\
        --out ds_parallel_config/gpus8/dp2_tp2_pp2.json
    python -m hetu_b200.models.generate_ds_config --hetero-layers 20,12 12,20 --hetero-tp 2 --out hetero.json   (Malleus form)
"""
import argparse
import os

from .parallel_config import generate_ds_parallel_config, generate_hetero_ds_parallel_config, save_ds_parallel_config


def generate_gpt_4d_config(num_layers=32, num_gpus=8, dp=2, cp=1, tp=2, pp=2, zero=True, recompute_layers=()):
    return generate_ds_parallel_config(num_layers, num_gpus, dp, tp, pp, cp=cp, zero=zero, recompute_layers=recompute_layers, model="gpt")


def generate_llama_4d_config(num_layers=32, num_gpus=8, dp=2, cp=1, tp=2, pp=2, zero=True, recompute_layers=()):
    return generate_ds_parallel_config(num_layers, num_gpus, dp, tp, pp, cp=cp, zero=zero, recompute_layers=recompute_layers, model="llama")


def generate_gpt_3d_config(rank_to_device_mapping=None, unused_rank=(), hetero_layers=None, hetero_stages=None, num_layers=32, num_gpus=8,
                           dp=2, tp=2, pp=2, zero=True):
    """homogeneous call == generate_gpt_4d_config; with hetero_layers ([[layers per stage] per pipeline]) the Malleus form"""
    if not hetero_layers:
        return generate_gpt_4d_config(num_layers, num_gpus, dp, 1, tp, pp, zero)
    mapping = rank_to_device_mapping or {}
    pipelines, rank = [], 0
    for stages in hetero_layers:
        lo, pl = 0, []
        for nl in stages:
            devs = [mapping.get(r, r) for r in range(rank, rank + tp) if r not in unused_rank]
            pl.append({"devices": devs, "layers": [lo, lo + nl - 1]})
            lo += nl
            rank += tp
        pipelines.append({"stages": pl})
    return generate_hetero_ds_parallel_config(num_layers, pipelines, zero=zero)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt", choices=["gpt", "llama"])
    ap.add_argument("--num-layers", type=int, default=32)
    ap.add_argument("--num-gpus", type=int, default=8)
    ap.add_argument("--dp", type=int, default=1); ap.add_argument("--cp", type=int, default=1)
    ap.add_argument("--tp", type=int, default=1); ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--zero", action="store_true")
    ap.add_argument("--recompute-layers", type=str, default="")
    ap.add_argument("--hetero-layers", nargs="*", default=None, help="one comma-separated list of layers-per-stage per pipeline")
    ap.add_argument("--hetero-tp", type=int, default=0)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.hetero_layers:
        hl = [[int(v) for v in s.split(",")] for s in a.hetero_layers]
        cfg = generate_gpt_3d_config(hetero_layers=hl, num_layers=sum(hl[0]), tp=a.hetero_tp or a.tp, zero=a.zero)
    else:
        rl = [int(v) for v in a.recompute_layers.split(",") if v]
        fn = generate_llama_4d_config if a.model == "llama" else generate_gpt_4d_config
        cfg = fn(a.num_layers, a.num_gpus, a.dp, a.cp, a.tp, a.pp, a.zero, rl)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    save_ds_parallel_config(cfg, a.out)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
