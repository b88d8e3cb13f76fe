"""save: train a tiny GPT 2 steps under (dp, tp), checkpoint, report the loss of step 3.
load: build the same model under another (dp, tp), load the checkpoint, report the loss of its first step."""
import json, os, sys
import numpy as np, torch
import hetu_b200 as ht
from hetu_b200.models import GPTConfig, GPTLMHeadModel, generate_ds_parallel_config
from hetu_b200.utils.checkpoint import temp_load_split, temp_save_split

mode, dp, tp, path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
world = dp * tp
ht.init_comm_group(world); rank = int(os.environ.get("RANK", "0")); ht.set_seed(11)
S, Bg = 16, 8
cfg = GPTConfig(vocab_size=128, n_positions=S, n_embd=32, n_layer=2, n_head=4)
with ht.graph("define_and_run", create_new=True) as g:
    dsc = [generate_ds_parallel_config(2, world, dp, tp, 1)]
    model = GPTLMHeadModel(cfg, dsc)
    in_ds, in_dg = ht.nn.parallel.config2ds(dsc[0]["input"])
    ids, pos, lab = (ht.parallel_placeholder("int64", [Bg * S], [in_ds], device_group_hierarchy=[in_dg], name=n) for n in ("ids", "pos", "lab"))
    loss = model(ids, pos, lab, seq_len=S)
    opt = ht.AdamOptimizer(lr=1e-2)
    train_op = opt.minimize(loss)
rng = np.random.RandomState(0)
X = rng.randint(0, 128, (Bg, S)); L = np.roll(X, -1, axis=1); P = np.tile(np.arange(S), (Bg, 1))
d_idx = rank // tp
per = Bg // dp
sh = lambda a: torch.as_tensor(a[d_idx * per:(d_idx + 1) * per].reshape(-1))
feed = {ids: sh(X), pos: sh(P), lab: sh(L)}


def step():
    out = g.run(loss, [loss, train_op], feed, grad_scale=1.0 / dp)
    lv = out[0].float().reshape(1)
    if dp > 1:
        lv = ht._C.comm_all_reduce(lv, [d * tp + rank % tp for d in range(dp)], "sum") / dp
    return float(lv)


if mode == "save":
    step(); step()
    temp_save_split(model, opt, path, step=2)
    nxt = step()
else:
    g.run(loss, [loss], feed, run_level="alloc")          # materialise parameters, then overwrite from disk
    for p in model.parameters():
        g.get_param(p)
    for p in model.parameters():
        for st in opt.get_states(p).values():
            g.get_param(st)
    loaded, missing = temp_load_split(model, opt, path, strict=True)
    nxt = step()
if rank == 0:
    print("CKPT " + json.dumps({"next_loss": nxt}))
