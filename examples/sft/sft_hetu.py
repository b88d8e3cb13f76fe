"""YAML-driven supervised fine-tuning (the reference's examples/sft/sft_hetu.py with hydra configs):

    python examples/sft/sft_hetu.py --config-name gpt_lora trainer.steps=20 sft.lora_rank=4
    python examples/sft/sft_hetu.py --config-name gpt_lora --data chats.jsonl --save-adapters out/adapters.pt
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/sft/sft_hetu.py ds_parallel.dp=2

--data: a JSON / JSONL file of {"messages": [{"role": ..., "content": ...}, ...]} records (or alpaca-style records with
sft.dataset_format=alpaca); without it a small synthetic arithmetic chat set is used.  The loss covers the assistant turns only
(sft.train_on_prompt=false); with sft.lora_rank > 0 only the adapters train and --save-adapters writes them (and the merged weights)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.engine import build_trainer
from hetu_b200.peft import lora_state_dict, merge_lora_weights
from hetu_b200.utils.parallel import distributed_init

ap = argparse.ArgumentParser()
ap.add_argument("--config-path", default=os.path.join(os.path.dirname(__file__), "config"))
ap.add_argument("--config-name", default="gpt_lora")
ap.add_argument("--data", default=None)
ap.add_argument("--save-adapters", default=None)
ap.add_argument("--prompt", default=None, help="after training: let the tuned model answer this user message (single-process runs)")
ap.add_argument("overrides", nargs="*")
a = ap.parse_args()
path = os.path.join(a.config_path, a.config_name + ("" if a.config_name.endswith((".yaml", ".yml")) else ".yaml"))
distributed_init()
overrides = list(a.overrides) + ([f"trainer.train_dataset_path={a.data}"] if a.data else [])
records = None if a.data else [{"messages": [{"role": "user", "content": f"what is {i} + {i}?"}, {"role": "assistant", "content": f"{i} + {i} = {2 * i}"}]}
                               for i in range(64)]
trainer = build_trainer(path, overrides, train_dataset=records)
losses = trainer.train()
model = trainer.trainer_states.model
if ht.distributed.rank() in trainer._loss_ranks():
    print(f"steps {len(losses)}  loss {losses[0]:.3f} -> {losses[-1]:.3f}")
    adapters = lora_state_dict(model)
    print(f"{len(adapters)} adapter tensors, {sum(v.numel() for v in adapters.values())} trainable parameters")
    if a.prompt and ht.distributed.world_size() == 1:
        # the tuned model = base weights with the adapters merged in; decode greedily behind the chat template's user turn
        from hetu_b200.data import ChatTemplate
        from hetu_b200.models import Generator, generate_ds_parallel_config
        tok = trainer.tokenizer
        base = trainer.model_wrapper.inner if hasattr(trainer.model_wrapper, "inner") else trainer.model_wrapper
        cfg = base.model_config
        layers = getattr(cfg, "n_layer", None) or cfg.num_hidden_layers
        window = int(getattr(cfg, "n_positions", None) or cfg.max_position_embeddings)
        gen = Generator(lambda: base.model_class(cfg, [generate_ds_parallel_config(int(layers), 1, 1, 1, 1, zero=False)]), batch=1, window=window,
                        pad_id=getattr(tok, "pad_id", 0))
        gen.load_state_dict(merge_lora_weights(model))
        tpl = ChatTemplate()
        ids = tok.encode(tpl.user_prefix + a.prompt + tpl.turn_suffix + tpl.assistant_prefix, add_special_tokens=False)
        out = gen.generate([ids], max_new_tokens=24, eos_id=tok.encode(tpl.turn_suffix, add_special_tokens=False)[-1])[0]
        print("prompt:", a.prompt)
        print("answer:", tok.decode(out[len(ids):]).strip())
    if a.save_adapters:
        os.makedirs(os.path.dirname(os.path.abspath(a.save_adapters)), exist_ok=True)
        torch.save({"adapters": adapters, "merged": merge_lora_weights(model)}, a.save_adapters)
        print("saved", a.save_adapters)
