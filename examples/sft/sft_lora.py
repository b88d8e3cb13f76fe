"""Supervised fine-tuning on chat data with LoRA adapters on every parallel linear."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200.data import ByteTokenizer
from hetu_b200.engine import ModelWrapper, SFTConfig, SFTTrainer
from hetu_b200.models import GPTConfig, GPTLMHeadModel
from hetu_b200.peft import lora_state_dict, merge_lora_weights

ht.init_comm_group(1)
records = [{"messages": [{"role": "user", "content": f"what is {i} + {i}?"}, {"role": "assistant", "content": f"{i} + {i} = {2 * i}"}]} for i in range(64)]
cfg = GPTConfig(vocab_size=259, n_positions=128, n_embd=128, n_layer=2, n_head=4)
sft = SFTConfig(packing=False, micro_batch_size=8, global_load_size=8, max_seq_length=64, steps=30, learning_rate=5e-3, lora_rank=8, pack_alignment=16)
trainer = SFTTrainer(sft, ModelWrapper(GPTLMHeadModel, cfg), ByteTokenizer(), None, records)
losses = trainer.train()
m = trainer.trainer_states.model
print(f"loss {losses[0]:.3f} -> {losses[-1]:.3f}; {len(lora_state_dict(m))} adapter tensors; merged state dict has {len(merge_lora_weights(m))} tensors")
