from .lora import (LoraConfig, LoraColumnParallelLinear, LoraRowParallelLinear, MultiLoraColumnParallelLinear,  # noqa: F401
                   MultiLoraRowParallelLinear, LoraModel, MultiLoraModel, get_peft_model, wrap_model_factory,
                   merge_lora_weights, lora_state_dict)
