from .config import LoraConfig  # noqa: F401
from .layer import (LoraColumnParallelLinear, LoraRowParallelLinear, MultiLoraColumnParallelLinear,  # noqa: F401
                    MultiLoraRowParallelLinear)
from .model import LoraModel, MultiLoraModel, get_peft_model, wrap_model_factory, merge_lora_weights, lora_state_dict  # noqa: F401
