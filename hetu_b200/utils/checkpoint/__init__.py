from .ht_safetensors import (WEIGHTS_NAME, WEIGHTS_FORMAT, TEMP_SPLITS, SPLIT_DIMS, save_file, load_file, save_model, load_model,  # noqa: F401
                             temp_save, temp_load, temp_save_split, temp_load_split, split_keys_for_shard, assemble_from_splits)
from .model_saver import ModelSaver  # noqa: F401
from .legacy import save_checkpoint, load_checkpoint, load_checkpoint_from_megatron, convert_llama_hf_to_ht  # noqa: F401
from .converters import (convert_gpt2_hf_to_ht, convert_gpt2_ht_to_hf, convert_llama_ht_to_hf, examine_checkpoint)  # noqa: F401,E402
