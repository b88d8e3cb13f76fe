"""Reference module paths that have no file of their own here.

The reference spreads some functionality over many small modules (`hetu.utils.parallel.read_ds`, `hetu.rpc.pssh_start`,
`hetu.data.tokenizers.hf_tokenizer`, ...).  Here that code lives in fewer, larger modules; a user's imports keep working through
ONE table and an import hook instead of a file per path:

  * an entry with a single source and no name list is an ALIAS -- the very same module object (`python -m <alias>` runs the
    source module as `__main__`),
  * any other entry is a FACADE -- a module object whose namespace is assembled from the listed names of the listed modules.

`hetu/__init__.py` maps `hetu.X` onto `hetu_b200.X` with the same loader.
"""
import importlib
import importlib.abc
import importlib.util
import sys
import types

_PKG = __name__.rsplit(".", 1)[0]

# path (below the package) -> [(source module (below the package), names or None = the module itself)]
TABLE = {
    "context": [("core", ["autocast", "context", "control_dependencies", "cpu_offload", "graph", "merge_strategy", "profiler", "recompute",
                          "run_level", "subgraph"])],
    "data.data_collator": [("engine.data_collator", None)],
    "data.utils": [("data.bucket", ["generate_cp_pack_data", "get_sorted_batch_and_len", "pack_sequences", "pad_sequences"]),
                   ("data.dataloader", ["build_data_loader", "parallel_data_provider"])],
    "data.messages.message_template": [("data.messages", ["ChatTemplate", "build_chat_sample"])],
    "data.messages.prompt_template": [("data.messages", ["PromptTemplate"])],
    "data.messages.utils": [("data.messages", ["build_chat_sample"])],
    "data.tokenizers.tokenizer": [("data.tokenizers", ["ByteTokenizer", "build_tokenizer"])],
    "data.tokenizers.gpt2_tokenizer": [("data.tokenizers", ["GPT2BPETokenizer", "build_tokenizer"])],
    "data.tokenizers.hf_tokenizer": [("data.tokenizers", ["HFTokenizer", "build_tokenizer"])],
    "data.tokenizers.sentencepiece_tokenizer": [("data.tokenizers", ["SentencePieceTokenizer", "build_tokenizer"])],
    "data.tokenizers.tiktoken_tokenizer": [("data.tokenizers", ["TikTokenizer", "build_tokenizer"])],
    "models.gpt.gpt_config": [("models.gpt.gpt_model", ["GPTConfig"])],
    "models.gpt.gpt_tokenizer": [("data.tokenizers", ["GPT2BPETokenizer"])],
    "models.llama.llama_config": [("models.llama.llama_model", ["LlamaConfig"])],
    "models.llama.llama_tokenizer": [("data.tokenizers", ["LlamaTokenizer", "SentencePieceTokenizer"])],
    "models.utils.converter.convert_llama_hf_to_ht": [("utils.checkpoint.legacy", ["convert_llama_hf_to_ht"])],
    "rpc.pssh_start": [("rpc.launcher", None)],
    "rpc.local_start": [("rpc.launcher", None)],
    "rpc.pssh_workers": [("rpc.launcher", None)],
    "rpc.pssh_start_elastic": [("rpc.elastic_server", None)],
    "rpc.heturpc_elastic_server": [("rpc.elastic_server", None)],
    "rpc.elastic_arg_parser": [("rpc.elastic_server", None)],
    "rpc.heturpc_async_server": [("rpc.heturpc_polling_server", None)],   # one threaded controller serves both variants
    "rpc.heturpc_async_server_exp": [("rpc.heturpc_polling_server", None)],
    "rpc.pssh_start_exp": [("rpc.launcher", None)],
    "rpc.kv_store.client": [("rpc.kv_store", ["KeyValueStoreClient", "RemoteDict"])],
    "rpc.kv_store.server": [("rpc.kv_store", ["KeyValueStoreServer"])],
    "rpc.kv_store.producer_consumer": [("rpc.kv_store", ["ProducerConsumer"])],
    "rpc.kv_store.const": [("rpc.kv_store", ["TIMEOUT", "DEFAULT_PORT"])],
    "engine.sft_config": [("engine.trainer_config", ["SFTConfig"])],
    "_binding": [("tools", [])],
    "_binding.codegen": [("tools", [])],
    "_binding.codegen.gen_py_ops": [("tools.gen_py_ops", ["gen_ops", "gen_stubs", "load_manifest", "check_manifest", "dump_registry", "main"])],
    "_binding.codegen.args_bridge": [("tools.gen_py_ops", ["ArgType", "parse_args"])],
    "engine.utils": [("engine.strategy", ["Args", "TrainerCtxs", "TrainerDatasetArgs", "TrainerStrategyArgs", "TrainerCommArgs", "TrainerCommAllArgs",
                                          "TrainerEnvs"])],
    "data.tokenizers.utils": [("data.tokenizers", ["SpecialToken", "BaseTokenizer"])],
    "data.tokenizers.bert_tokenizer": [("data.tokenizers.wordpiece", None)],
    "data.tokenizers.pretrained_tokenizer": [("data.tokenizers", ["PreTrainedTokenizer"])],
    "models.utils.common_utils": [("models.utils.pretrained", ["split_hetu_state_dict_into_shards", "split_state_dict_into_shards", "parse_size"])],
    "models.utils.config_utils": [("models.utils.pretrained", ["PreTrainedConfig", "CONFIG_NAME"])],
    "models.utils.hub": [("models.utils.pretrained", ["is_remote_url"])],
    "models.utils.model_utils": [("models.utils.pretrained", ["PreTrainedModel", "get_state_dict_dtype", "get_parameter_dtype", "WEIGHTS_INDEX_NAME"]),
                                 ("peft.lora.model", ["lora_state_dict"])],
    "models.utils.converter.convert_utils": [("utils.checkpoint.ht_safetensors", ["save_model", "save_file"])],
    "nn.functional": [("nn.init", ["generalized_xavier_", "xavier_uniform_", "xavier_normal_", "kaiming_uniform_", "kaiming_normal_", "lecun_uniform_",
                                   "lecun_normal_", "calculate_gain"])],
    "nn.parameter": [("nn", ["Parameter"])],
    "nn.modules": [("nn.module", "*"), ("nn.layers", "*"), ("nn.parallel", "*")],      # a package facade: its children are listed below
    "nn.modules.activation": [("nn.layers", ["ReLU", "GELU", "SiLU", "Sigmoid", "Tanh", "LeakyReLU", "Softmax", "NewGeLU"])],
    "nn.modules.batchnorm": [("nn.layers", ["BatchNorm"])],
    "nn.modules.container": [("nn.module", ["ModuleList", "Sequential", "ModuleDict"])],
    "nn.modules.conv": [("nn.layers", ["Conv2d"])],
    "nn.modules.dropout": [("nn.layers", ["Dropout", "Dropout2d"])],
    "nn.modules.instancenorm": [("nn.layers", ["InstanceNorm"])],
    "nn.modules.linear": [("nn.layers", ["Linear", "Identity"])],
    "nn.modules.loss": [("nn.layers", ["MSELoss", "BCELoss", "NLLLoss", "KLDivLoss", "CrossEntropyLoss", "SoftmaxCrossEntropySparse"])],
    "nn.modules.module": [("nn.module", ["Module"])],
    "nn.modules.normalization": [("nn.layers", ["LayerNorm", "RMSNorm"])],
    "nn.modules.padding": [("nn.layers", ["ConstantPad2d", "ZeroPad2d"])],
    "nn.modules.pooling": [("nn.layers", ["MaxPool2d", "AvgPool2d"])],
    "nn.modules.sparse": [("nn.layers", ["Embedding"])],
    "nn.modules.parallel": [("nn.parallel", None)],
    "nn.modules.parallel_ds": [("nn.parallel", None)],
    "nn.modules.parallel_multi_ds": [("nn.parallel", None)],
    "nn.modules.parallel_utils": [("nn.parallel", ["config2ds", "get_multi_ds_parallel_config"]),
                                  ("data.dataloader", ["parallel_data_provider"]),
                                  ("utils.parallel", ["get_local_index", "get_device_index"])],
    "nn.modules.utils": [("nn.layers", ["_nlist"])],
    "optim.optimizer": [("optim", ["Optimizer"])],
    "optim.sgd": [("optim", ["SGDOptimizer", "SGD"])],
    "utils.checkpoint.load_checkpoint": [("utils.checkpoint.legacy", ["convert_llama_hf_to_ht", "load_checkpoint", "load_checkpoint_from_megatron"])],
    "utils.checkpoint.save_checkpoint": [("utils.checkpoint.legacy", ["save_checkpoint"])],
    "utils.data.dataloader": [("data.dataloader", None)],
    "utils.data.dataset": [("data.dataset", None)],
    "utils.parallel.distributed": [("utils.parallel", ["distributed_init", "get_device_index", "get_local_index", "get_dg_from_union"])],
    "utils.parallel.ds_config": [("utils.parallel", ["RecomputeConfig", "StrategyConfig", "convert_strategy", "generate_recompute_config"])],
    "utils.parallel.generate_ds": [("models.parallel_config", ["generate_ds_parallel_config", "generate_hetero_ds_parallel_config",
                                                                "save_ds_parallel_config"]),
                                   ("utils.parallel", ["convert_strategy"])],
    "utils.parallel.read_ds": [("models.parallel_config", ["read_ds_parallel_config"]),
                               ("nn.parallel", ["config2ds", "get_multi_ds_parallel_config"]),
                               ("utils.parallel", ["parse_multi_ds_parallel_config"])],
}


class AliasLoader(importlib.abc.Loader):
    """hands the import machinery an existing module object under another name"""

    def __init__(self, real):
        self.real = real
        self.saved = (getattr(real, "__spec__", None), getattr(real, "__loader__", None))

    def create_module(self, spec):
        return self.real

    def get_code(self, fullname):       # `python -m <alias>` (runpy) executes the real module's code as __main__
        return self.saved[1].get_code(self.real.__name__)

    def exec_module(self, module):      # the import machinery re-stamped __spec__ / __loader__ with the alias: put the real ones back
        module.__spec__, module.__loader__ = self.saved


class FacadeLoader(importlib.abc.Loader):
    def __init__(self, sources):
        self.sources = sources

    def create_module(self, spec):
        return types.ModuleType(spec.name)

    def exec_module(self, module):
        names = []
        for src, wanted in self.sources:
            real = importlib.import_module(f"{_PKG}.{src}")
            if wanted == "*":
                wanted = [n for n in getattr(real, "__all__", None) or dir(real) if not n.startswith("_")]
            for n in wanted:
                setattr(module, n, getattr(real, n))
                names.append(n)
        module.__all__ = names
        module.__doc__ = "names re-exported from " + ", ".join(f"{_PKG}.{s}" for s, _ in self.sources)


class RefPathFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PKG + "."):
            return None
        sources = TABLE.get(fullname[len(_PKG) + 1:])
        if sources is None:
            return None
        if len(sources) == 1 and sources[0][1] is None:
            real = importlib.import_module(f"{_PKG}.{sources[0][0]}")
            return importlib.util.spec_from_loader(fullname, AliasLoader(real), origin=getattr(real.__spec__, "origin", None))
        rel = fullname[len(_PKG) + 1:]
        has_children = any(k.startswith(rel + ".") for k in TABLE)          # e.g. nn.modules -> nn.modules.linear
        return importlib.util.spec_from_loader(fullname, FacadeLoader(sources), is_package=has_children)


def install():
    if not any(isinstance(f, RefPathFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, RefPathFinder())
