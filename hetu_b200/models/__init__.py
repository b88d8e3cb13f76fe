from .gpt import GPTConfig, GPTLMHeadModel  # noqa: F401
from .llama import LlamaConfig, LlamaLMHeadModel  # noqa: F401
from .moe import MoEConfig, GPTMoELMHeadModel  # noqa: F401
from .parallel_config import generate_ds_parallel_config, read_ds_parallel_config  # noqa: F401
from .ctr import WDL, DeepFM, DCN, DeepCrossing, NCF  # noqa: F401
from .vision import LogReg, MLP, LeNet, CNN3, AlexNet, VGG, ResNet  # noqa: F401
from .rnn import RNN, LSTM  # noqa: F401
from .gnn import GCN, GCNLayer, GraphSageLayer, normalise_adjacency, partition_15d, dist_gcn_15d_forward  # noqa: F401
from .bert import BertConfig, BertModel, BertForPreTraining, BertForMaskedLM, BertForSequenceClassification, convert_bert_hf_to_ht  # noqa: F401,E402
from .transformer import Transformer, TransformerConfig, MultiHeadAttention  # noqa: F401,E402
from .generation import Generator  # noqa: F401,E402
