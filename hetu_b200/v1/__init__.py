"""Hetu v1 ("legacy") API surface on top of the new graph: `Variable`, `placeholder_op`, `*_op` constructors,
`Executor(eval_nodes, ctx=..., comm_mode=...)`.run(feed_dict), SGD/Momentum/AdaGrad/Adam optimizers with
`.minimize(loss)`, the parameter-server / hybrid communication modes, the HET embedding cache (cstable), the HetuMoE
layers and the v1 auto-parallel search strategies.
(ref: hetu/v1/python/hetu/{gpu_ops/executor.py, gpu_ops/*.py, optimizer.py, cstable.py, layers/, distributed_strategies/})
"""
from .executor import (Variable, placeholder_op, Executor, HetuConfig, gradients, cpu, gpu, rcpu, rgpu,  # noqa: F401
                       matmul_op, linear_op, relu_op, sigmoid_op, tanh_op, gelu_op, softmax_op, softmaxcrossentropy_op,
                       softmaxcrossentropy_sparse_op, add_op, mul_op, addbyconst_op, mulbyconst_op, reduce_mean_op, reduce_sum_op,
                       array_reshape_op, embedding_lookup_op, concat_op, dropout_op, layer_normalization_op, batch_matmul_op,
                       transpose_op, broadcastto_op, binarycrossentropy_op, mse_op, slice_op, sqrt_op, exp_op, log_op, conv2d_op,
                       conv2d_add_bias_op, max_pool2d_op, avg_pool2d_op, batch_normalization_op, instance_normalization2d_op, pad_op,
                       div_op, minus_op, opposite_op, abs_op, pow_op, rsqrt_op, leaky_relu_op, mish_op, silu_op, where_op, one_hot_op,
                       reduce_max_op, reduce_min_op, concatenate_op, split_op, sum_op, reset_graph)
from . import initializers, initializers as init, layers, lr_scheduler, metrics, dataloader, onnx  # noqa: F401
from .dataloader import Dataloader, dataloader_op, GNNDataLoaderOp, BatchIndices, RawData  # noqa: F401
from .optimizer import (SGDOptimizer, MomentumOptimizer, AdaGradOptimizer, AdamOptimizer, AMSGradOptimizer, AdamWOptimizer,  # noqa: F401
                        LambOptimizer)
from . import optimizer as optim  # noqa: F401  (v1: ht.optim.SGDOptimizer)
from .ps import PSContext, ShardedPSContext, CacheSparseTable  # noqa: F401
from . import strategies as dist  # noqa: F401
from . import lr_scheduler as lr, data  # noqa: F401
from .ndarray import NDArray, ND_Sparse_Array, array, empty, empty_like, sparse_array, IndexedSlices, is_gpu_ctx  # noqa: F401
from .context import DeviceGroup, NodeStatus, ContextStack  # noqa: F401
from .preduce import PartialReduce  # noqa: F401
from .logger import HetuLogger, WandbLogger  # noqa: F401
from .memory_pool import HetuMemoryPool  # noqa: F401
from . import stream  # noqa: F401
from ..data.tokenizers.wordpiece import BertTokenizer  # noqa: F401
from .profiler import HetuProfiler, NCCLProfiler, HetuSimulator, NCCLOP  # noqa: F401
from .strategies import (DataParallel, ModelParallel4CNN, ModelParallel4LM, OneWeirdTrick4CNN, MegatronLM, FlexFlowSearching,  # noqa: F401
                         OptCNNSearching, GPipeSearching, PipeDreamSearching, PipeOptSearching)
from ..models.moe import MoELayer, TopKGate, KTop1Gate, HashGate, BalanceGate, SAMGate  # noqa: F401
from .ops import *  # noqa: F401,F403,E402  (the long tail of v1 `*_op` constructors)
from .grad_ops import *  # noqa: F401,F403,E402  (explicit gradient-node constructors, quantised tables, pipeline send / receive)
from . import ops as gpu_ops  # noqa: F401,E402
from .runtime_api import (Communicator, wrapped_mpi_nccl_init, new_group_comm, get_mpi_communicate, get_nccl_communicate,  # noqa: F401,E402
                          scheduler_init, scheduler_finish, server_init, server_finish, worker_init, worker_finish, get_worker_communicate,
                          DistConfig, context, get_current_context, dispatch, softmax_func, random)
