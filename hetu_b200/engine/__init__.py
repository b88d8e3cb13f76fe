from .trainer_config import TrainingConfig, SFTConfig, DataLoadLevel  # noqa: F401
from .wrapper import ModelWrapper, ModelWrapperFromConfig, OptimizerWrapper, DatasetWrapper  # noqa: F401
from .trainer import Trainer, TrainerStates  # noqa: F401
from .sft_trainer import SFTTrainer  # noqa: F401
from .straggler import Straggler, WorkloadInfo  # noqa: F401
from .strategy import StrategyModel, TPGroup, LayersProp, TrainerCtxs, TrainerStrategyArgs  # noqa: F401
from .strategy_ampelos import AmpelosStrategyModel, HMP, partition_into_k_groups, replan_after_failure  # noqa: F401
from .data_collator import DataCollatorForLanguageModel  # noqa: F401
from .hydraulis import StrategyCost, dispatch_batch, HydraulisPlanner  # noqa: F401
from .config_loader import load_experiment, build_trainer  # noqa: F401
from .hot_trainers import HotSPaTrainer, MalleusTrainer  # noqa: F401
from .hetero import HeteroSession  # noqa: F401
from . import lobra  # noqa: F401
