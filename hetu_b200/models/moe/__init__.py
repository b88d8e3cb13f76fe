from .moe_model import MoEConfig, MoELayer, GPTMoELMHeadModel, TopKGate, KTop1Gate, HashGate, BalanceGate, SAMGate  # noqa: F401
