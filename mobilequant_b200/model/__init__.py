from .hf_config import HFConfig
from .hf_model import HFForCausalLM, HFModel, HFDecoderLayer, HFAttention, HFMLP, HFRMSNorm
