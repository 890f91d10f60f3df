"""Model switches of the unified decoder (mirror of mobilellm/model/hf_config.py:96-137).  A plain attribute bag --
no dependency on `transformers` -- that reads/writes the same config.json keys."""
import json, os

_DEFAULTS = dict(
    vocab_size=51200, hidden_size=2048, intermediate_size=8192, head_dim=None, num_hidden_layers=24,
    num_attention_heads=32, num_key_value_heads=None, resid_pdrop=0.0, embd_pdrop=0.0, attention_bias=False,
    attention_dropout=0.0, hidden_act="gelu_new", max_position_embeddings=2048, initializer_range=0.02,
    layer_norm_eps=1e-5, use_cache=True, tie_word_embeddings=False, rope_theta=10000.0, rope_scaling=None,
    partial_rotary_factor=1.0, qk_layernorm=False, bos_token_id=1, eos_token_id=2, pad_token_id=None,
    sliding_window=None, num_experts_per_tok=1, num_local_experts=1, mlp_bias=False, norm_class="rmsnorm",
    num_linears_per_mlp=3, shared_attention_norm=False, parallel_residual=False, normalize_embed=False,
    static_causal_mask=False, use_qkv_bias_only=False, use_matmul_as_module=False, l2norm_as_rmsnorm=False,
    torch_dtype="float32",
)


class HFConfig:
    model_type = "hfmodel"

    def __init__(self, **kw):
        for k, v in _DEFAULTS.items():
            setattr(self, k, kw.pop(k, v))
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        self._attn_implementation = kw.pop("_attn_implementation", "eager")
        self.extra = kw
        if self.num_local_experts != 1:
            raise NotImplementedError("MoE blocks (hf_model.py:1065-1162) are outside the MobileQuant hot path")

    def to_dict(self):
        d = {k: getattr(self, k) for k in _DEFAULTS}
        d["model_type"] = self.model_type
        d["architectures"] = ["HFForCausalLM"]
        return d

    def save_pretrained(self, out_dir):
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "config.json"), "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)

    @classmethod
    def from_pretrained(cls, path):
        with open(os.path.join(path, "config.json") if os.path.isdir(path) else path) as f:
            d = json.load(f)
        d.pop("model_type", None); d.pop("architectures", None)
        return cls(**d)


# Shapes of the three model families the reference evaluates (sim_model.py:42-47).
MODEL_SHAPES = {
    "tinyllama-1.1b": dict(vocab_size=32000, hidden_size=2048, intermediate_size=5632, num_hidden_layers=22,
                           num_attention_heads=32, num_key_value_heads=4, hidden_act="silu"),
    "gemma-2b": dict(vocab_size=256000, hidden_size=2048, intermediate_size=16384, num_hidden_layers=18,
                     num_attention_heads=8, num_key_value_heads=1, head_dim=256, hidden_act="gelu",
                     normalize_embed=True, tie_word_embeddings=True),
    "stablelm-2-1.6b": dict(vocab_size=100352, hidden_size=2048, intermediate_size=5632, num_hidden_layers=24,
                            num_attention_heads=32, num_key_value_heads=32, hidden_act="silu", norm_class="layernorm",
                            attention_bias=True, use_qkv_bias_only=True, partial_rotary_factor=0.25),
}


def named_config(name, **over):
    kw = dict(MODEL_SHAPES[name], use_cache=False, use_matmul_as_module=True, l2norm_as_rmsnorm=True)
    kw.update(over)
    return HFConfig(**kw)
