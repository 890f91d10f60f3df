"""Unified decoder model -- eager dense path only (mirror of mobilellm/model/hf_model.py, the subset every PTQ / eval
script of the reference actually runs: `_attn_implementation="eager"`, `use_matmul_as_module=True`,
`l2norm_as_rmsnorm=True`, ptq/mobilequant.py:140-142).  Module and parameter names equal the reference's, so
`named_modules()` paths -- the keys of act_dict.json / default_qcfg.json / parameters.pth -- are identical.

Flash/SDPA attention variants (hm:552-1027), MoE (hm:1065-1162) and the HF cache classes are out of scope.
These unquantised modules are the FP teacher of the calibration loops (alg:471-479, 674-688); the quantised hot path
swaps them for mobilequant_b200.quantization.qmodule.Q* modules (CUDA kernels) or compiles them into
mobilequant_b200.engine (integer forward).
"""
import math
from types import SimpleNamespace
import torch
import torch.nn as nn
from .hf_config import HFConfig
from .ops import L2Norm, ElementwiseAdd, ElementwiseMul, FMatMul


class HFRMSNorm(nn.Module):
    """hm:162-198.  With l2norm_as_rmsnorm the layer is sqrt(d) * x/max(||x||,1e-12) * w (eps unused)."""

    def __init__(self, dim, eps=1e-6, device=None, dtype=None, bias=None, l2norm_as_rmsnorm=False):
        super().__init__()
        self.eps = eps
        self.alpha = math.sqrt(dim)
        self.weight = nn.Parameter(torch.ones(dim, device=device, dtype=dtype))
        self.bias = None if bias is None else nn.Parameter(torch.zeros(dim, device=device, dtype=dtype))
        self.l2norm_as_rmsnorm = l2norm_as_rmsnorm
        if l2norm_as_rmsnorm:
            self.l2norm = L2Norm()
        self.elementwisemul = ElementwiseMul()
        torch.nn.init.normal_(self.weight)          # hm:179-182

    def _norm(self, x):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps)

    def forward_impl(self, x, weight, bias):
        if self.l2norm_as_rmsnorm:
            output = self.alpha * self.l2norm(x)
        else:
            output = self._norm(x.float()).type_as(x)
        output = self.elementwisemul(weight, output)
        if bias is not None:
            output = output + bias
        return output

    def forward(self, x):
        return self.forward_impl(x, self.weight, self.bias)


def rope_cos_sin(position_ids, dim, base, device, dtype):
    """hm:308-318 (Gemma-style call, position_ids given).  Returns cos, sin of shape [B, T, dim]."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64, device=device).float() / dim))
    inv_freq_expanded = inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
    freqs = (inv_freq_expanded @ position_ids[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin):
    """hm:338-367 with unsqueeze_dim=1."""
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def repeat_kv(x, n_rep):
    b, h, t, d = x.shape
    if n_rep == 1:
        return x
    return x[:, :, None, :, :].expand(b, h, n_rep, t, d).reshape(b, h * n_rep, t, d)


_ACT = {"silu": nn.SiLU, "gelu": nn.GELU, "gelu_new": lambda: nn.GELU(approximate="tanh"),
        "gelu_pytorch_tanh": lambda: nn.GELU(approximate="tanh")}


class HFAttention(nn.Module):
    """Eager attention, hm:382-549."""

    def __init__(self, config, layer_idx=None):
        super().__init__()
        self.config = config
        self.layer_idx = layer_idx
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = config.head_dim if config.head_dim is not None else self.hidden_size // self.num_heads
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.rope_theta = config.rope_theta
        self.rotary_dim = int(config.partial_rotary_factor * self.head_dim)
        b = config.attention_bias
        self.q_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_dim, bias=b)
        self.k_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=b)
        self.v_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=b)
        self.o_proj = nn.Linear(self.num_heads * self.head_dim, self.hidden_size, bias=b and not config.use_qkv_bias_only)
        self.qk_bmm = FMatMul()
        self.pv_bmm = FMatMul()

    def forward(self, hidden_states, attention_mask=None, position_ids=None, **kwargs):
        bsz, q_len, _ = hidden_states.size()
        fused = getattr(self.qk_bmm, "fused_attention", None)       # Q* modules: the block's attention half in four fused kernels
        if fused is not None:
            out = fused(self, hidden_states, attention_mask, position_ids)
            if out is not None:
                return self.o_proj(out), None, None
        q = self.q_proj(hidden_states).view(bsz, q_len, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.k_proj(hidden_states).view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        v = self.v_proj(hidden_states).view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        cos, sin = rope_cos_sin(position_ids, self.rotary_dim, self.rope_theta, hidden_states.device, hidden_states.dtype)
        if self.rotary_dim == self.head_dim:
            q, k = apply_rotary_pos_emb(q, k, cos, sin)
        else:                                        # partial rotary (StableLM), hm:489-501
            qr, kr = apply_rotary_pos_emb(q[..., : self.rotary_dim], k[..., : self.rotary_dim], cos, sin)
            q = torch.cat((qr, q[..., self.rotary_dim:]), dim=-1)
            k = torch.cat((kr, k[..., self.rotary_dim:]), dim=-1)
        k = repeat_kv(k, self.num_key_value_groups)
        v = repeat_kv(v, self.num_key_value_groups)
        fused = getattr(self.qk_bmm, "fused_probs", None)          # QMatMul pair: one kernel for the element-wise attention core
        attn = fused(q, k.transpose(2, 3), self.head_dim, attention_mask, self.pv_bmm) if fused is not None else None
        if attn is not None:
            out = self.pv_bmm(attn, v, input_quantized=True)
        else:
            attn = self.qk_bmm(q, k.transpose(2, 3)) / math.sqrt(self.head_dim)
            if attention_mask is not None:
                attn = attn + attention_mask
            attn = nn.functional.softmax(attn, dim=-1, dtype=torch.float32).to(q.dtype)
            out = self.pv_bmm(attn, v)
        out = out.transpose(1, 2).contiguous().view(bsz, q_len, -1)
        return self.o_proj(out), None, None


class HFMLP(nn.Module):
    """hm:1042-1062."""

    def __init__(self, config):
        super().__init__()
        self.num_linears_per_mlp = config.num_linears_per_mlp
        self.w1 = nn.Linear(config.hidden_size, config.intermediate_size, bias=config.mlp_bias)
        self.w2 = nn.Linear(config.intermediate_size, config.hidden_size, bias=config.mlp_bias)
        if self.num_linears_per_mlp == 3:
            self.w3 = nn.Linear(config.hidden_size, config.intermediate_size, bias=config.mlp_bias)
            self.elementwisemul = ElementwiseMul()
        self.act_fn = _ACT[config.hidden_act]()

    def forward(self, x):
        fused = getattr(self.act_fn, "fused_mlp", None) if self.num_linears_per_mlp == 3 else None
        if fused is not None:                               # Q* modules: one GEMM for w1 | w3 + one kernel for the element-wise core
            h = fused(x, self.w1, self.w3, self.w2)
            if h is not None:
                return self.w2(h, input_quantized=True)
        h = self.act_fn(self.w1(x))
        if self.num_linears_per_mlp == 3:
            h = self.elementwisemul(h, self.w3(x))
        return self.w2(h)


def _make_norm(config, in_layer):
    if config.norm_class.lower() == "layernorm":
        return nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
    if config.norm_class.lower() == "rmsnorm":
        # hm:1198-1203: decoder-layer norms get l2norm_as_rmsnorm, the final model.norm does not (hm:1441)
        return HFRMSNorm(config.hidden_size, eps=config.layer_norm_eps,
                         l2norm_as_rmsnorm=config.l2norm_as_rmsnorm if in_layer else False)
    raise NotImplementedError(config.norm_class)


class HFDecoderLayer(nn.Module):
    """hm:1165-1283."""

    def __init__(self, config, layer_idx):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = HFAttention(config, layer_idx)
        self.shared_attention_norm = config.shared_attention_norm
        self.parallel_residual = config.parallel_residual
        self.mlp = HFMLP(config)
        self.input_layernorm = _make_norm(config, True)
        if not self.shared_attention_norm:
            self.post_attention_layernorm = _make_norm(config, True)
        self.resid_add_1 = ElementwiseAdd()
        self.resid_add_2 = ElementwiseAdd()

    def forward(self, hidden_states, attention_mask=None, position_ids=None, **kwargs):
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        attn_out, _, _ = self.self_attn(hidden_states=hidden_states, attention_mask=attention_mask,
                                        position_ids=position_ids)
        residual = self.resid_add_1(residual, attn_out)
        if not self.parallel_residual:
            hidden_states = residual
        if not self.shared_attention_norm:
            hidden_states = self.post_attention_layernorm(hidden_states)
        hidden_states = self.resid_add_2(residual, self.mlp(hidden_states))
        return (hidden_states,)


def causal_mask_4d(bsz, seq_len, dtype, device):
    """What transformers' _prepare_4d_causal_attention_mask(None, ...) returns (hm:1548-1555): finfo.min above the
    diagonal, 0 elsewhere, shape [B, 1, T, T]."""
    m = torch.full((seq_len, seq_len), torch.finfo(dtype).min, dtype=dtype, device=device)
    m = torch.triu(m, diagonal=1)
    return m[None, None].expand(bsz, 1, seq_len, seq_len)


class HFModel(nn.Module):
    """hm:1421-1627."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, config.pad_token_id)
        self.layers = nn.ModuleList([HFDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = _make_norm(config, False)

    def forward(self, input_ids=None, inputs_embeds=None, position_ids=None, **kwargs):
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        bsz, seq_len = inputs_embeds.shape[:2]
        if position_ids is None:
            position_ids = torch.arange(seq_len, device=inputs_embeds.device).unsqueeze(0)
        mask = causal_mask_4d(bsz, seq_len, inputs_embeds.dtype, inputs_embeds.device)
        h = inputs_embeds
        if self.config.normalize_embed:
            h = h * (self.config.hidden_size ** 0.5)
        for layer in self.layers:
            h = layer(h, attention_mask=mask, position_ids=position_ids)[0]
        h = self.norm(h)
        return SimpleNamespace(last_hidden_state=h)


class HFForCausalLM(nn.Module):
    """hm:1675-1899 (forward only; generation utilities are out of scope)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = HFModel(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=config.mlp_bias)        # hm:1682
        self.apply(self._init_weights)
        if config.tie_word_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight

    def _init_weights(self, module):                      # hm:1318-1327
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()

    def forward(self, input_ids=None, inputs_embeds=None, position_ids=None, **kwargs):
        out = self.model(input_ids=input_ids, inputs_embeds=inputs_embeds, position_ids=position_ids)
        logits = self.lm_head(out.last_hidden_state).float()
        return SimpleNamespace(logits=logits, last_hidden_state=out.last_hidden_state)

    def save_pretrained(self, out_dir, safe_serialization=False):
        """Same artefact names as PreTrainedModel.save_pretrained(safe_serialization=False) (ptq/mobilequant.py:245)."""
        import os
        self.config.save_pretrained(out_dir)
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        torch.save(sd, os.path.join(out_dir, "pytorch_model.bin"))

    @classmethod
    def from_pretrained(cls, path, **kw):
        import os
        cfg = HFConfig.from_pretrained(path)
        for k, v in kw.items():
            setattr(cfg, k, v)
        m = cls(cfg)
        sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        m.load_state_dict(sd, strict=False)
        return m
