"""Fold a calibrated HF-layout checkpoint into the on-device "sim" layout (reference: device/convert_sim.py:138-178) --
the artifact `sim_<name>.pth` that the unchanged AIMET / QNN toolchain of the reference consumes.

    norm weights            * sqrt(n_embd)            (the sim model's FRMSNorm is the plain L2 form)         :163-164
    embed_tokens (Gemma)    * sqrt(n_embd)                                                                      :166-167
    lm_head, and w2 when impl_sym_pch_as_slinear:  per-row scale = max|W| ; weight / scale  -> .scale, .linear.weight  :146-162
    q_proj.weight           / sqrt(head_dim)          (attention scaling fused into the projection)           :172-174

The per-row maxima are one pass of mq_minmax_2d on the device; everything else is bookkeeping on the state dict.  The AIMET
`.encodings` writer (device/utils.py:278-303, device/calibrate.py:279-302) needs the ONNX node names AIMET assigns and is
out of reach offline (parity unpinned at that boundary); the learned ranges are exported as act_dict.json."""
import math, os
from collections import OrderedDict
import torch
from .. import kernels as K

# mobilellm/model/sim_model.py:40-45
SIM_CONFIGS = {
    "llama-1.1b-mobilequant-w4a8-s1024-e60-sym-hf": dict(n_layer=22, n_head=32, n_kv_head=4, head_dim=64, n_embd=2048, intermediate_size=5632, vocab_size=32000, block_size=1024, norm_eps=1e-5),
    "llama-1.1b-mobilequant-w8a8-s1024-e60-hf": dict(n_layer=22, n_head=32, n_kv_head=4, head_dim=64, n_embd=2048, intermediate_size=5632, vocab_size=32000, block_size=1024, norm_eps=1e-5, impl_sym_pch_as_slinear=True),
    "gemma-2b-mobilequant-w4a8-s1024-e60-sym-hf": dict(n_layer=18, n_head=8, n_kv_head=1, head_dim=256, n_embd=2048, intermediate_size=16384, vocab_size=256000, block_size=1024, norm_eps=1e-6, act_fn="gelu"),
    "gemma-2b-mobilequant-w8a8-s1024-e60-hf": dict(n_layer=18, n_head=8, n_kv_head=1, head_dim=256, n_embd=2048, intermediate_size=16384, vocab_size=256000, block_size=1024, norm_eps=1e-6, act_fn="gelu", impl_sym_pch_as_slinear=True),
}


def sim_state_keys(n_layer, impl_sym_pch_as_slinear=False, attention_bias=False, mlp_bias=False):
    """Parameter names of SimModel (sim_model.py:88-101, 223-260): what `elif k in sim_state` (:168) lets through."""
    keys = ["embed_tokens.weight", "norm.weight", "lm_head.scale", "lm_head.linear.weight"]
    if mlp_bias:
        keys.append("lm_head.linear.bias")
    for i in range(n_layer):
        p = f"layers.{i}."
        keys += [p + "input_layernorm.weight", p + "post_attention_layernorm.weight"]
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            keys.append(p + f"self_attn.{n}.weight")
            if attention_bias:
                keys.append(p + f"self_attn.{n}.bias")
        for n in ("w1", "w3"):
            keys.append(p + f"mlp.{n}.weight")
            if mlp_bias:
                keys.append(p + f"mlp.{n}.bias")
        if impl_sym_pch_as_slinear:
            keys += [p + "mlp.w2.scale", p + "mlp.w2.linear.weight"]
            if mlp_bias:
                keys.append(p + "mlp.w2.linear.bias")
        else:
            keys.append(p + "mlp.w2.weight")
            if mlp_bias:
                keys.append(p + "mlp.w2.bias")
    return set(keys)


def _row_absmax(w):
    """max_k |W[n, k]| per row on the device (== torch.max(torch.abs(W), dim=1)[0], convert_sim.py:148,156)."""
    w = w.float().contiguous()
    if not w.is_cuda:
        raise RuntimeError("convert_sim folds checkpoints on a CUDA device (no CPU fallback)")
    mn, mx = K.minmax_2d(w, per_row=True)
    return torch.maximum(mn.abs(), mx.abs())


@torch.no_grad()
def convert_state_dict(hf_state, n_embd, head_dim, n_layer, impl_sym_pch_as_slinear=False, is_gemma=False, device="cuda"):
    """hf_state: state dict of the calibrated HFForCausalLM.  Returns the OrderedDict saved as sim_<name>.pth (CPU fp32)."""
    sim_keys = sim_state_keys(n_layer, impl_sym_pch_as_slinear)
    new_state = OrderedDict()
    for k in list(hf_state.keys()):
        v = hf_state[k].detach().to(torch.float32)
        k = k.replace("model.", "")
        split = (impl_sym_pch_as_slinear and "w2.weight" in k) or ("lm_head" in k)
        if split:
            vd = v.to(device)
            scale = _row_absmax(vd)
            base = k.replace("w2.weight", "w2") if "w2.weight" in k else k.replace("lm_head.weight", "lm_head")
            new_state[base + ".scale"] = scale.cpu()
            new_state[base + ".linear.weight"] = (vd / scale.unsqueeze(1)).cpu()
        if "norm.weight" in k:
            new_state[k] = (v * math.sqrt(n_embd)).cpu()
        elif is_gemma and "embed_tokens" in k:
            new_state[k] = (v * math.sqrt(n_embd)).cpu()
        elif k in sim_keys:
            new_state[k] = v.cpu()
    missing = sim_keys - set(new_state.keys())
    if missing:                                             # load_state_dict(strict=True) of the reference (:169)
        raise KeyError(f"checkpoint lacks {sorted(missing)[:4]} ...")
    for i in range(n_layer):                                # :172-174
        new_state[f"layers.{i}.self_attn.q_proj.weight"] = new_state[f"layers.{i}.self_attn.q_proj.weight"] / math.sqrt(head_dim)
    return new_state


def convert(model, model_name, output_dir):
    """model: calibrated HFForCausalLM (tied lm_head already materialised).  Writes <output_dir>/sim_<model_name>.pth."""
    c = SIM_CONFIGS[model_name]
    dev = next(model.parameters()).device
    out = convert_state_dict(model.state_dict(), c["n_embd"], c["n_embd"] // c["n_head"], c["n_layer"], c.get("impl_sym_pch_as_slinear", False),
                             "gemma" in model_name.lower(), dev)
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, f"sim_{model_name}.pth")
    torch.save(out, path)
    return path
