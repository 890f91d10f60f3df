"""SmoothQuant initialiser: closed-form per-channel migration of activation outliers into the weights
(reference: ptq/smoothquant.py:50-139), the step before the learned LET of ptq/mobilequant.py.

    s = act_scales^alpha / max_fc |W|_col^(1 - alpha)        ln.w /= s ; fc.W *= s          (smooth_ln_fcs, :50-77)
    v_proj.W /= s ; o_proj.W *= s   (MHA only)    w3.W /= s ; w2.W *= s                     (smooth_fc_fcs, :82-107)

Same names, arguments and in-place semantics as the reference; the per-column |W| maxima come from mq_minmax_2d, the
rest is a handful of elementwise fp32 operations on the device.

    python -m mobilequant_b200.ptq.smoothquant --hf_path <dir> [--alpha 0.5] [--act_scales_path p] [--output_dir d]"""
import argparse, os
import torch
import torch.nn as nn
from .. import kernels as K
from ..model.hf_model import HFForCausalLM, HFRMSNorm, HFDecoderLayer


def _weight_col_absmax(fcs):
    """max over the fcs of the per-input-channel |W| maxima, clamped at 1e-5 (:60-61, :89-90)."""
    cols = []
    for fc in fcs:
        w = fc.weight.detach().float().contiguous()
        if not w.is_cuda:
            raise RuntimeError("SmoothQuant initialisation runs on a CUDA device (no CPU fallback)")
        mn, mx = K.minmax_2d(w, per_row=False)
        cols.append(torch.maximum(mn.abs(), mx.abs()).unsqueeze(0))
    return torch.cat(cols, dim=0).max(dim=0)[0].clamp(min=1e-5)


@torch.no_grad()
def smooth_ln_fcs(ln, fcs, act_scales, alpha=0.5):
    if not isinstance(fcs, list):
        fcs = [fcs]
    assert isinstance(ln, (nn.LayerNorm, HFRMSNorm))
    for fc in fcs:
        assert isinstance(fc, nn.Linear)
        assert len(ln.weight.data) == fc.in_features == len(act_scales)
    device, dtype = fcs[0].weight.device, fcs[0].weight.dtype
    act_scales = act_scales.to(device=device, dtype=dtype)
    weight_scales = _weight_col_absmax(fcs).to(dtype)
    scales = (act_scales.pow(alpha) / weight_scales.pow(1 - alpha)).clamp(min=1e-5).to(device).to(dtype)
    ln.weight.div_(scales)
    if hasattr(ln, "bias") and ln.bias is not None:
        ln.bias.div_(scales)
    for fc in fcs:
        fc.weight.mul_(scales.view(1, -1))
    for p in list(ln.parameters()) + [q for fc in fcs for q in fc.parameters()]:
        assert torch.isnan(p).sum() == 0


@torch.no_grad()
def smooth_fc_fcs(fc1, fcs, act_scales, alpha=0.5):
    if not isinstance(fcs, list):
        fcs = [fcs]
    device, dtype = fcs[0].weight.device, fcs[0].weight.dtype
    act_scales = act_scales.to(device=device, dtype=dtype)
    weight_scales = _weight_col_absmax(fcs).to(dtype)
    scales = (act_scales.pow(alpha) / weight_scales.pow(1 - alpha)).clamp(min=1e-5).to(device).to(dtype)
    fc1.weight.div_(scales.view(-1, 1))
    if fc1.bias is not None:
        fc1.bias.div_(scales.view(-1))
    for fc in fcs:
        fc.weight.mul_(scales.view(1, -1))
    for p in list(fc1.parameters()) + [q for fc in fcs for q in fc.parameters()]:
        assert torch.isnan(p).sum() == 0


@torch.no_grad()
def smooth_lm(model, scales, alpha=0.5, original_smoothquant=False, original_omniquant=False):
    """ptq/smoothquant.py:110-139."""
    for name, module in model.named_modules():
        if isinstance(module, HFDecoderLayer):
            at, mlp = module.self_attn, module.mlp
            if model.config.shared_attention_norm:
                fcs = [at.q_proj, at.k_proj, at.v_proj, mlp.w1]
                if model.config.num_linears_per_mlp == 3:
                    fcs.append(mlp.w3)
                smooth_ln_fcs(module.input_layernorm, fcs, scales[name + ".self_attn.q_proj_input"], alpha)
            else:
                smooth_ln_fcs(module.input_layernorm, [at.q_proj, at.k_proj, at.v_proj], scales[name + ".self_attn.q_proj_input"], alpha)
                fcs = [mlp.w1]
                if model.config.num_linears_per_mlp == 3:
                    fcs.append(mlp.w3)
                smooth_ln_fcs(module.post_attention_layernorm, fcs, scales[name + ".mlp.w1_input"], alpha)
            if not original_smoothquant:
                if at.v_proj.weight.shape[0] == at.o_proj.weight.shape[1]:
                    smooth_fc_fcs(at.v_proj, at.o_proj, scales[name + ".self_attn.o_proj_input"], alpha)
                if not original_omniquant and model.config.num_linears_per_mlp == 3:
                    smooth_fc_fcs(mlp.w3, mlp.w2, scales[name + ".mlp.w2_input"], alpha)


def main(argv=None):
    from .generate_act_range import random_samples
    from .generate_act_scale_shift import get_act_scales
    p = argparse.ArgumentParser()
    p.add_argument("--hf_path", type=str, required=True)
    p.add_argument("--act_scales_path", type=str, default=None)
    p.add_argument("--alpha", type=float, default=0.5)
    p.add_argument("--seq_len", type=int, default=4096)
    p.add_argument("--num_samples", type=int, default=512)
    p.add_argument("--use_rand_samples", default=False, action="store_true")
    p.add_argument("--original_smoothquant", default=False, action="store_true")
    p.add_argument("--original_omniquant", default=False, action="store_true")
    p.add_argument("--output_dir", default=None, type=str)
    args = p.parse_args(argv)
    out_dir = args.output_dir or args.hf_path
    path = args.act_scales_path or os.path.join(args.hf_path, "act_scales.pth")
    torch.manual_seed(1337)
    model = HFForCausalLM.from_pretrained(args.hf_path, use_matmul_as_module=True, l2norm_as_rmsnorm=True).float().cuda()
    if os.path.exists(path):
        act_scales = torch.load(path)
    else:                                                   # the reference requires the file (:164); computed here when absent
        samples = random_samples(args.num_samples, args.seq_len, model.config.vocab_size, model.config.bos_token_id or 1)
        act_scales = get_act_scales(model, samples)
        torch.save(act_scales, path)
    smooth_lm(model, act_scales, args.alpha, args.original_smoothquant, args.original_omniquant)
    model.save_pretrained(out_dir, safe_serialization=False)


if __name__ == "__main__":
    main()
