"""Per-channel activation scales (abs-max) and shifts ((max+min)/2, EMA) of every Linear / norm input and output --
the statistics the SmoothQuant initialiser consumes (reference: ptq/generate_act_scale_shift.py:42-149).

The reference reduces each hooked tensor on the host (`.abs().max(dim=0)` then `.float().cpu()`, :49-51 / :104-106); here
the per-column min and max of a tensor are one pass of mq_minmax_2d on the device and the running statistics stay there
until the single copy at the end.  abs-max == max(|col min|, |col max|) exactly; the shift EMA (:108-111) is the same three
fp32 operations, so both dictionaries are bit-identical to the reference's.

    python -m mobilequant_b200.ptq.generate_act_scale_shift --hf_path <dir> [--num_samples N] [--seq_len T]
writes act_scales.pth / act_shifts.pth next to the checkpoint (generate_act_scale_shift.py:176-183)."""
import argparse, os
from functools import partial
import torch
import torch.nn as nn
from .. import kernels as K
from ..model.hf_model import HFForCausalLM, HFRMSNorm

HOOKED = (nn.Linear, nn.LayerNorm, HFRMSNorm)            # generate_act_scale_shift.py:66, :121


def _col_minmax(t):
    t2 = t.detach().float().reshape(-1, t.shape[-1]).contiguous()
    return K.minmax_2d(t2, per_row=False)                 # (min[C], max[C]) on the device


def _collect(model, samples, update):
    model.eval()
    device = next(model.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("activation statistics are collected on a CUDA device (no CPU fallback)")

    def hook(m, x, y, name):
        update(name, "input", x[0] if isinstance(x, tuple) else x)
        update(name, "output", y[0] if isinstance(y, tuple) else y)

    hooks = [m.register_forward_hook(partial(hook, name=n)) for n, m in model.named_modules() if isinstance(m, HOOKED)]
    for ids in samples:
        model(ids.to(device))
    for h in hooks:
        h.remove()


@torch.no_grad()
def get_act_scales(model, samples):
    """{"<module>_input" / "<module>_output": float32 CPU tensor [C]} = running per-channel abs-max (:46-55)."""
    stats = {}

    def update(name, field, t):
        mn, mx = _col_minmax(t)
        cur = torch.maximum(mn.abs(), mx.abs())
        key = f"{name}_{field}"
        stats[key] = torch.maximum(stats[key], cur) if key in stats else cur

    _collect(model, samples, update)
    return {k: v.cpu() for k, v in stats.items()}


@torch.no_grad()
def get_act_shifts(model, samples):
    """{"<module>_input" / "<module>_output": float32 CPU tensor [C]} = EMA(0.99) of (max + min) / 2 per channel (:101-111)."""
    stats = {}

    def update(name, field, t):
        mn, mx = _col_minmax(t)
        cur = (mx + mn) / 2
        key = f"{name}_{field}"
        stats[key] = 0.99 * stats[key] + 0.01 * cur if key in stats else cur

    _collect(model, samples, update)
    return {k: v.cpu() for k, v in stats.items()}


def main(argv=None):
    from .generate_act_range import random_samples
    p = argparse.ArgumentParser()
    p.add_argument("--hf_path", type=str, required=True)
    p.add_argument("--seq_len", type=int, default=4096)
    p.add_argument("--num_samples", type=int, default=512)
    p.add_argument("--use_rand_samples", default=False, action="store_true")
    p.add_argument("--output_dir", default=None, type=str)
    args = p.parse_args(argv)
    out_dir = args.output_dir or args.hf_path
    torch.manual_seed(1337)
    model = HFForCausalLM.from_pretrained(args.hf_path, use_matmul_as_module=True, l2norm_as_rmsnorm=True).float().cuda()
    # no tokenizer / dataset access offline: the calibration set is the reference's own random-id convention (:83-85)
    samples = random_samples(args.num_samples, args.seq_len, model.config.vocab_size, model.config.bos_token_id or 1)
    os.makedirs(out_dir, exist_ok=True)
    torch.save(get_act_scales(model, samples), os.path.join(out_dir, "act_scales.pth"))
    torch.save(get_act_shifts(model, samples), os.path.join(out_dir, "act_shifts.pth"))


if __name__ == "__main__":
    main()
