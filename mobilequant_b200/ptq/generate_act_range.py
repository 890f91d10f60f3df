"""Per-tensor / per-channel activation-range calibration (mirror of ptq/generate_act_range.py:49-158).

Same hooks on the same module types, same act_dict.json layout -- but the running [min, max] of every hooked tensor
lives in one device buffer updated by the mq_minmax / mq_minmax_2d kernels (no `.min().item()` host sync per tensor,
generate_act_range.py:65; ~1150 syncs per sample in the reference) and is read back once at the end.  With
torch.distributed initialised, samples are sharded round-robin over ranks and the packed ranges are combined with a
single all-reduce (MAX over [-min, max]): the result is bit-identical to the single-GPU run.

    python -m mobilequant_b200.ptq.generate_act_range --hf_path <dir> [--use_rand_samples] [--per_channel]
"""
import argparse, os
from collections import defaultdict
from functools import partial
import torch
import torch.nn as nn
import torch.distributed as dist
from .. import kernels as K
from ..model.hf_model import HFRMSNorm, HFForCausalLM
from ..model.ops import FMatMul
from ..utils.io import json_save

HOOKED = (nn.Linear, nn.SiLU, nn.Softmax, nn.GELU, nn.LayerNorm, HFRMSNorm, FMatMul)   # generate_act_range.py:93


@torch.no_grad()
def get_act_range(model, samples, per_channel=False):
    """samples: list of LongTensor[1, T] on any device.  Returns the act_dict ({name: {field: [min, max]}} or, with
    per_channel, {name: {field: Tensor[2, C]}})."""
    model.eval()
    device = next(model.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("activation-range calibration runs on a CUDA device (no CPU fallback)")
    stats = {}          # (name, field) -> float32[2] (per tensor) or (min[C], max[C])
    order = []

    def update(name, field, t):
        key = (name, field)
        t = t.detach().float()
        if per_channel:
            t2 = t.reshape(-1, t.shape[-1])
            if key not in stats:
                stats[key] = K.minmax_2d(t2, per_row=False); order.append(key)
            else:
                K.minmax_2d(t2, False, stats[key][0], stats[key][1], accumulate=True)
        else:
            if key not in stats:
                stats[key] = K.minmax(t); order.append(key)
            else:
                K.minmax(t, stats[key], accumulate=True)

    def hook(m, xx, yy, name):
        update(name, "input", xx[0] if isinstance(xx, tuple) else xx)
        update(name, "output", yy[0] if isinstance(yy, tuple) else yy)
        if isinstance(m, FMatMul):
            update(name, "input2", xx[1])

    hooks = [m.register_forward_hook(partial(hook, name=n)) for n, m in model.named_modules() if isinstance(m, HOOKED)]
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
    for i, ids in enumerate(samples):
        if i % world == rank:
            model(ids.to(device))
    for h in hooks:
        h.remove()
    if world > 1 and len(samples) < world:
        raise ValueError(f"sample-sharded act-range calibration needs at least one sample per rank ({len(samples)} samples, {world} ranks)")
    if world > 1 and not per_channel:
        from ..utils.dist import allreduce_ranges
        packed = allreduce_ranges(torch.stack([stats[k] for k in order]))   # [n, 2]; one exchange: max(-min), max(max)
        for k, row in zip(order, packed):
            stats[k] = row
    act_dict = defaultdict(dict)
    if per_channel:
        for (name, field) in order:
            mn, mx = stats[(name, field)]
            if world > 1:
                neg = -mn
                dist.all_reduce(neg, op=dist.ReduceOp.MAX); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                mn = -neg
            act_dict[name][field] = torch.stack((mn, mx), dim=0).cpu()
    else:
        host = torch.stack([stats[k] for k in order]).cpu().tolist()    # the single D2H
        for (name, field), (mn, mx) in zip(order, host):
            act_dict[name][field] = [mn, mx]
    return dict(act_dict)


def random_samples(num_samples, seq_len, vocab_size, bos_token_id=1, seed=1337):
    """generate_act_range.py:106-108 (--use_rand_samples): uniform ids in [bos+1, vocab-1)."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(bos_token_id + 1, vocab_size - 1, (1, seq_len), generator=g) for _ in range(num_samples)]


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--hf_path", type=str, required=True)
    p.add_argument("--seq_len", type=int, default=4096)
    p.add_argument("--num_samples", type=int, default=512)
    p.add_argument("--per_channel", default=False, action="store_true")
    p.add_argument("--use_rand_samples", default=False, action="store_true")
    p.add_argument("--samples_path", default=None, type=str,
                   help="torch-saved list of LongTensor [1, T] token ids (the reference tokenises the Pile validation split, "
                        "generate_act_range.py:97-104; there is no dataset / tokenizer access here)")
    p.add_argument("--output_dir", default=None, type=str)
    args = p.parse_args(argv)
    if args.samples_path is None and not args.use_rand_samples:
        raise FileNotFoundError("no calibration text available offline: pass --samples_path <token-id list> or --use_rand_samples "
                                "(uniformly random token ids, the reference's own convention for its extra samples)")
    out_dir = args.output_dir or args.hf_path
    torch.manual_seed(1337)
    model = HFForCausalLM.from_pretrained(args.hf_path, use_matmul_as_module=True, l2norm_as_rmsnorm=True).float().cuda()
    samples = []
    if args.samples_path is not None:
        samples += [t.view(1, -1)[:, :args.seq_len].long() for t in torch.load(args.samples_path, weights_only=False)][:args.num_samples]
    if args.use_rand_samples:     # generate_act_range.py:105-108: one random-id sample per text sample (all of them without text)
        samples += random_samples(len(samples) or args.num_samples, args.seq_len, model.config.vocab_size, model.config.bos_token_id or 1)
    act_dict = get_act_range(model, samples, args.per_channel)
    os.makedirs(out_dir, exist_ok=True)
    if not args.per_channel:
        json_save(os.path.join(out_dir, "act_dict.json"), act_dict)
    else:
        torch.save(act_dict, os.path.join(out_dir, "act_dict_per_channel.pth"))


if __name__ == "__main__":
    main()
