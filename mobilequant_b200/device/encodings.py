"""AIMET-schema `.encodings` artefacts emitted from the calibrated ranges, and the packed integer weight export
(reference: device/utils.py:278-530 `update_encodings`, device/calibrate.py:255-302).

The on-device toolchain of the reference consumes
  * `<name>.encodings`     {"activation_encodings": {node: {"input"|"output": {"0"|"1": enc}}}, "param_encodings": {...}}
                           enc = {bitwidth, dtype "int", is_symmetric "False", max, min, offset, scale}
  * `<name>_kv_cache.encodings`   {"k_cache": enc, "v_cache": enc}: min / max over the layers of qk_bmm.input2 / pv_bmm.input2.
In the reference the node names come from AIMET's ONNX export and `update_encodings` overrides that file's entries with
the act_dict.json ranges.  Here the overriding is a TABLE (`block_rules`): one row per (ONNX node of a decoder block,
source module, source field, target field, target slot, q-scaling) -- the same rows the reference applies statement by
statement -- so the very same function serves both uses:
  update_encodings(ori, act_dict, ...)     drop-in for the reference's function on an AIMET-written encodings file;
  encodings_from_act_dict(act_dict, ...)   builds the node skeleton itself (canonical names: the reference's prefixes
                                           and counters) and fills it, no AIMET / ONNX needed.
The arithmetic of an entry is device/utils.py:278-284: scale = (max - min) / (2^b - 1), offset = int(min * qmax / (max - min)).
Weights: `export_quantized_weights` writes the integer codes the tcgen05 GEMMs consume (int8 one per byte, int4 packed two per
byte, low nibble first) with their per-channel / per-tensor scale + offset and the matching "param_encodings".
"""
import math
import torch


def encoding_from_min_max(fmin, fmax, bitwidth):
    """device/utils.py:278-284 (`update_encodings_from_min_max`)."""
    qmax = 2 ** int(bitwidth) - 1
    return {"max": fmax, "min": fmin, "scale": (fmax - fmin) / qmax, "offset": int((fmin * qmax) / (fmax - fmin))}


def _blank(bitwidth):
    return {"bitwidth": int(bitwidth), "dtype": "int", "is_symmetric": "False", "max": 0.0, "min": 0.0, "offset": 0, "scale": 0.0}


# --------------------------------------------------------------------------------------------------------------------
# One row per override of device/utils.py:303-530: (node, exact?, source module suffix, source field, target field, slot, q?)
#   node      name pattern of the ONNX node; `exact` False = unique prefix match (utils.py:287-293)
#   q?        True: the range is multiplied by q_proj_factor (attention scaling folded into q_proj, convert_sim.py:172-174)
# --------------------------------------------------------------------------------------------------------------------
def block_rules(i, impl_sym_pch_as_slinear=False, has_sigmoid=True):
    L, A, M = f"layers.{i}.", f"layers.{i}.self_attn.", f"layers.{i}.mlp."
    n_in, n_post = "input_layernorm", "post_attention_layernorm"
    qk, pv = "self_attn.qk_bmm", "self_attn.pv_bmm"
    R = []
    add = lambda node, exact, src, sf, tf, slot, q=False: R.append((node, exact, src, sf, tf, slot, q))
    for norm in (n_in, n_post):                                                                   # :305-330
        add(L + norm + ".module_normalize", False, norm, "input", "input", "0")
        add(L + norm + ".module_mul", False, norm, "output", "output", "0")
    for proj, q in (("q_proj", True), ("k_proj", False), ("v_proj", False)):                        # :333-353
        add(A + proj, False, n_in, "output", "input", "0")
        add(A + proj, False, "self_attn." + proj, "output", "output", "0", q)
    add(A + "o_proj", False, "self_attn.o_proj", "output", "output", "0")                         # :356-359
    add(A + "o_proj", False, pv, "output", "input", "0")
    for w in ("w1", "w3"):                                                                        # :362-372
        add(M + w, False, n_post, "output", "input", "0")
        add(M + w, False, "mlp." + w, "output", "output", "0")
    if has_sigmoid:                                                                               # :376-388 (QSiLU only)
        add(M + "act.sigmoid", False, "mlp.w1", "output", "input", "0")
        add(M + "act.sigmoid", False, "mlp.act_fn", "input2", "output", "0")
        add(M + "act.mul", False, "mlp.w1", "output", "input", "0")
        add(M + "act.mul", False, "mlp.act_fn", "input2", "input", "1")
        add(M + "act.mul", False, "mlp.act_fn", "output", "output", "0")
    if impl_sym_pch_as_slinear:                                                                   # :390-397
        add(M + "w2.linear", False, "mlp.w2", "input", "input", "0")
    else:
        add(M + "w2", False, "mlp.w2", "input", "input", "0")
        add(M + "w2", False, "mlp.w2", "output", "output", "0")
    mm = "module_matmul" if i == 0 else f"module_matmul_{2 * i}"                                   # :403-414
    add(mm, True, qk, "input", "input", "0", True)
    add(mm, True, qk, "input2", "input", "1")
    add(mm, True, qk, "output", "output", "0", True)
    mm = f"module_matmul_{2 * i + 1}"
    add(mm, True, pv, "input", "input", "0")
    add(mm, True, pv, "input2", "input", "1")
    add(mm, True, pv, "output", "output", "0")
    add(A + "softmax", False, pv, "input", "output", "0")                                         # :418-420
    add(f"module_add_{5 * i + 3}", True, "self_attn.o_proj", "output", "input", "1")              # :426-434
    add(f"module_add_{5 * i + 4}", True, "mlp.w2", "output", "input", "1")
    # reshape / transpose / concat nodes carry the range of the projection they move (:441-516)
    rs = [A + "module_reshape" if i == 0 else A + f"module_reshape_{6 * i}"] + [A + f"module_reshape_{6 * i + j}" for j in range(1, 6)]
    for node, (src, q) in zip(rs, (("self_attn.q_proj", True), ("self_attn.k_proj", False), ("self_attn.v_proj", False),
                                   ("self_attn.k_proj", False), ("self_attn.v_proj", False), (pv, False))):
        add(node, True, src, "output", "input", "0", q)
        add(node, True, src, "output", "output", "0", q)
    ts = [A + "module_transpose" if i == 0 else A + f"module_transpose_{5 * i}"] + [A + f"module_transpose_{5 * i + j}" for j in range(1, 5)]
    for node, (src, q) in zip(ts, (("self_attn.q_proj", True), ("self_attn.k_proj", False), ("self_attn.v_proj", False),
                                   ("self_attn.k_proj", False), (pv, False))):
        add(node, True, src, "output", "input", "0", q)
        add(node, True, src, "output", "output", "0", q)
    for node, src, q in (("module_cat" if i == 0 else f"module_cat_{2 * i}", "self_attn.q_proj", True),
                         (f"module_cat_{2 * i + 1}", "self_attn.k_proj", False)):
        add(node, True, src, "output", "input", "0", q)
        add(node, True, src, "output", "input", "1", q)
        add(node, True, src, "output", "output", "0", q)
    m = 8 if impl_sym_pch_as_slinear else 7                                                       # :520-549 (RoPE / gate multiplies)
    add(f"module_mul_{m * i + 1}", True, "self_attn.q_proj", "output", "input", "0", True)
    add(f"module_mul_{m * i + 2}", True, "self_attn.q_proj", "output", "input", "0", True)
    add(f"module_mul_{m * i + 3}", True, "self_attn.k_proj", "output", "input", "0")
    add(f"module_mul_{m * i + 4}", True, "self_attn.k_proj", "output", "input", "0")
    add(f"module_mul_{m * i + 6}", True, "mlp.act_fn", "output", "input", "0")
    add(f"module_mul_{m * i + 6}", True, "mlp.w3", "output", "input", "1")
    add(f"module_mul_{m * i + 6}", True, "mlp.w2", "input", "output", "0")
    if impl_sym_pch_as_slinear:
        add(f"module_mul_{m * i + 7}", True, "mlp.w2", "output", "output", "0")
    return R


def _resolve(act, node, exact):
    if exact:
        if node not in act:
            raise KeyError(f"encodings file has no node {node!r}")
        return node
    hits = [k for k in act if k.startswith(node)]
    if len(hits) != 1:                                                      # utils.py:292 asserts a unique match
        raise KeyError(f"prefix {node!r} matches {len(hits)} nodes")
    return hits[0]


def update_encodings(ori_encodings, new_act_dict, num_blocks, q_proj_factor, impl_sym_pch_as_slinear=False):
    """Drop-in for device/utils.py:296-560: override the activation encodings of an AIMET-written file with the calibrated
    act_dict.json ranges.  Returns (encodings, names of the nodes that were NOT overridden) -- the reference prints those."""
    act = ori_encodings["activation_encodings"]
    untouched = set(act.keys())
    for i in range(num_blocks):
        has_sigmoid = any(k.startswith(f"layers.{i}.mlp.act.sigmoid") for k in act)
        for node, exact, src, sf, tf, slot, q in block_rules(i, impl_sym_pch_as_slinear, has_sigmoid):
            name = _resolve(act, node, exact)
            entry = new_act_dict[f"model.layers.{i}.{src}"]
            if sf not in entry and sf == "input2" and src.endswith("act_fn"):
                fmin, fmax = 0.0, 1.0              # sigmoid output: the range QSiLU.set_scale_offset assumes when none was recorded (qm:731-734)
            else:
                fmin, fmax = entry[sf]
            f = q_proj_factor if q else 1.0
            enc = act[name][tf][slot]
            enc.update(encoding_from_min_max(fmin * f, fmax * f, enc["bitwidth"]))
            untouched.discard(name)
    ori_encodings["activation_encodings"] = act
    return ori_encodings, sorted(untouched)


def encodings_skeleton(qcfg, num_blocks, impl_sym_pch_as_slinear=False, has_sigmoid=True):
    """The nodes `update_encodings` addresses, under canonical names, with the bitwidth of the quantizer each slot is filled
    from (default_qcfg.json)."""
    act = {}
    for i in range(num_blocks):
        for node, exact, src, sf, tf, slot, q in block_rules(i, impl_sym_pch_as_slinear, has_sigmoid):
            bits = int(qcfg[f"model.layers.{i}.{src}"][sf]["bitwidth"])
            act.setdefault(node, {}).setdefault(tf, {})[slot] = _blank(bits)
    return {"activation_encodings": act, "param_encodings": {}}


def encodings_from_act_dict(act_dict, qcfg, num_blocks, head_dim, impl_sym_pch_as_slinear=False):
    """`.encodings` content straight from the calibrated artefacts (no AIMET, no ONNX).  q_proj_factor = 1 / sqrt(head_dim):
    the sim model folds the attention scaling into q_proj (device/convert_sim.py:172-174, device/calibrate.py)."""
    has_sigmoid = "input2" in qcfg["model.layers.0.mlp.act_fn"]
    enc = encodings_skeleton(qcfg, num_blocks, impl_sym_pch_as_slinear, has_sigmoid)
    enc, _ = update_encodings(enc, act_dict, num_blocks, 1.0 / math.sqrt(head_dim), impl_sym_pch_as_slinear)
    return enc


def kv_cache_encodings(act_dict, num_blocks, bitwidth=8):
    """device/calibrate.py:275-285: one encoding for all key caches and one for all value caches -- min / max over the layers of
    the matmul operands that read them (qk_bmm.input2, pv_bmm.input2)."""
    out = {}
    for tag, mod in (("k_cache", "qk_bmm"), ("v_cache", "pv_bmm")):
        rng = [act_dict[f"model.layers.{i}.self_attn.{mod}"]["input2"] for i in range(num_blocks)]
        mn, mx = min(r[0] for r in rng), max(r[1] for r in rng)
        scale = (mx - mn) / (2 ** bitwidth - 1)
        out[tag] = {"bitwidth": bitwidth, "dtype": "int", "is_symmetric": "False", "max": mx, "min": mn, "offset": int(mn / scale), "scale": scale}
    return out


@torch.no_grad()
def export_quantized_weights(model, qcfg, pack4=True):
    """Integer weight tensors of every QLinear-to-be of a calibrated (fused) float model, as the tcgen05 GEMMs consume them:
    {name: {"codes": int8 / uint8 [N, K] (4-bit: uint8 [N, K/2], two codes per byte, low nibble first), "scale": [N] or [1],
            "offset": [N] or [1], "bitwidth", "is_symmetric", "is_per_channel"}} + the matching AIMET "param_encodings" rows.
    One fused pass per matrix on the device (mq_wprep_fwd: min/max, scale/offset of qm:40-61, codes, packing)."""
    from .. import kernels as K
    weights, params = {}, {}
    for name, mod in model.named_modules():
        if not isinstance(mod, torch.nn.Linear) or name not in qcfg or "weight" not in qcfg[name]:
            continue
        c = qcfg[name]["weight"]
        bits, sym, pc = int(c["bitwidth"]), c["is_symmetric"] in ("True", "true"), c["is_per_channel"] in ("True", "true")
        if bits > 8:
            continue
        w = mod.weight.detach().float().contiguous()
        if not w.is_cuda:
            raise RuntimeError("export_quantized_weights runs on a CUDA device (no CPU fallback)")
        do_pack = pack4 and bits == 4 and w.shape[1] % 2 == 0 and (w.shape[0] * w.shape[1]) % 32 == 0
        out = K.wprep_fwd(w, bits, sym, pc, want_fq=False, want_codes=True, pack4=do_pack)
        codes = out["codes"].view(torch.uint8) if do_pack else out["codes"]        # packed bytes are raw nibble pairs (two's complement for symmetric)
        weights[name] = {"codes": codes.cpu(), "scale": out["scale"].cpu(), "offset": out["offset"].cpu(), "bitwidth": bits,
                         "is_symmetric": sym, "is_per_channel": pc, "packed": bool(do_pack), "shape": tuple(w.shape)}
        qmin, qmax = (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1) if sym else (0, 2 ** bits - 1)
        rows = []
        for s, o in zip(out["scale"].cpu().tolist(), out["offset"].cpu().tolist()):
            rows.append({"bitwidth": bits, "dtype": "int", "is_symmetric": str(sym), "scale": s, "offset": int(o) if not sym else qmin,
                         "min": (qmin - o) * s, "max": (qmax - o) * s})
        params[name + ".weight"] = rows
    return weights, params


def export_all(model, act_dict, qcfg, output_dir, name="model", impl_sym_pch_as_slinear=False, kv_cache_bitwidth=8, pack4=True):
    """Write <name>.encodings, <name>_kv_cache.encodings and <name>_qweights.pth for a calibrated (fused) float model."""
    import os
    from ..utils.io import json_save
    cfg = model.config
    hd = cfg.head_dim if cfg.head_dim is not None else cfg.hidden_size // cfg.num_attention_heads
    enc = encodings_from_act_dict(act_dict, qcfg, cfg.num_hidden_layers, hd, impl_sym_pch_as_slinear)
    weights, params = export_quantized_weights(model, qcfg, pack4)
    enc["param_encodings"] = params
    os.makedirs(output_dir, exist_ok=True)
    json_save(os.path.join(output_dir, f"{name}.encodings"), enc)
    json_save(os.path.join(output_dir, f"{name}_kv_cache.encodings"), kv_cache_encodings(act_dict, cfg.num_hidden_layers, kv_cache_bitwidth))
    torch.save(weights, os.path.join(output_dir, f"{name}_qweights.pth"))
    return enc, weights


def main(argv=None):
    """python -m mobilequant_b200.device.encodings --hf_path <calibrated output dir of ptq/mobilequant.py>"""
    import argparse, os
    from ..model.hf_model import HFForCausalLM
    from ..utils.io import json_load
    p = argparse.ArgumentParser()
    p.add_argument("--hf_path", required=True, help="directory holding the fused fp checkpoint, act_dict.json and default_qcfg.json")
    p.add_argument("--output_dir", default=None)
    p.add_argument("--name", default=None)
    p.add_argument("--impl_sym_pch_as_slinear", action="store_true")
    p.add_argument("--kv_cache_bitwidth", type=int, default=8)
    a = p.parse_args(argv)
    model = HFForCausalLM.from_pretrained(a.hf_path, use_matmul_as_module=True, l2norm_as_rmsnorm=True).float().cuda()
    act = json_load(os.path.join(a.hf_path, "act_dict.json"))
    qcfg = json_load(os.path.join(a.hf_path, "default_qcfg.json"))
    export_all(model, act, qcfg, a.output_dir or a.hf_path, a.name or os.path.basename(os.path.normpath(a.hf_path)),
               a.impl_sym_pch_as_slinear, a.kv_cache_bitwidth)


if __name__ == "__main__":
    main()
