"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the fake-quant decoder and of the calibration loops.

A functional restatement (plain dict of tensors in, tensors out; torch CPU fp32 + autograd) of
  * the eager decoder forward           mobilellm/model/hf_model.py:187-195, 338-367, 426-549, 1057-1062, 1208-1283
  * the Q* module forwards              mobilellm/quantization/qmodule.py:341-358, 453-466, 515-531, 625-642, 739-753, 790-799
  * the mixed-precision recipe          ptq/mobilequant.py:175-201
  * act-range calibration               ptq/generate_act_range.py:49-122
  * LET / LWC / LRL training loops      mobilellm/quantization/algorithm.py:187-233, 381-584, 587-787
Weights are addressed by the reference's state_dict names ("model.layers.0.self_attn.q_proj.weight", ...), quantizer
state by the reference's module paths, so fixtures generated from the real reference plug in unchanged.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this file.
"""
import math
from collections import OrderedDict
import torch
import torch.nn.functional as F
from . import fakequant_ref as fr

MODE_NONE, MODE_DIV, MODE_MUL = 0, 1, 2


# ---------------------------------------------------------------------------------------------------------------
# configuration helpers
# ---------------------------------------------------------------------------------------------------------------
def head_dim(cfg):
    return cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]


def default_recipe(cfg, w_bits=8, w_sym=False, w_per_channel=False, a_bits=8, softmax8_in=False, softmax8_out=False):
    """default_qcfg.json content after create_sim_qmodel (qm:835-865) + update_quant_cfg (ptq/mobilequant.py:175-201).
    Values are (bits, symmetric, per_channel); missing slot == no quantizer."""
    W, A, A16 = (w_bits, w_sym, w_per_channel), (a_bits, False, False), (16, False, False)
    rec = {}
    L = cfg["num_hidden_layers"]
    for i in range(L):
        p = f"model.layers.{i}."
        norms = ["input_layernorm"] + ([] if cfg.get("shared_attention_norm") else ["post_attention_layernorm"])
        for n in norms:
            rec[p + n] = dict(input=A16, weight=(16, False, False), output=A)
        for n in ("q_proj", "k_proj", "v_proj"):
            rec[p + "self_attn." + n] = dict(weight=W, output=A)
        rec[p + "self_attn.o_proj"] = dict(weight=W, output=A16)
        rec[p + "self_attn.qk_bmm"] = dict(input=A, input2=A, output=A if softmax8_in else A16)
        rec[p + "self_attn.pv_bmm"] = dict(input=A if softmax8_out else A16, input2=A, output=A)
        rec[p + "mlp.w1"] = dict(weight=W, output=A)
        if cfg.get("num_linears_per_mlp", 3) == 3:
            rec[p + "mlp.w3"] = dict(weight=W, output=A)
        rec[p + "mlp.w2"] = dict(input=A, weight=(w_bits, w_sym, True), output=A16)
        if cfg["hidden_act"] == "silu":
            rec[p + "mlp.act_fn"] = dict(input2=A, output=A)
        else:
            rec[p + "mlp.act_fn"] = dict(output=A)
    return rec


def recipe_from_qcfg_json(qcfg):
    """default_qcfg.json (string valued, qm:100-107) -> recipe."""
    t = ("True", "true")
    return {name: {slot: (int(c["bitwidth"]), c["is_symmetric"] in t, c["is_per_channel"] in t) for slot, c in slots.items()}
            for name, slots in qcfg.items()}


class QState:
    """scale/offset of every static activation quantizer (qm:216-245, set from act_dict.json via qm:965-970)."""

    def __init__(self, recipe, act_dict, learnable=False):
        self.recipe = recipe
        self.p = OrderedDict()          # "module.slot_quantizer.scale" -> tensor
        self.rng = {}
        for name, slots in recipe.items():
            for slot, (bits, sym, _pc) in slots.items():
                if slot == "weight" or bits > 16:
                    continue
                if slot == "input2" and slot not in act_dict.get(name, {}):
                    mn, mx = 0.0, 1.0                      # qm:731-734
                else:
                    mn, mx = act_dict[name][slot]
                s, o, qmin, qmax = fr.scale_offset_from_minmax(mn, mx, bits, sym)
                self.p[f"{name}.{slot}_quantizer.scale"] = s.clone().requires_grad_(learnable)
                self.p[f"{name}.{slot}_quantizer.offset"] = o.clone().requires_grad_(learnable)
                self.rng[f"{name}.{slot}"] = (qmin, qmax)

    def fq(self, name, slot, x):
        if slot not in self.recipe.get(name, {}) or self.recipe[name][slot][0] > 16:
            return x
        qmin, qmax = self.rng[f"{name}.{slot}"]
        return fr.fake_quant(x, self.p[f"{name}.{slot}_quantizer.scale"], self.p[f"{name}.{slot}_quantizer.offset"], qmin, qmax)

    def codes(self, name, slot, x):
        qmin, qmax = self.rng[f"{name}.{slot}"]
        return fr.quant_codes(x, self.p[f"{name}.{slot}_quantizer.scale"], self.p[f"{name}.{slot}_quantizer.offset"], qmin, qmax)

    def act_dict(self):
        """export_act_range, qm:908-937"""
        out = {}
        for name, slots in self.recipe.items():
            e = {}
            for slot, (bits, sym, _pc) in slots.items():
                if slot == "weight" or bits > 16:
                    continue
                mn, mx = fr.minmax_from_scale_offset(self.p[f"{name}.{slot}_quantizer.scale"].detach(),
                                                     self.p[f"{name}.{slot}_quantizer.offset"].detach(), bits, sym)
                e[slot] = [mn.item(), mx.item()]
            out[name] = e
        return out


# ---------------------------------------------------------------------------------------------------------------
# model pieces
# ---------------------------------------------------------------------------------------------------------------
def rope_cos_sin(position_ids, dim, base):
    """hm:308-318"""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    inv = inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
    freqs = (inv @ position_ids[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def causal_mask(bsz, T, dtype=torch.float32):
    m = torch.triu(torch.full((T, T), torch.finfo(dtype).min, dtype=dtype), diagonal=1)
    return m[None, None].expand(bsz, 1, T, T)


def rms_l2norm(x, weight, alpha, bias=None):
    """hm:187-195 with l2norm_as_rmsnorm: weight * (alpha * normalize(x))"""
    out = weight * (alpha * F.normalize(x, p=2, dim=-1, eps=1e-12))
    return out if bias is None else out + bias


class Block:
    """One decoder block; `let`/`lwc` are dicts of learnable tensors (None -> plain weights)."""

    def __init__(self, sd, i, cfg):
        self.sd, self.i, self.cfg = sd, i, cfg
        self.p = f"model.layers.{i}."
        self.let = None        # {"qkv_smooth_scale": ..., ...}
        self.lwc = None        # {"self_attn.q_proj.weight_quantizer.upbound_factor": ..., ...}
        self.use_shift = False
        self.original_omniquant = False
        self.sim = False       # True: the model has been rewritten by create_sim_qmodel (QSiLU/QGELU even when disabled)

    def w(self, n):
        return self.sd[self.p + n + ".weight"]

    def b(self, n):
        return self.sd.get(self.p + n + ".bias")

    # ---- LET plan: alg:195-220 -------------------------------------------------------------------------------
    def _let_of(self, n):
        """(col_fac, col_mode, row_fac, row_mode, bias') for module n under the current LET parameters."""
        cf = cm = rf = rm = None
        bias = self.b(n)
        if self.let is None:
            return None, 0, None, 0, bias
        L, cfg = self.let, self.cfg
        three = cfg.get("num_linears_per_mlp", 3) == 3
        sh = lambda k: L.get(f"{k}_smooth_shift")
        vo = self.w("self_attn.v_proj").shape[0] == self.w("self_attn.o_proj").shape[1]
        qk = self.w("self_attn.q_proj").shape[0] == self.w("self_attn.k_proj").shape[0]
        fc2 = three and not self.original_omniquant
        cm = rm = 0
        ln_key = {"input_layernorm": "qkv", "post_attention_layernorm": "fc1"}
        fc_key = {"self_attn.q_proj": "qkv", "self_attn.k_proj": "qkv", "self_attn.v_proj": "qkv", "mlp.w1": "fc1", "mlp.w3": "fc1"}
        if cfg.get("shared_attention_norm"):
            fc_key["mlp.w1"] = fc_key["mlp.w3"] = "qkv"
        if n in ln_key:
            k = ln_key[n]; s = L[f"{k}_smooth_scale"]
            cf, cm = s, MODE_DIV
            bias = (bias - sh(k)) / s if bias is not None else (-1 * sh(k)) / s                   # alg:56-59
        elif n in fc_key:
            k = fc_key[n]; s = L[f"{k}_smooth_scale"]
            cf, cm = s, MODE_MUL
            bias = bias + self.w(n) @ sh(k) if bias is not None else self.w(n) @ sh(k)            # alg:64-67
            if n == "self_attn.v_proj" and vo:
                rf, rm = L["out_smooth_scale"], MODE_DIV
                bias = (bias - sh("out")) / rf.view(-1)                                            # alg:76
            if n == "mlp.w3" and fc2:
                rf, rm = L["fc2_smooth_scale"], MODE_DIV
                bias = (bias - sh("fc2")) / rf.view(-1)
            if n == "self_attn.q_proj" and qk:
                rf, rm = L["qkt_smooth_scale"], MODE_DIV
                bias = bias / rf.view(-1)                                                          # alg:94
            if n == "self_attn.k_proj" and qk:
                rf, rm = L["qkt_smooth_scale"], MODE_MUL
                bias = bias * rf.view(-1)                                                          # alg:96
        elif n == "self_attn.o_proj" and vo:
            cf, cm = L["out_smooth_scale"], MODE_MUL
            bias = bias + self.w(n) @ sh("out") if bias is not None else self.w(n) @ sh("out")    # alg:83-86
        elif n == "mlp.w2" and fc2:
            cf, cm = L["fc2_smooth_scale"], MODE_MUL
            bias = bias + self.w(n) @ sh("fc2") if bias is not None else self.w(n) @ sh("fc2")
        return cf, cm, rf, rm, bias

    def qweight(self, n, qs, quant):
        """(fake-quantised weight, bias) of module n: LET (alg:60-96) then weight Quantizer (qm:262-290)."""
        w = self.w(n)
        cf, cm, rf, rm, bias = self._let_of(n)
        is_vec = w.dim() == 1
        w2 = w.view(1, -1) if is_vec else w
        wt = fr.let_weight(w2, cf, cm, rf, rm)
        spec = qs.recipe.get(self.p + n, {}).get("weight") if quant else None
        if spec is not None and spec[0] <= 16:
            bits, sym, pc = spec
            su = sl = None
            if self.lwc is not None:
                su = torch.sigmoid(self.lwc[n + ".weight_quantizer.upbound_factor"])
                sl = torch.sigmoid(self.lwc[n + ".weight_quantizer.lowbound_factor"])
            wt = fr.dynamic_fake_quant(wt, bits, sym, pc, su, sl)
        return (wt.view(-1) if is_vec else wt), bias

    def linear(self, n, x, qs, quant):
        w, bias = self.qweight(n, qs, quant)
        name = self.p + n
        if quant:
            x = qs.fq(name, "input", x)
        y = F.linear(x, w, bias)
        return qs.fq(name, "output", y) if quant else y

    def norm(self, n, x, qs, quant, hook=None):
        w, bias = self.qweight(n, qs, quant)
        name = self.p + n
        xin = x
        if quant:
            x = qs.fq(name, "input", x)
        if self.cfg.get("norm_class", "rmsnorm") == "layernorm":
            y = F.layer_norm(x, x.shape[-1:], weight=w, bias=bias, eps=self.cfg.get("layer_norm_eps", 1e-5))
        else:
            y = rms_l2norm(x, w, math.sqrt(x.shape[-1]), bias)
        if hook:
            hook(name, "input", xin); hook(name, "output", y)
        return qs.fq(name, "output", y) if quant else y

    def forward(self, h, qs=None, quant=False, mask=None, position_ids=None, hook=None):
        cfg = self.cfg
        B, T, H = h.shape
        nh, nkv, hd = cfg["num_attention_heads"], cfg["num_key_value_heads"], head_dim(cfg)
        p = self.p
        hk = (lambda n, f, t: hook(p + n, f, t)) if hook else None
        lin = lambda n, x: self._hooked(n, x, qs, quant, hk)
        residual = h
        x = self.norm("input_layernorm", h, qs, quant, hook)
        q = lin("self_attn.q_proj", x).view(B, T, nh, hd).transpose(1, 2)
        k = lin("self_attn.k_proj", x).view(B, T, nkv, hd).transpose(1, 2)
        v = lin("self_attn.v_proj", x).view(B, T, nkv, hd).transpose(1, 2)
        rot = int(cfg.get("partial_rotary_factor", 1.0) * hd)
        cos, sin = rope_cos_sin(position_ids, rot, cfg.get("rope_theta", 10000.0))
        cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
        rope = lambda t: (t * cos) + (rotate_half(t) * sin)
        if rot == hd:
            q, k = rope(q), rope(k)
        else:
            q = torch.cat((rope(q[..., :rot]), q[..., rot:]), dim=-1)
            k = torch.cat((rope(k[..., :rot]), k[..., rot:]), dim=-1)
        rep = nh // nkv
        if rep > 1:
            k = k[:, :, None].expand(B, nkv, rep, T, hd).reshape(B, nh, T, hd)
            v = v[:, :, None].expand(B, nkv, rep, T, hd).reshape(B, nh, T, hd)
        kt = k.transpose(2, 3)
        name = p + "self_attn.qk_bmm"
        s = torch.matmul(qs.fq(name, "input", q), qs.fq(name, "input2", kt)) if quant else torch.matmul(q, kt)
        if hk:
            hk("self_attn.qk_bmm", "input", q); hk("self_attn.qk_bmm", "input2", kt); hk("self_attn.qk_bmm", "output", s)
        if quant:
            s = qs.fq(name, "output", s)
        s = s / math.sqrt(hd)
        if mask is not None:
            s = s + mask
        pr = F.softmax(s, dim=-1, dtype=torch.float32)
        name = p + "self_attn.pv_bmm"
        o = torch.matmul(qs.fq(name, "input", pr), qs.fq(name, "input2", v)) if quant else torch.matmul(pr, v)
        if hk:
            hk("self_attn.pv_bmm", "input", pr); hk("self_attn.pv_bmm", "input2", v); hk("self_attn.pv_bmm", "output", o)
        if quant:
            o = qs.fq(name, "output", o)
        o = o.transpose(1, 2).contiguous().view(B, T, nh * hd)
        attn = lin("self_attn.o_proj", o)
        residual = residual + attn
        x = residual if not cfg.get("parallel_residual") else x
        if not cfg.get("shared_attention_norm"):
            x = self.norm("post_attention_layernorm", x, qs, quant, hook)
        g = lin("mlp.w1", x)
        name = p + "mlp.act_fn"
        if cfg["hidden_act"] == "silu":
            if quant:                                                    # QSiLU, qm:739-753
                a = g * qs.fq(name, "input2", torch.sigmoid(g))
            elif self.sim:                                               # QSiLU with its quantizers disabled (FP targets)
                a = g * torch.sigmoid(g)
            else:                                                        # nn.SiLU of the raw FP model, hm:1055
                a = F.silu(g)
        elif cfg["hidden_act"] == "gelu" or quant or self.sim:           # QGELU always uses erf-GELU, qm:790-799
            a = F.gelu(g)
        else:
            a = F.gelu(g, approximate="tanh")
        if hk:
            hk("mlp.act_fn", "input", g); hk("mlp.act_fn", "output", a)
        if quant:
            a = qs.fq(name, "output", a)
        if cfg.get("num_linears_per_mlp", 3) == 3:
            a = a * lin("mlp.w3", x)
        m = lin("mlp.w2", a)
        return residual + m

    def _hooked(self, n, x, qs, quant, hk):
        y_pre = None
        if hk is not None and not quant:
            w, bias = self.qweight(n, qs, False)
            y_pre = F.linear(x, w, bias)
            hk(n, "input", x); hk(n, "output", y_pre)
            return y_pre
        return self.linear(n, x, qs, quant)


def embed(sd, cfg, ids):
    h = F.embedding(ids, sd["model.embed_tokens.weight"])
    if cfg.get("normalize_embed"):
        h = h * (cfg["hidden_size"] ** 0.5)
    return h


def final_norm(sd, cfg, h):
    w = sd["model.norm.weight"]
    if cfg.get("norm_class", "rmsnorm") == "layernorm":
        return F.layer_norm(h, h.shape[-1:], weight=w, bias=sd.get("model.norm.bias"), eps=cfg.get("layer_norm_eps", 1e-5))
    eps = cfg.get("layer_norm_eps", 1e-5)                               # hm:184-185, 1441 (plain RMSNorm, eps used)
    return w * (h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps))


def model_forward(sd, cfg, ids, qs=None, quant=False, hook=None, blocks=None):
    """HFForCausalLM.forward: logits [B, T, V] and the final hidden state."""
    B, T = ids.shape
    h = embed(sd, cfg, ids)
    mask = causal_mask(B, T)
    pos = torch.arange(T).unsqueeze(0)
    blocks = blocks or [Block(sd, i, cfg) for i in range(cfg["num_hidden_layers"])]
    for blk in blocks:
        h = blk.forward(h, qs, quant, mask, pos, hook)
    hn = final_norm(sd, cfg, h)
    if hook:
        hook("model.norm", "input", h); hook("model.norm", "output", hn)
    head = sd.get("lm_head.weight", sd["model.embed_tokens.weight"])
    logits = F.linear(hn, head)
    if hook:
        hook("lm_head", "input", hn); hook("lm_head", "output", logits)
    return logits, h


# ---------------------------------------------------------------------------------------------------------------
# act-range calibration -- generate_act_range.py:49-122 (per-tensor)
# ---------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def act_range(sd, cfg, samples):
    act = {}

    def hook(name, field, t):
        mn, mx = t.min().item(), t.max().item()
        e = act.setdefault(name, {})
        e[field] = [mn, mx] if field not in e else [min(e[field][0], mn), max(e[field][1], mx)]

    for ids in samples:
        model_forward(sd, cfg, ids, hook=hook)
    return act


# ---------------------------------------------------------------------------------------------------------------
# calibration loops
# ---------------------------------------------------------------------------------------------------------------
def get_lr(max_lr, min_lr, it, warmup_iters, max_iters):
    if it < warmup_iters:
        return max_lr * it / warmup_iters
    if it > max_iters:
        return min_lr
    r = (it - warmup_iters) / (max_iters - warmup_iters)
    return min_lr + 0.5 * (1.0 + math.cos(math.pi * r)) * (max_lr - min_lr)


def init_let(blk, original_omniquant=False, e2e=False):
    """alg:484-496 / 692-706: ones / zeros, created in the reference's registration order."""
    cfg = blk.cfg
    L = OrderedDict()
    qk = blk.w("self_attn.q_proj").shape[0] == blk.w("self_attn.k_proj").shape[0]
    vo = blk.w("self_attn.v_proj").shape[0] == blk.w("self_attn.o_proj").shape[1]
    if qk:
        L["qkt_smooth_scale"] = torch.ones(blk.w("self_attn.q_proj").shape[0], requires_grad=True)
    order = [("qkv", cfg["hidden_size"])]
    if vo:
        order.append(("out", blk.w("self_attn.o_proj").shape[1]))
    order.append(("fc1", cfg["hidden_size"]))
    if cfg.get("num_linears_per_mlp", 3) == 3 and (e2e or not original_omniquant):
        order.append(("fc2", cfg["intermediate_size"]))
    for k, n in order:
        L[f"{k}_smooth_shift"] = torch.zeros(n, requires_grad=True)
        L[f"{k}_smooth_scale"] = torch.ones(n, requires_grad=True)
    return L


def init_lwc(blk, recipe):
    """enable_lwc, qm:133-151: 4.0, [N,1] per-channel or [1] per-tensor."""
    out = OrderedDict()
    for name, slots in recipe.items():
        if not name.startswith(blk.p) or "weight" not in slots:
            continue
        n = name[len(blk.p):]
        w = blk.w(n)
        shape = (w.shape[0], 1) if slots["weight"][2] else (1,)
        out[n + ".weight_quantizer.upbound_factor"] = (torch.ones(shape) * 4.0).requires_grad_(True)
        out[n + ".weight_quantizer.lowbound_factor"] = (torch.ones(shape) * 4.0).requires_grad_(True)
    return out


def truncate_(L, use_shift=False, thr=1e-2):
    """alg:27-42, 190-193 (in place, no grad)."""
    with torch.no_grad():
        for k, t in L.items():
            if ("smooth" if use_shift else "smooth_scale") in k:
                small = t.abs() < thr
                t[small] = t[small].sign() * thr


def fuse_block(blk, recipe):
    """smooth_lm_inplace, alg:147-184: LET folded into the weights (+ zero-shift bias buffers), then run_lwc clamp."""
    sd = blk.sd
    with torch.no_grad():
        names = [n[len(blk.p):-len(".weight")] for n in list(sd.keys()) if n.startswith(blk.p) and n.endswith(".weight")]
        new = {}
        for n in names:
            if blk.p + n not in recipe or "weight" not in recipe[blk.p + n]:
                continue
            w = blk.w(n)
            is_vec = w.dim() == 1
            cf, cm, rf, rm, bias = blk._let_of(n) if blk.let is not None else (None, 0, None, 0, blk.b(n))
            wt = fr.let_weight(w.view(1, -1) if is_vec else w, cf, cm, rf, rm)
            pc = recipe[blk.p + n]["weight"][2]
            mn, mx = fr.tensor_minmax(wt, pc)
            if blk.lwc is not None:
                mx = torch.sigmoid(blk.lwc[n + ".weight_quantizer.upbound_factor"]) * mx
                mn = torch.sigmoid(blk.lwc[n + ".weight_quantizer.lowbound_factor"]) * mn
            wt = wt.clamp(mn, mx)
            new[n] = (wt.view(-1) if is_vec else wt, bias)
        for n, (wt, bias) in new.items():
            sd[blk.p + n + ".weight"] = wt.detach().clone()
            if bias is not None:
                sd[blk.p + n + ".bias"] = bias.detach().clone()
    blk.let = None
    blk.lwc = None


def calibrate(sd, cfg, recipe, act_dict, embeds, mode="e2e", epochs=1, batch_size=1, let=True, lwc=True, lrl=True,
              let_lr=1e-3, lwc_lr=1e-2, lrl_lr=1e-6, let_min_lr=None, lwc_min_lr=None, lrl_min_lr=None, wd=0.0,
              warmup_epochs=0, original_omniquant=False, log=None):
    """omniquant (alg:381-584) / e2equant (alg:587-787) on CPU.  embeds: [nsamples, T, H] first-layer inputs.
    Returns dict(params={layer: OrderedDict}, act_dict=..., losses=[...], sd=fused state dict)."""
    let_min_lr = let_lr if let_min_lr is None else let_min_lr
    lwc_min_lr = lwc_lr if lwc_min_lr is None else lwc_min_lr
    lrl_min_lr = lrl_lr if lrl_min_lr is None else lrl_min_lr
    sd = {k: v.clone() for k, v in sd.items()}
    nsamples, T, H = embeds.shape
    L = cfg["num_hidden_layers"]
    blocks = [Block(sd, i, cfg) for i in range(L)]
    for b in blocks:
        b.original_omniquant = original_omniquant and mode != "e2e"
        b.sim = True
    qs = QState(recipe, act_dict, learnable=lrl)
    mask1 = causal_mask(1, T); maskb = causal_mask(batch_size, T)
    pos = torch.arange(T).unsqueeze(0)
    quant_inps = embeds.clone(); fp_inps = embeds.clone()
    steps = nsamples // batch_size
    max_iters, warm = epochs * steps, warmup_epochs * steps
    losses, params = [], {}

    def lrl_params(prefixes):
        return [t for k, t in qs.p.items() if any(k.startswith(p) for p in prefixes)]

    def train(blks, fwd):
        groups = [{"params": [t for b in blks for k, t in b.let.items() if "smooth_scale" in k] if let else [], "lr": let_lr},
                  {"params": [t for b in blks for t in b.lwc.values()] if lwc else [], "lr": lwc_lr}]
        if lrl or mode == "e2e":
            groups.append({"params": lrl_params([b.p for b in blks]) if lrl else [], "lr": lrl_lr})
        groups = [g for g in groups if len(g["params"]) > 0 or True]
        opt = torch.optim.AdamW(groups, weight_decay=wd)
        for ep in range(epochs):
            for j in range(steps):
                it = ep * steps + j
                opt.param_groups[0]["lr"] = get_lr(let_lr, let_min_lr, it, warm, max_iters)
                opt.param_groups[1]["lr"] = get_lr(lwc_lr, lwc_min_lr, it, warm, max_iters)
                if len(opt.param_groups) > 2:
                    opt.param_groups[2]["lr"] = get_lr(lrl_lr, lrl_min_lr, it, warm, max_iters)
                for b in blks:
                    if let:
                        truncate_(b.let)
                idx = j * batch_size
                out = fwd(quant_inps[idx:idx + batch_size])
                loss = F.mse_loss(out, fp_inps[idx:idx + batch_size])      # MSELoss(fp, quant), alg:533
                losses.append(loss.item())
                opt.zero_grad()
                loss.backward()
                opt.step()
                if log:
                    log(f"step {it} loss {loss.item():.6e}")

    def snapshot(b):
        d = OrderedDict()
        if let:
            for k, t in b.let.items():
                if "smooth_scale" in k:
                    d[k] = t.detach().clone()
        if lwc:
            for k, t in b.lwc.items():
                d[k] = t.detach().clone()
        for k, t in qs.p.items():
            if k.startswith(b.p) and lrl:
                d[k[len(b.p):]] = t.detach().clone()
        return d

    if mode == "e2e":
        with torch.no_grad():
            for j in range(steps):
                idx = j * batch_size
                h = fp_inps[idx:idx + batch_size]
                for b in blocks:
                    h = b.forward(h, qs, False, maskb, pos)
                fp_inps[idx:idx + batch_size] = h
        for b in blocks:
            b.let = init_let(b, e2e=True) if let else None
            b.lwc = init_lwc(b, recipe) if lwc else None

        def fwd(h):
            for b in blocks:
                h = b.forward(h, qs, True, maskb, pos)
            return h
        if epochs > 0:
            train(blocks, fwd)
        for b in blocks:
            params[b.i] = snapshot(b)
            fuse_block(b, recipe)
    else:
        for b in blocks:
            with torch.no_grad():
                for j in range(nsamples):
                    fp_inps[j] = b.forward(fp_inps[j:j + 1], qs, False, mask1, pos)[0]
            b.let = init_let(b, original_omniquant) if let else None
            b.lwc = init_lwc(b, recipe) if lwc else None
            if epochs > 0:
                train([b], lambda h, b=b: b.forward(h, qs, True, maskb, pos))
            params[b.i] = snapshot(b)
            fuse_block(b, recipe)
            with torch.no_grad():
                for j in range(nsamples):
                    quant_inps[j] = b.forward(quant_inps[j:j + 1], qs, True, mask1, pos)[0]
    return dict(params=params, act_dict=qs.act_dict(), losses=losses, sd=sd, qs=qs)
