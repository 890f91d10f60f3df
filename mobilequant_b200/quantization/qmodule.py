"""Fake-quant module surface of MobileQuant, backed by fused sm_100a kernels (drop-in mirror of
mobilellm/quantization/qmodule.py -- same class names, constructor signatures, attribute names and JSON artefacts).

What differs from the reference is only *how* a forward/backward is executed: one Quantizer call is one fused CUDA
kernel (libmqb200: mq_fq_fwd / mq_fq_bwd / mq_wprep_fwd / mq_wprep_bwd) instead of ~8 forward and ~10 backward ATen
launches, and a QLinear under LET does transform + min/max + LWC + fake-quant in a single pass over the weight.
There is no CPU execution path: tensors must live on a CUDA device (host logic -- configs, rewriting, JSON export --
works anywhere).
"""
import math
import os
from copy import deepcopy
from dataclasses import dataclass
import torch
import torch.nn as nn
from ..model.hf_model import HFRMSNorm
from ..model.ops import FMatMul
from .functional import (StaticFakeQuantFn, LetLwcWeightQuantFn, AttnProbsFn, SiluGateFn, RmsNormL2Fn, QkvRopeFn,
                         GroupedWeightFn)
from .. import kernels as K

CLIPMIN = 1e-5   # qm:11
CLIPMAX = 1e6    # qm:12
MODE_NONE, MODE_DIV, MODE_MUL = 0, 1, 2


# ---------------------------------------------------------------------------------------------------------------
# scalar helpers (host side; tiny tensors) -- qm:40-76
# ---------------------------------------------------------------------------------------------------------------
def compute_scale_offset_from_min_max(min_val, max_val, bitwdith, is_symmetric):
    if not isinstance(min_val, torch.Tensor):
        min_val = torch.tensor(min_val)
    if not isinstance(max_val, torch.Tensor):
        max_val = torch.tensor(max_val)
    if is_symmetric:
        alpha = torch.maximum(min_val.abs(), max_val.abs())
        beta = 0
        q_min, q_max = -2 ** (bitwdith - 1), 2 ** (bitwdith - 1) - 1
    else:
        alpha = max_val - min_val
        beta = min_val
        q_min, q_max = 0, 2 ** bitwdith - 1
    scale = (alpha / q_max).clamp(min=CLIPMIN, max=CLIPMAX)
    offset = -(beta / scale).round()
    return scale, offset, alpha, beta, q_min, q_max


def compute_min_max_from_scale_offset(scale, offset, bitwidth, is_symmetric):
    q_max = 2 ** (bitwidth - 1) - 1 if is_symmetric else 2 ** bitwidth - 1
    scale = scale.clamp(min=CLIPMIN, max=CLIPMAX)
    alpha = scale * q_max
    beta = -offset * scale
    max_val = alpha + beta
    min_val = beta if not is_symmetric else -max_val
    return min_val, max_val


@dataclass
class QuantConfig:                       # qm:81-107 (string-valued dict form is the default_qcfg.json schema)
    bitwidth: int = 32
    group_size: int = -1
    is_symmetric: bool = False
    is_per_channel: bool = False
    is_dynamic: bool = False

    @classmethod
    def from_dict(cls, cfg):
        t = ("True", "true")
        return cls(bitwidth=int(cfg["bitwidth"]), group_size=int(cfg["group_size"]),
                   is_symmetric=cfg["is_symmetric"] in t, is_per_channel=cfg["is_per_channel"] in t,
                   is_dynamic=cfg["is_dynamic"] in t)

    def to_dict(self):
        return {k: str(getattr(self, k)) for k in ("bitwidth", "group_size", "is_symmetric", "is_per_channel", "is_dynamic")}


# ---------------------------------------------------------------------------------------------------------------
class Quantizer(nn.Module):
    """qm:112-295."""

    def __init__(self, qcfg):
        super().__init__()
        self.qcfg = deepcopy(qcfg)
        self.lwc = False
        self.enable = True

    def update_qcfg(self, qcfg):
        if not isinstance(qcfg, QuantConfig):
            assert isinstance(qcfg, dict)
            qcfg = QuantConfig.from_dict(qcfg)
        self.qcfg = deepcopy(qcfg)
        for a in ("scale", "offset"):
            if hasattr(self, a):
                delattr(self, a)

    def export_qcfg(self):
        return self.qcfg.to_dict()

    def _check_supported(self):
        if self.qcfg.group_size != -1:
            # the reference's own enable_lwc raises NameError on this path (qm:139-141, SURVEY.md hard part 8)
            raise NotImplementedError("per-group quantisation (group_size != -1) is not on the MobileQuant hot path")

    def enable_lwc(self, w):             # qm:133-151
        self._check_supported()
        self.lwc = True
        n = w.shape[0] if self.qcfg.is_per_channel else None
        shape = (n, 1) if self.qcfg.is_per_channel else (1,)
        self.upbound_factor = nn.Parameter(torch.ones(shape, dtype=w.dtype, device=w.device) * 4.0)
        self.lowbound_factor = nn.Parameter(torch.ones(shape, dtype=w.dtype, device=w.device) * 4.0)

    def disable_lwc(self):               # qm:153-157
        self.lwc = False
        for a in ("upbound_factor", "lowbound_factor"):
            if hasattr(self, a):
                delattr(self, a)

    @torch.no_grad()
    def run_lwc(self, input_):
        """qm:159-185: clamp the weight to its (learned) clipping range at fuse time; consumes the LWC parameters."""
        self._check_supported()
        x = input_
        if self.qcfg.is_per_channel:
            min_val, max_val = torch.amin(x, dim=-1, keepdim=True), torch.amax(x, dim=-1, keepdim=True)
        else:
            y = x.contiguous().view(-1)
            min_val, max_val = torch.amin(y, dim=-1), torch.amax(y, dim=-1)
        if self.lwc:
            max_val = torch.sigmoid(self.upbound_factor) * max_val
            min_val = torch.sigmoid(self.lowbound_factor) * min_val
            for a in ("scale", "offset"):
                if hasattr(self, a):
                    delattr(self, a)
            self.disable_lwc()
        return x.clamp(min_val, max_val).type(input_.dtype)

    def set_scale_offset_from_minmax(self, min_val, max_val, cache_mode=None, device=None):   # qm:216-245
        scale, offset, _, _, q_min, q_max = compute_scale_offset_from_min_max(min_val, max_val, self.qcfg.bitwidth,
                                                                             self.qcfg.is_symmetric)
        self.qmin, self.qmax = q_min, q_max
        scale, offset = scale.to(device), offset.to(device)
        for a in ("scale", "offset"):
            if hasattr(self, a) and cache_mode in ("parameter", "buffer"):
                delattr(self, a)
        if cache_mode == "parameter":
            self.register_parameter("scale", nn.Parameter(scale))
            self.register_parameter("offset", nn.Parameter(offset))
        elif cache_mode == "buffer":
            self.register_buffer("scale", scale)
            self.register_buffer("offset", offset)
        else:
            self.scale, self.offset = scale, offset

    def set_scale_offset_from_tensor(self, x, cache_mode=None):
        if self.qcfg.is_per_channel:
            mn, mx = torch.amin(x, dim=-1, keepdim=True), torch.amax(x, dim=-1, keepdim=True)
        else:
            mn, mx = x.min(), x.max()
        self.set_scale_offset_from_minmax(mn, mx, cache_mode, x.device)

    def _qrange(self):
        b = self.qcfg.bitwidth
        return (-2 ** (b - 1), 2 ** (b - 1) - 1) if self.qcfg.is_symmetric else (0, 2 ** b - 1)

    def forward(self, input_, use_scale_offset_as="parameter", let=None, out=None):
        """qm:251-295.  `let` (only passed by QLinear/QRMSNorm under smooth_lm_temporary) fuses the LET transform of
        the raw weight into the same kernel: dict(col_fac, col_mode, row_fac, row_mode)."""
        if not self.enable or self.qcfg.bitwidth > 16:
            return input_ if let is None else materialize_let(input_, let)
        self._check_supported()
        x = input_
        dynamic = self.qcfg.is_dynamic or self.lwc or not hasattr(self, "scale") or not hasattr(self, "offset")
        if let is not None and not (self.qcfg.is_dynamic or self.lwc):
            # --let without --lwc: the reference quantises the materialised temp_weight with a range it caches on the
            # first forward and reuses afterwards (qm:262-277); only the LWC / dynamic path recomputes the range per step
            x = materialize_let(x, let)
            let = None
        if dynamic:
            x2 = x.reshape(-1, x.shape[-1]) if x.dim() != 2 else x
            if x.dim() == 1:
                x2 = x.reshape(1, -1)
            su = torch.sigmoid(self.upbound_factor) if self.lwc else None     # qm:271-272
            sl = torch.sigmoid(self.lowbound_factor) if self.lwc else None
            lt = let or {}
            y, scale, offset = LetLwcWeightQuantFn.apply(
                x2, lt.get("col_fac"), lt.get("col_mode", 0), lt.get("row_fac"), lt.get("row_mode", 0), su, sl,
                self.qcfg.bitwidth, self.qcfg.is_symmetric, self.qcfg.is_per_channel, out if x2.shape == x.shape else None)
            self.qmin, self.qmax = self._qrange()
            if self.qcfg.is_per_channel:
                scale, offset = scale.reshape(-1, 1), offset.reshape(-1, 1)
            else:
                scale, offset = scale.reshape(()), offset.reshape(())
            if self.qcfg.is_dynamic or self.lwc:
                if isinstance(getattr(self, "scale", None), nn.Parameter):
                    raise TypeError("cannot assign a dynamic scale over a cached nn.Parameter (reference qm:244)")
                self.scale, self.offset = scale, offset
                return y.reshape(input_.shape)
            # first forward of a static weight quantizer: cache the range (qm:276-277), then quantise with it
            for a in ("scale", "offset"):
                if hasattr(self, a):
                    delattr(self, a)
            if use_scale_offset_as == "parameter":
                self.register_parameter("scale", nn.Parameter(scale)); self.register_parameter("offset", nn.Parameter(offset))
            elif use_scale_offset_as == "buffer":
                self.register_buffer("scale", scale); self.register_buffer("offset", offset)
            else:
                self.scale, self.offset = scale, offset
        if self.scale.device != x.device:
            self.scale.data = self.scale.to(x.device)
        if self.offset.device != x.device:
            self.offset.data = self.offset.to(x.device)
        return StaticFakeQuantFn.apply(x, self.scale, self.offset, self.qmin, self.qmax)


def materialize_let(w, let):
    """The temp_weight of alg:60-96 as plain tensor ops (only used when a weight quantizer is disabled)."""
    t = w if w.dim() == 2 else w.reshape(1, -1)
    if let.get("col_mode", 0) == MODE_MUL:
        t = t * let["col_fac"].view(1, -1)
    elif let.get("col_mode", 0) == MODE_DIV:
        t = t / let["col_fac"].view(1, -1)
    if let.get("row_mode", 0) == MODE_DIV:
        t = t / let["row_fac"].view(-1, 1)
    elif let.get("row_mode", 0) == MODE_MUL:
        t = t * let["row_fac"].view(-1, 1)
    return t.reshape(w.shape)


def _static_params(quant, device):
    """[scale, offset, qmin, qmax] of a per-tensor static quantizer for the fused calibration kernels ([None, None, 0, 0] when it
    is absent / disabled); False when the fused kernels do not cover it (dynamic, LWC, per-channel, range not cached yet)."""
    if quant is None or not quant.enable or quant.qcfg.bitwidth > 16:
        return [None, None, 0.0, 0.0]
    if quant.qcfg.is_dynamic or quant.lwc or quant.qcfg.is_per_channel or quant.qcfg.group_size != -1 \
            or not hasattr(quant, "scale") or not hasattr(quant, "offset") or quant.scale.numel() != 1:
        return False
    for t in (quant.scale, quant.offset):                          # same lazy move as Quantizer.forward
        if t.device != device:
            t.data = t.to(device)
    return [quant.scale, quant.offset, quant.qmin, quant.qmax]


def _active(quant):
    return quant is not None and quant.enable and quant.qcfg.bitwidth <= 16


def _grouped_weight(mods):
    """cat([m._fq_weight() for m in mods], 0) for the single GEMM of sibling projections.  After the first call each module's
    weight pass writes straight into its row slice of one shared buffer (m._wout), so from then on this is a zero-copy view."""
    ws = [m._fq_weight() for m in mods]
    buf = getattr(mods[0], "_wgroup", None)
    if buf is not None and all(getattr(m, "_wout", None) is not None and w.data_ptr() == m._wout.data_ptr() and w.shape == m._wout.shape
                               for m, w in zip(mods, ws)):
        return GroupedWeightFn.apply(buf, *ws)
    if buf is None and all(w.dtype == torch.float32 and w.dim() == 2 for w in ws):
        buf = torch.empty((sum(w.shape[0] for w in ws), ws[0].shape[1]), dtype=torch.float32, device=ws[0].device)
        r0 = 0
        for m, w in zip(mods, ws):
            m._wout = buf[r0:r0 + w.shape[0]]
            r0 += w.shape[0]
        mods[0]._wgroup = buf
    return torch.cat(ws, dim=0)


def _fused_enabled(name):
    """MQB200_FUSED_<NAME>=0 runs that piece of the calibration block op by op (A/B checks)."""
    return os.environ.get(f"MQB200_FUSED_{name}", "1") != "0"


_rope_cache = {}


def _rope_tables(builder, position_ids, rot, theta, device):
    """cos / sin of hm:308-318 for these position ids, built once per (positions, width, base) instead of once per layer call."""
    key = (id(position_ids), position_ids._version, rot, float(theta), str(device))
    hit = _rope_cache.get(key)
    if hit is None or hit[2] is not position_ids:
        if rot == 0:
            hit = (None, None, position_ids)
        else:
            with torch.no_grad():
                cos, sin = builder(position_ids, rot, theta, device, torch.float32)
            hit = (cos.contiguous(), sin.contiguous(), position_ids)      # the reference keeps id() from being reused
        if len(_rope_cache) > 8:
            _rope_cache.clear()
        _rope_cache[key] = hit
    return hit[0], hit[1]


_causal_cache = {}


def is_causal_mask(mask, tq, tk):
    """True when `mask` is the additive causal mask transformers builds for a full prompt (hm:1548-1555: finfo.min above the
    diagonal, 0 elsewhere, [B, 1, T, T]).  Checked by content once per mask tensor (never during stream capture: an unknown
    mask then counts as not causal)."""
    if mask is None or tq != tk or mask.dim() != 4 or mask.shape[1] != 1 or mask.shape[-2:] != (tq, tk) or mask.dtype != torch.float32:
        return False
    key = (id(mask), mask._version)
    hit = _causal_cache.get(key)
    if hit is None or hit[1] is not mask:
        if mask.is_cuda and torch.cuda.is_current_stream_capturing():
            return False
        ref = torch.triu(torch.full((tq, tk), torch.finfo(mask.dtype).min, dtype=mask.dtype, device=mask.device), diagonal=1)
        hit = (bool((mask == ref).all()), mask)                            # the reference keeps id() from being reused
        if len(_causal_cache) > 8:
            _causal_cache.clear()
        _causal_cache[key] = hit
    return hit[0]


# ---------------------------------------------------------------------------------------------------------------
class _QBase:
    """Shared (de)serialisation of the quantizer triplets (qm:314-339 and friends)."""
    _slots = ("input", "weight", "output")

    def _q(self, slot):
        return getattr(self, f"{slot}_quantizer", None)

    def export_qcfg(self):
        return {s: self._q(s).export_qcfg() for s in self._slots if self._q(s) is not None}

    def _update(self, **cfgs):
        for s, cfg in cfgs.items():
            if self._q(s) is not None and cfg is not None:
                self._q(s).update_qcfg(cfg)

    def _set_ranges(self, act_scale, use_scale_offset_as, device=None):
        for s in ("input", "input2", "output"):
            q = self._q(s)
            if q is None or s not in self._slots:
                continue
            if s == "input2" and s not in act_scale:
                q.set_scale_offset_from_minmax(0.0, 1.0, use_scale_offset_as, device)      # qm:731-734
            else:
                q.set_scale_offset_from_minmax(act_scale[s][0], act_scale[s][1], use_scale_offset_as, device)

    def _fq_weight(self):
        """The (LET-transformed,) fake-quantised weight of a QLinear / QRMSNorm / QLayerNorm forward.  The calibration loops may
        compute it ahead of time on a side stream (algorithm.py:_prefetch_weights: the weight pass is HBM-bound and independent
        of the activations, the rest of a step is a chain of small kernels); the stashed tensor is picked up here after waiting
        for its event."""
        pre = getattr(self, "_prepared_weight", None)
        if pre is not None:
            self._prepared_weight = None
            w, ev = pre
            torch.cuda.current_stream().wait_event(ev)
            w.record_stream(torch.cuda.current_stream())
            return w
        return self._fq_weight_now()

    def _fq_weight_now(self):
        weight, let = self._let_weight(self.weight)
        if self.weight_quantizer is not None:
            out = getattr(self, "_wout", None)              # row slice of a buffer shared with sibling projections (_grouped_weight)
            if out is not None and (out.shape != weight.shape or out.device != weight.device):
                out = None
            return self.weight_quantizer(weight, let=let, out=out)
        if let is not None:
            return materialize_let(weight, let)
        return weight

    def _let_weight(self, raw_weight):
        """weight fed to the weight quantizer + the fused LET description (None when not under smooth_lm_temporary)."""
        if not getattr(self, "use_temporary_parameter", False):
            return raw_weight, None
        let = getattr(self, "let", None)
        if let is None:
            return getattr(self, "temp_weight", raw_weight), None
        return raw_weight, let


class QLinear(nn.Linear, _QBase):
    """qm:298-405."""

    def __init__(self, kargs, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        super().__init__(**kargs)
        self.input_quantizer = Quantizer(input_quant_cfg) if input_quant_cfg is not None else None
        self.weight_quantizer = Quantizer(weight_quant_cfg) if weight_quant_cfg is not None else None
        self.output_quantizer = Quantizer(output_quant_cfg) if output_quant_cfg is not None else None
        self.use_temporary_parameter = False

    def update_qcfg(self, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        self._update(input=input_quant_cfg, weight=weight_quant_cfg, output=output_quant_cfg)

    def set_scale_offset(self, act_scale, use_scale_offset_as="parameter"):
        self._set_ranges(act_scale, use_scale_offset_as, self.weight.device)

    def _bias(self):
        return self.bias if not self.use_temporary_parameter else getattr(self, "temp_bias", self.bias)

    def forward(self, input_, input_quantized=False):
        bias = self._bias()
        weight = self._fq_weight()
        if self.input_quantizer is not None and not input_quantized:
            input_ = self.input_quantizer(input_)
        out = nn.functional.linear(input_, weight, bias=bias)
        if self.output_quantizer is not None:
            out = self.output_quantizer(out)
        return out

    @staticmethod
    def from_float(module, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        kargs = dict(in_features=module.in_features, out_features=module.out_features, bias=module.bias is not None,
                     device=module.weight.device, dtype=module.weight.dtype)
        out = QLinear(kargs, input_quant_cfg, weight_quant_cfg, output_quant_cfg)
        with torch.no_grad():
            out.weight.copy_(module.weight)
            if out.bias is not None:
                out.bias.copy_(module.bias)
        return out

    @staticmethod
    def to_float(module):
        bias = getattr(module, "bias", None)
        out = nn.Linear(module.in_features, module.out_features, bias=bias is not None, device=module.weight.device,
                        dtype=module.weight.dtype)
        with torch.no_grad():
            out.weight.copy_(module.weight)
            if bias is not None:
                out.bias.copy_(bias)
        return out


class QMatMul(nn.Module, _QBase):
    """qm:408-466."""
    _slots = ("input", "input2", "output")

    def __init__(self, input_quant_cfg, input2_quant_cfg, output_quant_cfg):
        super().__init__()
        self.input_quantizer = Quantizer(input_quant_cfg) if input_quant_cfg is not None else None
        self.input2_quantizer = Quantizer(input2_quant_cfg) if input2_quant_cfg is not None else None
        self.output_quantizer = Quantizer(output_quant_cfg) if output_quant_cfg is not None else None

    def update_qcfg(self, input_quant_cfg, input2_quant_cfg, output_quant_cfg):
        self._update(input=input_quant_cfg, input2=input2_quant_cfg, output=output_quant_cfg)

    def set_scale_offset(self, act_scale, use_scale_offset_as="parameter"):
        self._set_ranges(act_scale, use_scale_offset_as)

    def forward(self, x1, x2, input_quantized=False):
        if self.input_quantizer is not None and not input_quantized:
            x1 = self.input_quantizer(x1)
        if self.input2_quantizer is not None:
            x2 = self.input2_quantizer(x2)
        out = torch.matmul(x1, x2)
        if self.output_quantizer is not None:
            out = self.output_quantizer(out)
        return out

    def fused_attention(self, attn, hidden_states, attention_mask, position_ids):
        """HFAttention.forward (hm:470-540) up to o_proj's input, with this module as attn.qk_bmm:
          * q / k / v projections as ONE GEMM over the concatenated fake-quantised weights,
          * their output quantizers, head split, RoPE and the matmul input quantizers as one kernel (QkvRopeFn),
          * grouped-query heads addressed through views ([B, nkv, rep*T, hd] x [B, nkv, hd, T]) instead of repeat_kv copies,
          * the score quantizer / scale / mask / softmax / probability quantizer as one kernel (AttnProbsFn).
        None when a piece is not covered (the caller then runs the module graph op by op)."""
        from ..model.hf_model import rope_cos_sin
        pv, x = attn.pv_bmm, hidden_states
        lins = (attn.q_proj, attn.k_proj, attn.v_proj)
        if not _fused_enabled("QKV") or not _fused_enabled("PROBS") or not isinstance(pv, QMatMul) or not all(isinstance(m, QLinear) for m in lins) \
                or not x.is_cuda or x.dtype != torch.float32 or position_ids is None:
            return None
        B, T, _ = x.shape
        nh, nkv, hd, rot = attn.num_heads, attn.num_key_value_heads, attn.head_dim, attn.rotary_dim
        if hd % 8 or rot % 8 or (hd - rot) % 8 or nh % nkv or any(_active(m.input_quantizer) for m in lins) \
                or len({m.bias is None for m in lins}) != 1 or not is_causal_mask(attention_mask, T, T) or not K.attn_probs_supported(T):
            return None
        params = []
        for quant in (lins[0].output_quantizer, lins[1].output_quantizer, lins[2].output_quantizer, self.input_quantizer,
                      self.input2_quantizer, pv.input2_quantizer, self.output_quantizer, pv.input_quantizer):
            pq = _static_params(quant, x.device)
            if pq is False:
                return None
            params += pq
        if any(m.weight.dtype != torch.float32 for m in lins):
            return None
        bs = [m._bias() for m in lins]
        y = nn.functional.linear(x, _grouped_weight(lins), None if bs[0] is None else torch.cat(bs, dim=0))
        cos, sin = _rope_tables(rope_cos_sin, position_ids, rot, attn.rope_theta, x.device)
        q, k, v = QkvRopeFn.apply(y, cos, sin, nh, nkv, hd, rot, *params[:24])
        rep = nh // nkv
        scores = torch.matmul(q.view(B, nkv, rep * T, hd), k.transpose(2, 3)).view(B, nh, T, T)
        mul = (torch.ones((), dtype=torch.float32) / torch.tensor(math.sqrt(hd), dtype=torch.float32)).item()
        probs = AttnProbsFn.apply(scores, mul, *params[24:])
        out = torch.matmul(probs.view(B, nkv, rep * T, T), v).view(B, nh, T, hd)
        if pv.output_quantizer is not None:
            out = pv.output_quantizer(out)
        return out.transpose(1, 2).contiguous().view(B, T, nh * hd)

    def fused_probs(self, q, kt, head_dim, attention_mask, pv_bmm):
        """The attention core of HFAttention.forward (hm:514-534) with this module as qk_bmm:
            pv_bmm.input_quantizer(softmax(self(q, kt) / sqrt(head_dim) + attention_mask))
        as matmul + ONE fused kernel (csrc/calib_attn.cu) instead of the element-wise chain over [B, nh, T, T].  Returns None
        when the fused kernel does not cover the case (non-causal mask, cached ranges missing, dynamic / per-channel quantizers,
        reduced precision, T > 2048): the caller then runs the chain op by op."""
        if not _fused_enabled("PROBS") or not isinstance(pv_bmm, QMatMul) or not q.is_cuda or q.dtype != torch.float32 \
                or kt.dtype != torch.float32:
            return None
        if not is_causal_mask(attention_mask, q.shape[-2], kt.shape[-1]) or not K.attn_probs_supported(kt.shape[-1]):
            return None
        params = []
        for quant in (self.output_quantizer, pv_bmm.input_quantizer):
            pq = _static_params(quant, q.device)
            if pq is False:
                return None
            params += pq
        x1 = self.input_quantizer(q) if self.input_quantizer is not None else q
        x2 = self.input2_quantizer(kt) if self.input2_quantizer is not None else kt
        scores = torch.matmul(x1, x2)
        mul = (torch.ones((), dtype=torch.float32) / torch.tensor(math.sqrt(head_dim), dtype=torch.float32)).item()
        return AttnProbsFn.apply(scores, mul, *params)


class QRMSNorm(HFRMSNorm, _QBase):
    """qm:469-576."""

    def __init__(self, kargs, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        super().__init__(**kargs)
        self.input_quantizer = Quantizer(input_quant_cfg) if input_quant_cfg is not None else None
        self.weight_quantizer = Quantizer(weight_quant_cfg) if weight_quant_cfg is not None else None
        self.output_quantizer = Quantizer(output_quant_cfg) if output_quant_cfg is not None else None
        self.use_temporary_parameter = False

    def update_qcfg(self, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        self._update(input=input_quant_cfg, weight=weight_quant_cfg, output=output_quant_cfg)

    def set_scale_offset(self, act_scale, use_scale_offset_as="parameter"):
        self._set_ranges(act_scale, use_scale_offset_as, self.weight.device)

    def _qweight(self):
        return self._fq_weight()

    def _fused(self, input_, weight, bias):
        """The whole forward as one kernel (csrc/calib_act.cu) when it is the L2-norm form with per-tensor static quantizers."""
        if not _fused_enabled("NORM") or not self.l2norm_as_rmsnorm or not input_.is_cuda or input_.dtype != torch.float32 \
                or weight.dtype != torch.float32 or not K.rmsnorm_l2_supported(input_.shape[-1]):
            return None
        l2 = self.l2norm
        if l2.p != 2 or l2.dim not in (-1, input_.dim() - 1):
            return None
        params = []
        for quant in (self.input_quantizer, self.output_quantizer):
            pq = _static_params(quant, input_.device)
            if pq is False:
                return None
            params += pq
        return RmsNormL2Fn.apply(input_, weight, bias, self.alpha, l2.eps, *params)

    def forward(self, input_):
        weight = self._qweight()
        bias = self.bias if not self.use_temporary_parameter else getattr(self, "temp_bias", self.bias)
        out = self._fused(input_, weight, bias)
        if out is not None:
            return out
        if self.input_quantizer is not None:
            input_ = self.input_quantizer(input_)
        out = self.forward_impl(input_, weight, bias)
        if self.output_quantizer is not None:
            out = self.output_quantizer(out)
        return out

    @staticmethod
    def from_float(module, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        kargs = dict(dim=len(module.weight), eps=module.eps, device=module.weight.device, dtype=module.weight.dtype,
                     l2norm_as_rmsnorm=module.l2norm_as_rmsnorm)
        out = QRMSNorm(kargs, input_quant_cfg, weight_quant_cfg, output_quant_cfg)
        with torch.no_grad():
            out.weight.copy_(module.weight)
        return out

    @staticmethod
    def to_float(module):
        out = HFRMSNorm(dim=len(module.weight), eps=module.eps, device=module.weight.device, dtype=module.weight.dtype,
                        l2norm_as_rmsnorm=module.l2norm_as_rmsnorm)
        with torch.no_grad():
            out.weight.copy_(module.weight)
        b = getattr(module, "bias", None)
        if b is not None:                     # zero-shift buffer registered by smooth_ln_fcs_inplace (alg:109-110)
            out.bias = nn.Parameter(b.detach().clone())
        return out


class QLayerNorm(nn.LayerNorm, _QBase):
    """qm:579-688."""

    def __init__(self, kargs, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        super().__init__(**kargs)
        self.input_quantizer = Quantizer(input_quant_cfg) if input_quant_cfg is not None else None
        self.weight_quantizer = Quantizer(weight_quant_cfg) if weight_quant_cfg is not None else None
        self.output_quantizer = Quantizer(output_quant_cfg) if output_quant_cfg is not None else None
        self.use_temporary_parameter = False

    def update_qcfg(self, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        self._update(input=input_quant_cfg, weight=weight_quant_cfg, output=output_quant_cfg)

    def set_scale_offset(self, act_scale, use_scale_offset_as="parameter"):
        self._set_ranges(act_scale, use_scale_offset_as, self.weight.device)

    def forward(self, input_):
        bias = self.bias if not self.use_temporary_parameter else getattr(self, "temp_bias", self.bias)
        weight = self._fq_weight()
        if self.input_quantizer is not None:
            input_ = self.input_quantizer(input_)
        out = nn.functional.layer_norm(input_, input_.shape[-1:], weight=weight, bias=bias, eps=self.eps)
        if self.output_quantizer is not None:
            out = self.output_quantizer(out)
        return out

    @staticmethod
    def from_float(module, input_quant_cfg, weight_quant_cfg, output_quant_cfg):
        kargs = dict(normalized_shape=len(module.weight), eps=module.eps, elementwise_affine=module.elementwise_affine,
                     device=module.weight.device, dtype=module.weight.dtype)
        out = QLayerNorm(kargs, input_quant_cfg, weight_quant_cfg, output_quant_cfg)
        with torch.no_grad():
            out.weight.copy_(module.weight)
            if out.bias is not None and module.bias is not None:
                out.bias.copy_(module.bias)
        return out

    @staticmethod
    def to_float(module):
        out = nn.LayerNorm(len(module.weight), eps=module.eps, elementwise_affine=module.elementwise_affine,
                           bias=module.bias is not None, device=module.weight.device, dtype=module.weight.dtype)
        with torch.no_grad():
            out.weight.copy_(module.weight)
            if module.bias is not None:
                out.bias.copy_(module.bias)
        return out


class QSiLU(nn.Module, _QBase):
    """qm:691-753: fq_out(x * fq_in2(sigmoid(x)))."""
    _slots = ("input", "input2", "output")

    def __init__(self, input_quant_cfg, input2_quant_cfg, output_quant_cfg):
        super().__init__()
        self.input_quantizer = Quantizer(input_quant_cfg) if input_quant_cfg is not None else None
        self.input2_quantizer = Quantizer(input2_quant_cfg) if input2_quant_cfg is not None else None
        self.output_quantizer = Quantizer(output_quant_cfg) if output_quant_cfg is not None else None

    def update_qcfg(self, input_quant_cfg, input2_quant_cfg, output_quant_cfg):
        self._update(input=input_quant_cfg, input2=input2_quant_cfg, output=output_quant_cfg)

    def set_scale_offset(self, act_scale, use_scale_offset_as="parameter"):
        self._set_ranges(act_scale, use_scale_offset_as)

    def forward(self, x):
        if self.input_quantizer is not None:
            x = self.input_quantizer(x)
        y = torch.sigmoid(x)
        if self.input2_quantizer is not None:
            y = self.input2_quantizer(y)
        out = x * y
        if self.output_quantizer is not None:
            out = self.output_quantizer(out)
        return out


    def fused_mlp(self, x, w1, w3, w2):
        """HFMLP.forward (hm:1057-1061) with this module as act_fn, up to w2's GEMM:
            w2.input_quantizer( self(w1(x)) * w3(x) )
        as ONE GEMM over the concatenated (fake-quantised) w1 / w3 weights + ONE fused element-wise kernel (csrc/calib_act.cu:
        both output quantizers, the sigmoid and its quantizer, the products, w2's input quantizer).  None when not covered; the
        caller then goes module by module."""
        if not _fused_enabled("GATE") or not all(isinstance(m, QLinear) for m in (w1, w3, w2)) or not x.is_cuda \
                or x.dtype != torch.float32 or w1.out_features != w3.out_features or w1.out_features % 4 != 0:
            return None
        if any(_active(m.input_quantizer) for m in (self, w1, w3)) or (w1.bias is None) != (w3.bias is None):
            return None                                    # quantised by the producer in every MobileQuant recipe (qm:848-858)
        params = []
        for quant in (w1.output_quantizer, w3.output_quantizer, self.input2_quantizer, self.output_quantizer, w2.input_quantizer):
            pq = _static_params(quant, x.device)
            if pq is False:
                return None
            params += pq
        if w1.weight.dtype != torch.float32 or w3.weight.dtype != torch.float32:
            return None
        ba, bb = w1._bias(), w3._bias()
        y = nn.functional.linear(x, _grouped_weight((w1, w3)), None if ba is None else torch.cat((ba, bb), dim=0))
        return SiluGateFn.apply(y, *params)


class QGELU(nn.Module, _QBase):
    """qm:756-799: fq_out(erf-GELU(x)) (exact GELU even for tanh-GELU checkpoints, as the reference)."""
    _slots = ("input", "output")

    def __init__(self, input_quant_cfg, output_quant_cfg):
        super().__init__()
        self.input_quantizer = Quantizer(input_quant_cfg) if input_quant_cfg is not None else None
        self.output_quantizer = Quantizer(output_quant_cfg) if output_quant_cfg is not None else None

    def update_qcfg(self, input_quant_cfg, output_quant_cfg):
        self._update(input=input_quant_cfg, output=output_quant_cfg)

    def set_scale_offset(self, act_scale, use_scale_offset_as="parameter"):
        self._set_ranges(act_scale, use_scale_offset_as)

    def forward(self, x):
        if self.input_quantizer is not None:
            x = self.input_quantizer(x)
        out = nn.functional.gelu(x)
        if self.output_quantizer is not None:
            out = self.output_quantizer(out)
        return out


# ---------------------------------------------------------------------------------------------------------------
# model rewriting and (de)serialisation -- qm:835-970
# ---------------------------------------------------------------------------------------------------------------
_NO_INPUT_Q = ("q_proj", "k_proj", "v_proj", "o_proj", "w1", "w3")


def create_sim_qmodel(model, default_weight_qcfg=None, default_act_qcfg=None):
    """qm:835-865: in-place swap of Linear / FMatMul / SiLU / GELU / norms for their Q* versions (lm_head and the final
    model.norm stay in floating point)."""
    wq = default_weight_qcfg if default_weight_qcfg is not None else QuantConfig()
    aq = default_act_qcfg if default_act_qcfg is not None else QuantConfig()
    for name, module in reversed(list(model._modules.items())):
        if "lm_head" in name or ("norm" in name and "layernorm" not in name):
            continue
        if isinstance(module, nn.Linear):
            q = QLinear.from_float(module, aq, wq, aq)
            if any(k in name for k in _NO_INPUT_Q):
                q.input_quantizer = None            # already quantised by the producer (qm:848-850)
            model._modules[name] = q
        elif isinstance(module, FMatMul):
            model._modules[name] = QMatMul(aq, aq, aq)
        elif isinstance(module, nn.SiLU):
            q = QSiLU(aq, aq, aq); q.input_quantizer = None
            model._modules[name] = q
        elif isinstance(module, nn.GELU):
            q = QGELU(aq, aq); q.input_quantizer = None
            model._modules[name] = q
        elif isinstance(module, HFRMSNorm):
            model._modules[name] = QRMSNorm.from_float(module, aq, wq, aq)
        elif isinstance(module, nn.LayerNorm):
            model._modules[name] = QLayerNorm.from_float(module, aq, wq, aq)
        elif len(list(module.children())) > 0:
            create_sim_qmodel(module, wq, aq)
    return model


def create_fp_model(model):
    """qm:889-905."""
    for name, module in reversed(list(model._modules.items())):
        if isinstance(module, QLinear):
            model._modules[name] = QLinear.to_float(module)
        elif isinstance(module, QRMSNorm):
            model._modules[name] = QRMSNorm.to_float(module)
        elif isinstance(module, QLayerNorm):
            model._modules[name] = QLayerNorm.to_float(module)
        elif isinstance(module, QMatMul):
            model._modules[name] = FMatMul()
        elif isinstance(module, QSiLU):
            model._modules[name] = nn.SiLU()
        elif isinstance(module, QGELU):
            model._modules[name] = nn.GELU()
        elif len(list(module.children())) > 1:
            create_fp_model(module)
    return model


_QTYPES = (QLinear, QRMSNorm, QLayerNorm, QMatMul, QSiLU, QGELU)


def export_act_range(model):
    """qm:908-937 -> the act_dict.json payload {module: {input|input2|output: [min, max]}}."""
    act_dict = {}
    for name, m in model.named_modules():
        if not isinstance(m, _QTYPES):
            continue
        entry = act_dict.get(name, {})
        for slot in ("input", "input2", "output"):
            q = getattr(m, f"{slot}_quantizer", None)
            if q is None or slot not in m._slots:
                continue
            mn, mx = compute_min_max_from_scale_offset(q.scale.detach(), q.offset.detach(), q.qcfg.bitwidth,
                                                       q.qcfg.is_symmetric)
            entry[slot] = [mn.item(), mx.item()]
        act_dict[name] = entry
    return act_dict


def update_qcfg(model, override_qcfg):
    """qm:940-954."""
    for name, module in model.named_modules():
        if isinstance(module, (QLinear, QRMSNorm, QLayerNorm)):
            assert name in override_qcfg
            c = override_qcfg[name]
            module.update_qcfg(c.get("input", None), c["weight"], c["output"])
        elif isinstance(module, QMatMul):
            assert name in override_qcfg
            c = override_qcfg[name]
            module.update_qcfg(c["input"], c["input2"], c["output"])
        elif isinstance(module, QSiLU):
            assert name in override_qcfg
            c = override_qcfg[name]
            module.update_qcfg(c.get("input", None), c["input2"], c["output"])
        elif isinstance(module, QGELU):
            assert name in override_qcfg
            c = override_qcfg[name]
            module.update_qcfg(c.get("input", None), c["output"])
    return model


def export_qcfg(model):
    """qm:957-962 -> the default_qcfg.json payload."""
    return {name: m.export_qcfg() for name, m in model.named_modules() if isinstance(m, _QTYPES)}


def set_scale_and_offset(model, act_dict, use_scale_offset_as="buffer"):
    """qm:965-970."""
    for name, module in model.named_modules():
        if isinstance(module, _QTYPES):
            assert name in act_dict, name
            module.set_scale_offset(act_dict[name], use_scale_offset_as)
    return model


def update_quant_cfg(model, use_8bit_softmax_input=False, use_8bit_softmax_output=False):
    """The mixed-precision recipe ptq/mobilequant.py:175-201 applies before calibration (a script-local closure in the
    reference).  ptq/generate_qcfg.py has its own, different rule set: ptq/generate_qcfg.py:create_mixed_precision_model."""
    for name, module in reversed(list(model._modules.items())):
        if isinstance(module, QLinear):
            if any(k in name for k in _NO_INPUT_Q):
                module.input_quantizer = None
            if "w2" in name:
                module.weight_quantizer.qcfg.is_per_channel = True
                module.output_quantizer.qcfg.bitwidth = 16
            elif "o_proj" in name:
                module.output_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (QRMSNorm, QLayerNorm)):
            module.input_quantizer.qcfg.bitwidth = 16
            module.weight_quantizer.qcfg.bitwidth = 16
            module.weight_quantizer.qcfg.is_symmetric = False
            module.weight_quantizer.qcfg.is_per_channel = False
        elif isinstance(module, QMatMul):
            if "qk_bmm" in name and not use_8bit_softmax_input:
                module.output_quantizer.qcfg.bitwidth = 16
            if "pv_bmm" in name and not use_8bit_softmax_output:
                module.input_quantizer.qcfg.bitwidth = 16
        elif isinstance(module, (QSiLU, QGELU)):
            module.input_quantizer = None
        elif len(list(module.children())) > 1:
            update_quant_cfg(module, use_8bit_softmax_input, use_8bit_softmax_output)
    return model
