"""The statically-quantised integer forward (W8A8 / W4A8) of the unified decoder.

In the reference this stage does not exist in executable form: `eval/harness_eval.py --mode custom` (:81-89) *simulates*
it in fp32 (fake-quant, qm:285-290) and the true integer execution is handed to Qualcomm QNN after AIMET export
(device/export.py:311-363).  IntEngine compiles a calibrated model -- fused weights (alg:147-184), default_qcfg.json
(qm:957-962) and act_dict.json (qm:908-937) -- into integer tensors + per-column epilogue constants and runs every
decoder block on integer codes with the libmqb200 kernels:

    h (fp32 residual) --qnorm--> u8 --qgemm(QKV)--> u8 --qrope--> q,k,vT u8 --qattn--> u8 --qgemm(o_proj)+resid--> h
    h --qnorm--> u8 --qgemm(w1||w3)+act LUT*gate--> u8 --qgemm(w2)+resid--> h

Embedding lookup, the final (unquantised) norm and lm_head stay in floating point, as in the reference (qm:843-845).
"""
import math
import os
import numpy as np
import torch
from .. import kernels as K
from ..quantization.qmodule import compute_scale_offset_from_min_max


def _sq(act_dict, qcfg, name, slot):
    """(scale, offset, qmax) python floats of a static activation quantizer (qm:216-245 from act_dict.json)."""
    c = qcfg[name][slot]
    bits, sym = int(c["bitwidth"]), c["is_symmetric"] in ("True", "true")
    if sym:
        raise NotImplementedError("symmetric activation quantizers are not used by any MobileQuant recipe")
    if slot == "input2" and slot not in act_dict[name]:
        mn, mx = 0.0, 1.0
    else:
        mn, mx = act_dict[name][slot]
    s, o, _, _, qmin, qmax = compute_scale_offset_from_min_max(mn, mx, bits, sym)
    o = float(o)
    if o != round(o):
        raise ValueError(f"{name}.{slot}: offset {o} is not integral; reload ranges through act_dict.json (qm:60)")
    return float(s), o + 0.0, float(qmax)


def _f64_sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x.astype(np.float64)))


class IntEngine:
    def __init__(self, model, qcfg, act_dict, device=None, pack4=True):
        """model: float HFForCausalLM holding the *fused* weights (LET folded, LWC-clamped) -- e.g. the output of
        ptq.mobilequant.quantize / create_fp_model, or any float checkpoint for plain static PTQ.
        pack4: 4-bit weight matrices are kept packed (two codes per byte) in HBM and expanded into an L2-sized scratch
        buffer right before their GEMM (mq_unpack4); False keeps one code per byte."""
        self.pack4 = bool(pack4)
        self._scratch = {}
        self.cfg = cfg = model.config
        self.device = dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("IntEngine runs on a CUDA device only (no CPU fallback)")
        self.nh, self.nkv = cfg.num_attention_heads, cfg.num_key_value_heads
        self.hd = cfg.head_dim if cfg.head_dim is not None else cfg.hidden_size // cfg.num_attention_heads
        self.rot = int(cfg.partial_rotary_factor * self.hd)
        self.H, self.I = cfg.hidden_size, cfg.intermediate_size
        self.Ipad = (self.I + 127) // 128 * 128
        self.layernorm = cfg.norm_class.lower() == "layernorm"
        # block variants (hm:1171-1172, 1257-1263): sequential pre-norm (llama / gemma / stablelm) and the parallel block with
        # one shared norm (phi: attention and MLP both read input_layernorm's codes).  The two mixed combinations make the MLP
        # consume either post_attention_layernorm(input_layernorm(x)) or the unquantised residual -- no model family the
        # reference converts uses them (scripts/convert_ckpt.py) and the second is not an integer computation at all.
        self.parallel = bool(cfg.parallel_residual)
        if bool(cfg.shared_attention_norm) != self.parallel:
            raise NotImplementedError("IntEngine covers parallel_residual == shared_attention_norm (both False: llama / gemma / "
                                      "stablelm; both True: phi)")
        self.gated = cfg.num_linears_per_mlp == 3
        self.embed = model.model.embed_tokens.weight.detach().float().to(dev)
        self.final_norm = model.model.norm.to(dev).float()
        self.lm_head = model.lm_head.weight.detach().float().to(dev)
        b = getattr(model.lm_head, "bias", None)                   # hm:1682: the head carries a bias when config.mlp_bias (phi)
        self.lm_head_bias = None if b is None else b.detach().float().to(dev)
        self.layers = [self._build_layer(model.model.layers[i], f"model.layers.{i}", qcfg, act_dict) for i in range(cfg.num_hidden_layers)]
        self._rope_cache = {}
        self._bufs = {}
        # decode: epilogue inside the skinny GEMM (last CTA of a column group) or as its own launch.  Measured on B200
        # (TinyLlama, batch 8, graph replay): fused 1.76 ms/step, separate 1.67 ms/step -- the last-arriving CTA serialises
        # the group's epilogue while a graph node costs less than that -- so the two-launch form is the default.
        self.fused_gemv = os.environ.get("MQB200_GEMV_FUSED", "0") == "1"
        # prefill with packed 4-bit weights: mq_unpack4 into an L2-resident scratch + the W8 GEMM (default), or the fused
        # int4 x int8 GEMM mq_qgemm_w4a8 (MQB200_W4_FUSED=1).  Both are bit-identical; measured on B200 at batch 32 x seq 1024
        # (profiles/r2_bench_w4a8_b32*.json) the fused kernel's GEMMs take 45.6 ms per step against 25.0 + 0.8 ms: expanding
        # nibbles through shared memory (16 KB read + 32 KB written per k-slice) competes with the tensor core's own operand
        # fetch (96 B/clk of the 128 B/clk shared-memory port), while the scratch copy costs only L2 traffic.
        self.fused_w4 = os.environ.get("MQB200_W4_FUSED", "0") == "1"
        # decode: the residual epilogues of o_proj / w2 as their own launches (default) or riding on the following row norm
        # (MQB200_RESID_NORM=1, two launches per layer fewer).  With programmatic dependent launch along the chain an extra
        # launch costs less than the longer norm kernel: measured 1.155 ms (separate) against 1.181 ms (fused) per step.
        self.fused_resid_norm = os.environ.get("MQB200_RESID_NORM", "0") == "1"

    # ---- build ------------------------------------------------------------------------------------------------------
    def _wq(self, w, c, want_fq=False):
        bits, sym, pc = int(c["bitwidth"]), c["is_symmetric"] in ("True", "true"), c["is_per_channel"] in ("True", "true")
        w = w.detach().float().to(self.device).contiguous()
        w2 = w.reshape(1, -1) if w.dim() == 1 else w
        out = K.wprep_fwd(w2, bits, sym, pc, want_fq=want_fq, want_codes=bits <= 8)
        out["sym"] = sym
        out["rows"] = w2.shape[0]
        out["bits"] = bits
        return out

    def _percol(self, wq_list, sx, ox, Kdim, out_q, biases, pad_to=None):
        """Concatenate several quantised weight matrices along N and derive the per-column epilogue constants."""
        codes, sw, ow, cs, bias = [], [], [], [], []
        signed = wq_list[0]["sym"]
        # output quantizers are per tensor (qm:216-245): one (scale, offset) per fused segment, stored per group of
        # `qgroup` columns (the gcd of the segment widths; a multiple of the kernel's 32-column chunk)
        widths = [wq["rows"] for wq in wq_list]
        qgroup = math.gcd(*widths) if len(widths) > 1 else (widths[0] + 31) // 32 * 32
        if qgroup % 32:
            raise NotImplementedError(f"fused projection segments {widths} must share a multiple-of-32 column group")
        so = torch.tensor([oq[0] for oq, n in zip(out_q, widths) for _ in range(max(1, n // qgroup))], dtype=torch.float32, device=self.device)
        oo = torch.tensor([oq[1] for oq, n in zip(out_q, widths) for _ in range(max(1, n // qgroup))], dtype=torch.float32, device=self.device)
        for wq, oq, b in zip(wq_list, out_q, biases):
            n = wq["rows"]
            codes.append(wq["codes"].view(torch.int8 if signed else torch.uint8))
            sw.append(wq["scale"].expand(n) if wq["scale"].numel() == 1 else wq["scale"])
            ow.append(wq["offset"].expand(n) if wq["offset"].numel() == 1 else wq["offset"])
            cs.append(wq["colsum"])
            bias.append(torch.zeros(n, device=self.device) if b is None else b.detach().float().to(self.device))
        codes = torch.cat(codes); sw = torch.cat(sw).float(); ow = torch.cat(ow).to(torch.int64); cs = torch.cat(cs).to(torch.int64)
        sxw = (torch.tensor(sx, dtype=torch.float32, device=self.device) * sw).contiguous()
        c0 = (Kdim * int(ox) * ow - int(ox) * cs)
        assert c0.abs().max().item() < 2 ** 31
        bias = torch.cat(bias)
        return dict(wbits=max(wq["bits"] for wq in wq_list), codes=codes.contiguous(), sxw=sxw, ow=ow.to(torch.int32).contiguous(), c0=c0.to(torch.int32).contiguous(),
                    bias=bias.contiguous() if bias.abs().max().item() > 0 else None, so=so, oo=oo, qgroup=qgroup,
                    qmax=out_q[0][2], N=codes.shape[0], K=Kdim)

    def _build_layer(self, layer, p, qcfg, act):
        L = {}
        at, mlp = layer.self_attn, layer.mlp
        norms = [("n1", layer.input_layernorm, p + ".input_layernorm")]
        if not self.parallel:
            norms.append(("n2", layer.post_attention_layernorm, p + ".post_attention_layernorm"))
        for tag, mod, name in norms:
            wq = self._wq(mod.weight, qcfg[name]["weight"], want_fq=True)
            b = getattr(mod, "bias", None)
            b = None if b is None or float(b.detach().abs().max()) == 0.0 else b.detach().float().to(self.device).contiguous()
            L[tag] = dict(w_fq=wq["w_fq"].reshape(-1).contiguous(), bias=b, qin=_sq(act, qcfg, name, "input"), qout=_sq(act, qcfg, name, "output"),
                          eps=float(getattr(mod, "eps", 1e-5)))
        pa = p + ".self_attn."
        x1 = L["n1"]["qout"]
        qo = [_sq(act, qcfg, pa + n, "output") for n in ("q_proj", "k_proj", "v_proj")]
        L["qkv"] = self._percol([self._wq(m.weight, qcfg[pa + n]["weight"]) for m, n in ((at.q_proj, "q_proj"), (at.k_proj, "k_proj"), (at.v_proj, "v_proj"))],
                                x1[0], x1[1], self.H, qo, [getattr(at.q_proj, "bias", None), getattr(at.k_proj, "bias", None), getattr(at.v_proj, "bias", None)])
        qk_in, qk_in2, qk_out = (_sq(act, qcfg, pa + "qk_bmm", s) for s in ("input", "input2", "output"))
        pv_in, pv_in2, pv_out = (_sq(act, qcfg, pa + "pv_bmm", s) for s in ("input", "input2", "output"))
        # pv_bmm.input_quantizer: p >= 0, so with an offset o_p >= 0 the de-offset code clamp(rne(p/s)+o_p, 0, qmax) - o_p is
        # clamp(rne(p/s), 0, qmax - o_p): the kernels work on de-offset codes with the upper bound lowered by o_p.  A negative
        # offset (learned range minimum above 0) would give every masked position the weight |o_p|*s_p in the reference's
        # fake-quant simulation (it quantises the full [T,T] matrix, hm:527-534) -- not a causal computation; not supported.
        if pv_in[1] < 0.0:
            raise NotImplementedError("pv_bmm.input_quantizer offset must be >= 0 (softmax output range must include 0)")
        pv_in = (pv_in[0], 0.0, pv_in[2] - pv_in[1])
        L["rope_in"] = [(q[0], q[1]) for q in qo]
        L["rope_out"] = [(qk_in[0], qk_in[1]), (qk_in2[0], qk_in2[1]), (pv_in2[0], pv_in2[1])]
        f = np.float32
        L["attn"] = [qk_in[1], qk_in2[1], pv_in2[1], float(f(qk_in[0]) * f(qk_in2[0])), qk_out[0], qk_out[1], qk_out[2], pv_in[0], pv_in[2],
                     float(f(pv_in[0]) * f(pv_in2[0])), pv_out[0], pv_out[1]]
        # two-level exp table of the quantised softmax: E(k) = (A[k >> 8] * B[k & 255]) >> 31 (include/mqb200.h:mq_qattn)
        al = np.float64(f(qk_out[0])) / np.sqrt(np.float64(self.hd))
        i = np.arange(256, dtype=np.float64)
        lut = np.concatenate([np.rint(np.exp(-256.0 * i * al) * 2.0 ** 31), np.rint(np.exp(-i * al) * 2.0 ** 31)]).astype(np.uint32)
        L["attn_lut"] = torch.from_numpy(lut.view(np.int32)).to(self.device)
        L["o"] = self._percol([self._wq(at.o_proj.weight, qcfg[pa + "o_proj"]["weight"])], pv_out[0], pv_out[1], self.nh * self.hd,
                              [_sq(act, qcfg, pa + "o_proj", "output")], [getattr(at.o_proj, "bias", None)])
        # ---- MLP: w1 || w3 interleaved per 128 rows (padded to a multiple of 128), activation folded into a LUT
        pm = p + ".mlp."
        x2 = L["n1"]["qout"] if self.parallel else L["n2"]["qout"]
        q1 = _sq(act, qcfg, pm + "w1", "output")
        w1 = self._wq(mlp.w1.weight, qcfg[pm + "w1"]["weight"])
        pc1 = self._percol([w1], x2[0], x2[1], self.H, [q1], [getattr(mlp.w1, "bias", None)])
        if self.gated:
            q3 = _sq(act, qcfg, pm + "w3", "output")
            w3 = self._wq(mlp.w3.weight, qcfg[pm + "w3"]["weight"])
            pc3 = self._percol([w3], x2[0], x2[1], self.H, [q3], [getattr(mlp.w3, "bias", None)])
        else:
            # two-linear MLP (hm:1057-1062 with num_linears_per_mlp == 2): the gate operand of the fused activation epilogue
            # is the constant 1.0 -- zero weight codes, bias 1, "output quantizer" (scale 1, offset 0): code 1 -> 1.0 exactly
            q3 = (1.0, 0.0, 255.0)
            one = torch.ones(self.I, device=self.device)
            pc3 = dict(pc1, codes=torch.zeros_like(pc1["codes"]), sxw=torch.zeros_like(pc1["sxw"]), ow=torch.zeros_like(pc1["ow"]),
                       c0=torch.zeros_like(pc1["c0"]), bias=one, so=torch.ones_like(pc1["so"]), oo=torch.zeros_like(pc1["oo"]))
        L["w13"] = self._interleave(pc1, pc3)
        act_q = qcfg[pm + "act_fn"]
        q_aout = _sq(act, qcfg, pm + "act_fn", "output")
        c = np.arange(256, dtype=np.float32)
        xg = ((c - f(q1[1])) * f(q1[0])).astype(np.float32)
        if "input2" in act_q:                                   # QSiLU
            q_in2 = _sq(act, qcfg, pm + "act_fn", "input2")
            sg = _f64_sigmoid(xg).astype(np.float32)
            sgq = np.clip(np.rint(sg / f(q_in2[0])).astype(np.float32) + f(q_in2[1]), 0, f(q_in2[2])).astype(np.float32)
            sgq = ((sgq - f(q_in2[1])) * f(q_in2[0])).astype(np.float32)
            a = (xg * sgq).astype(np.float32)
        else:                                                   # QGELU (erf form, qm:794)
            xd = xg.astype(np.float64)
            a = (0.5 * xd * (1.0 + np.vectorize(math.erf)(xd / np.sqrt(2.0)))).astype(np.float32)
        aq = np.clip(np.rint(a / f(q_aout[0])).astype(np.float32) + f(q_aout[1]), 0, f(q_aout[2])).astype(np.float32)
        L["act_lut"] = torch.from_numpy(((aq - f(q_aout[1])) * f(q_aout[0])).astype(np.float32)).to(self.device)
        w2_in = _sq(act, qcfg, pm + "w2", "input")
        L["w2_in"] = w2_in
        w2 = self._wq(mlp.w2.weight, qcfg[pm + "w2"]["weight"])
        L["w2"] = self._percol([self._pad_cols(w2)], w2_in[0], w2_in[1], self.Ipad, [_sq(act, qcfg, pm + "w2", "output")], [getattr(mlp.w2, "bias", None)])
        if self.pack4:
            for k in ("qkv", "o", "w13", "w2"):
                self._pack(L[k])
        return L

    def _pack(self, g):
        """4-bit matrix (already arranged: fused / interleaved / K-padded) -> two UNSIGNED codes per byte (low nibble first); the
        int8 copy is dropped.  Symmetric weights (codes -8..7) are stored in offset-binary form code + 8 with the per-column
        zero point raised by 8 -- c0 = K*ox*ow - ox*colsum is invariant under that shift -- so one format serves the fused
        int4 x int8 GEMM (mq_qgemm_w4a8: nibbles expanded inside the kernel) and the decode path (mq_unpack4)."""
        c = g["codes"]
        if g["wbits"] > 4 or (c.shape[0] * c.shape[1]) % 32 or c.shape[1] % 32:
            return
        u = c.view(torch.uint8)
        if c.dtype == torch.int8:
            u = (u + 8) & 0xF
            g["ow"] = (g["ow"] + 8).contiguous()
        g["packed"] = ((u[:, 0::2] & 0xF) | ((u[:, 1::2] & 0xF) << 4)).contiguous()
        g["shape"], g["dtype"] = tuple(c.shape), torch.uint8
        g["codes"] = None

    def _codes(self, g):
        """The [N, K] one-code-per-byte operand of a GEMM: the resident tensor, or the packed one expanded into scratch."""
        if g["codes"] is not None:
            return g["codes"]
        n = g["shape"][0] * g["shape"][1]
        buf = self._scratch.get(g["dtype"])
        if buf is None or buf.numel() < n:
            nmax = max(L[k]["shape"][0] * L[k]["shape"][1] for L in self.layers for k in ("qkv", "o", "w13", "w2") if L[k]["codes"] is None) \
                if getattr(self, "layers", None) else n
            buf = self._scratch[g["dtype"]] = torch.empty(max(n, nmax), dtype=g["dtype"], device=self.device)
        out = buf[:n].view(g["shape"])
        K.unpack4(g["packed"], out)
        return out

    def weight_bytes(self):
        """Bytes of quantised weight codes resident in HBM."""
        return sum((g["codes"] if g["codes"] is not None else g["packed"]).numel() for L in self.layers for g in (L[k] for k in ("qkv", "o", "w13", "w2")))

    def _pad_cols(self, wq):
        """Pad K (=intermediate) to Ipad with the row's zero point so that padded columns contribute exactly 0."""
        if self.Ipad == self.I:
            return wq
        rows = wq["rows"]
        codes = wq["codes"]
        off = wq["offset"].expand(rows) if wq["offset"].numel() == 1 else wq["offset"]
        pad = off.to(codes.dtype).view(-1, 1).expand(rows, self.Ipad - self.I)
        wq = dict(wq)
        wq["codes"] = torch.cat([codes, pad], dim=1).contiguous()
        wq["colsum"] = (wq["colsum"].to(torch.int64) + off.to(torch.int64) * (self.Ipad - self.I)).to(torch.int32)
        return wq

    def _interleave(self, a, b):
        """[128 rows of w1 | 128 rows of w3] per 256-row GEMM tile; rows beyond I are zero-point padding."""
        I, Ipad = self.I, self.Ipad

        def pad(t, fill):
            if Ipad == I:
                return t
            extra = torch.full((Ipad - I,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
            return torch.cat([t, extra])

        def il(x, y):
            return torch.stack([x.view(Ipad // 128, 128, *x.shape[1:]), y.view(Ipad // 128, 128, *y.shape[1:])], dim=1).reshape(2 * Ipad, *x.shape[1:]).contiguous()

        out = dict(N=2 * Ipad, K=a["K"], qmax=a["qmax"], qgroup=128, wbits=max(a["wbits"], b["wbits"]))
        out["codes"] = il(pad(a["codes"], 0), pad(b["codes"], 0))
        for k, fill in (("sxw", 0.0), ("ow", 0), ("c0", 0)):
            out[k] = il(pad(a[k], fill), pad(b[k], fill))
        # one (so, oo) entry per 128-column half tile: w1's output quantizer, then w3's
        out["so"] = torch.stack([a["so"][:1].expand(Ipad // 128), b["so"][:1].expand(Ipad // 128)], dim=1).reshape(-1).contiguous()
        out["oo"] = torch.stack([a["oo"][:1].expand(Ipad // 128), b["oo"][:1].expand(Ipad // 128)], dim=1).reshape(-1).contiguous()
        if a["bias"] is None and b["bias"] is None:
            out["bias"] = None
        else:
            z = lambda d: d["bias"] if d["bias"] is not None else torch.zeros(I, device=self.device)
            out["bias"] = il(pad(z(a), 0.0), pad(z(b), 0.0))
        return out

    # ---- run --------------------------------------------------------------------------------------------------------
    def _rope(self, T):
        if T not in self._rope_cache:
            from ..model.hf_model import rope_cos_sin
            pos = torch.arange(T, device=self.device).unsqueeze(0)
            cos, sin = rope_cos_sin(pos, self.rot, self.cfg.rope_theta, self.device, torch.float32)
            self._rope_cache[T] = (cos[0].contiguous(), sin[0].contiguous())
        return self._rope_cache[T]

    def set_rope_tables(self, T, cos, sin):
        """Tests pass the oracle's tables so that both sides use bit-identical cos/sin."""
        self._rope_cache[T] = (cos.to(self.device).float().contiguous(), sin.to(self.device).float().contiguous())

    def _buffers(self, B, T):
        key = (B, T)
        if key not in self._bufs:
            M, dev = B * T, self.device
            u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
            i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
            self._bufs[key] = dict(x=u8(M, self.H), rs=i32(M), qkv=u8(M, (self.nh + 2 * self.nkv) * self.hd),
                                   rope=dict(q=u8(B, self.nh, T, self.hd), k=u8(B, self.nkv, T, self.hd), vt=u8(B, self.nkv, self.hd, T),
                                             rsq=i32(B, self.nh, T), rsk=i32(B, self.nkv, T)),
                                   attn=u8(M, self.nh * self.hd), rs_attn=i32(M), act=u8(M, self.Ipad), rs_act=i32(M))
        return self._bufs[key]

    def _gemm(self, a, g, rowsum, mode, **kw):
        if g["codes"] is None and self.fused_w4:        # packed 4-bit weights: expanded inside the GEMM (no scratch copy)
            return K.qgemm(a, g["packed"], rowsum, g["sxw"], g["ow"], g["c0"], mode, bias=g["bias"], so=g["so"], oo=g["oo"], qmax=g["qmax"],
                           qgroup=g["qgroup"], packed4=True, **kw)
        return K.qgemm(a, self._codes(g), rowsum, g["sxw"], g["ow"], g["c0"], mode, bias=g["bias"], so=g["so"], oo=g["oo"], qmax=g["qmax"],
                       qgroup=g["qgroup"], **kw)

    def _attn_inputs(self, h, L, B, T, bufs, cos, sin):
        """Input norm, fused q|k|v projection, RoPE + re-quantisation into the attention layouts (token-local work)."""
        K.qnorm(h, L["n1"]["qin"], L["n1"]["w_fq"], L["n1"]["bias"], L["n1"]["qout"], self.layernorm, L["n1"]["eps"], bufs["x"], bufs["rs"])
        self._gemm(bufs["x"], L["qkv"], bufs["rs"], K.EPI_QUANT, out=bufs["qkv"], out_bits=8)
        K.qrope(bufs["qkv"], B, T, self.nh, self.nkv, self.hd, self.rot, L["rope_in"], L["rope_out"], cos, sin, bufs["rope"])

    def _block_tail(self, h, L, bufs, trace=None):
        """o_proj + residual, (post-attention norm,) MLP + residual.  Sequential block: the MLP reads the norm of the updated
        residual stream; parallel block with shared norm (phi): it reads the input norm's codes, still in bufs["x"]."""
        self._gemm(bufs["attn"], L["o"], bufs["rs_attn"], K.EPI_RESID, resid=h)
        if trace is not None:
            trace.update(h_mid=h.clone())
        if not self.parallel:
            K.qnorm(h, L["n2"]["qin"], L["n2"]["w_fq"], L["n2"]["bias"], L["n2"]["qout"], self.layernorm, L["n2"]["eps"], bufs["x"], bufs["rs"])
        bufs["rs_act"].zero_()
        w2in = L["w2_in"]
        self._gemm(bufs["x"], L["w13"], bufs["rs"], K.EPI_ACTMUL, out=bufs["act"], lut=L["act_lut"], s2=w2in[0], o2=w2in[1], qmax2=w2in[2],
                   rowsum_out=bufs["rs_act"])
        if trace is not None:
            trace.update(x2=bufs["x"].clone(), act=bufs["act"].clone())
        self._gemm(bufs["act"], L["w2"], bufs["rs_act"], K.EPI_RESID, resid=h)
        return h

    @torch.no_grad()
    def block(self, h, L, B, T, bufs, trace=None, cache=None):
        """One decoder block on the fp32 residual stream h [B*T, H] (updated in place).  cache = (KVCache, layer index):
        the rotated k / v codes of the T tokens are also written to the decode cache."""
        cos, sin = self._rope(T)
        self._attn_inputs(h, L, B, T, bufs, cos, sin)
        if cache is not None:
            kv, li = cache
            kv.k[li][:, :, :T].copy_(bufs["rope"]["k"])
            kv.v[li][:, :, :T].copy_(bufs["rope"]["vt"].transpose(2, 3))
            kv.rsk[li][:, :, :T].copy_(bufs["rope"]["rsk"])
        bufs["rs_attn"].zero_()
        K.qattn(bufs["rope"], B, T, self.nh, self.nkv, self.hd, L["attn"], L["attn_lut"], bufs["attn"], bufs["rs_attn"])
        if trace is not None:
            trace.update(x1=bufs["x"].clone(), qkv=bufs["qkv"].clone(), q=bufs["rope"]["q"].clone(), k=bufs["rope"]["k"].clone(),
                         vt=bufs["rope"]["vt"].clone(), attn=bufs["attn"].clone())
        return self._block_tail(h, L, bufs, trace)

    # ---- sequence-sharded prefill (north star: "inference shards the KV/sequence across GPUs") -----------
    @staticmethod
    def seq_shard_plan(T, rank, world):
        """Zig-zag ownership of 2*world equal chunks: rank r owns chunks r and 2*world-1-r, which gives every rank the same
        number of causally visible (query, key) pairs.  Returns (chunk length, [chunk ids], absolute positions of the local tokens)."""
        if T % (2 * world * 128):
            raise ValueError(f"sequence-sharded prefill needs T % (256 * world) == 0 (T {T}, world {world})")
        Tc = T // (2 * world)
        chunks = [rank, 2 * world - 1 - rank]
        pos = torch.cat([torch.arange(c * Tc, (c + 1) * Tc) for c in chunks])
        return Tc, chunks, pos

    def _shard_buffers(self, B, Tl, T):
        dev = self.device
        u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=dev)
        b = dict(x=u8(B * Tl, self.H), rs=i32(B * Tl), qkv=u8(B * Tl, (self.nh + 2 * self.nkv) * self.hd),
                 rope=dict(q=u8(B, self.nh, Tl, self.hd), k=u8(B, self.nkv, Tl, self.hd), vt=u8(B, self.nkv, self.hd, Tl),
                           rsq=i32(B, self.nh, Tl), rsk=i32(B, self.nkv, Tl)),
                 attn=u8(B * Tl, self.nh * self.hd), rs_attn=i32(B * Tl), act=u8(B * Tl, self.Ipad), rs_act=i32(B * Tl),
                 k_all=u8(B, self.nkv, T, self.hd), vt_all=u8(B, self.nkv, self.hd, T), rsk_all=i32(B, self.nkv, T))
        return b

    @torch.no_grad()
    def seq_sharded_steps(self, input_ids, rank, world):
        """Generator form of the sequence-sharded prefill of rank `rank`: per layer it yields the packed K | V^T | key-code-sum
        codes of the local tokens (one uint8 tensor, 2*hd + 4 bytes per token and kv head) and expects the list of all ranks'
        packed tensors back (`send`); everything else is token-local.  Finishes by returning the fp32 hidden state of the
        local tokens [B, Tl, H] and their absolute positions.  Drivers: prefill_seq_sharded (torch.distributed all_gather) and
        the single-process lock-step simulation of the tests."""
        B, T = input_ids.shape
        if self.hd not in (64, 128):
            raise NotImplementedError("sequence-sharded attention runs on the tcgen05 kernel (head_dim 64 / 128)")
        Tc, chunks, pos = self.seq_shard_plan(T, rank, world)
        Tl = 2 * Tc
        pos_d = pos.to(self.device)
        h = self._embed(input_ids.to(self.device)[:, pos_d]).reshape(B * Tl, self.H).contiguous()
        cos, sin = self._rope(T)
        cosl, sinl = cos[pos_d].contiguous(), sin[pos_d].contiguous()
        bufs = self._shard_buffers(B, Tl, T)
        rope = bufs["rope"]
        nk, nr = rope["k"].numel(), rope["rsk"].numel() * 4
        for L in self.layers:
            self._attn_inputs(h, L, B, Tl, bufs, cosl, sinl)
            packed = torch.cat([rope["k"].reshape(-1), rope["vt"].reshape(-1), rope["rsk"].view(torch.uint8).reshape(-1)])
            gathered = yield packed
            for r, buf in enumerate(gathered):                        # place every rank's two chunks at their sequence positions
                k_r = buf[:nk].view(B, self.nkv, Tl, self.hd)
                v_r = buf[nk:2 * nk].view(B, self.nkv, self.hd, Tl)
                s_r = buf[2 * nk:2 * nk + nr].view(torch.int32).view(B, self.nkv, Tl)
                for s, c in enumerate((r, 2 * world - 1 - r)):
                    bufs["k_all"][:, :, c * Tc:(c + 1) * Tc].copy_(k_r[:, :, s * Tc:(s + 1) * Tc])
                    bufs["vt_all"][:, :, :, c * Tc:(c + 1) * Tc].copy_(v_r[:, :, :, s * Tc:(s + 1) * Tc])
                    bufs["rsk_all"][:, :, c * Tc:(c + 1) * Tc].copy_(s_r[:, :, s * Tc:(s + 1) * Tc])
            attn3 = bufs["attn"].view(B, Tl, self.nh * self.hd)
            rs2 = bufs["rs_attn"].view(B, Tl)
            for s, c in enumerate(chunks):                             # one causal call per owned chunk: queries c*Tc .. (c+1)*Tc-1
                q_s = rope["q"][:, :, s * Tc:(s + 1) * Tc].contiguous()
                rsq_s = rope["rsq"][:, :, s * Tc:(s + 1) * Tc].contiguous()
                rs_s = torch.zeros(B * Tc, dtype=torch.int32, device=self.device)
                out_s = K.qattn_shard(q_s, rsq_s, bufs["k_all"], bufs["vt_all"], bufs["rsk_all"], B, Tc, T, c * Tc, self.nh, self.nkv,
                                      self.hd, L["attn"], L["attn_lut"], rowsum_out=rs_s)
                attn3[:, s * Tc:(s + 1) * Tc].copy_(out_s.view(B, Tc, -1))
                rs2[:, s * Tc:(s + 1) * Tc].copy_(rs_s.view(B, Tc))
            self._block_tail(h, L, bufs)
        return h.view(B, Tl, self.H), pos

    @torch.no_grad()
    def prefill_seq_sharded(self, input_ids, group=None):
        """Sequence-sharded integer prefill over the ranks of a torch.distributed (NCCL) process group: every rank holds the
        weights, owns 2 of 2*world sequence chunks (zig-zag) and exchanges only the int8 K / V codes of its tokens, one
        all_gather per layer.  Returns (hidden [B, Tl, H] of the local tokens, their absolute positions); the logits of the
        last position live on rank 0 (it owns the last chunk): `last_token_logits`."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        gen = self.seq_sharded_steps(input_ids, rank, world)
        packed = next(gen)
        try:
            while True:
                out = [torch.empty_like(packed) for _ in range(world)]
                dist.all_gather(out, packed, group=group)
                packed = gen.send(out)
        except StopIteration as e:
            return e.value

    def last_token_logits(self, h_local, pos):
        """Logits of the sequence's last position from a rank's local hidden state (None on ranks that do not own it)."""
        T_last = int(pos.max().item())
        idx = int((pos == T_last).nonzero()[0])
        return self._head(h_local[:, idx, :])

    @torch.no_grad()
    def backbone(self, h, B, T, trace_layer=None, cache=None):
        bufs = self._buffers(B, T)
        trace = None
        for i, L in enumerate(self.layers):
            tr = {} if trace_layer == i else None
            self.block(h, L, B, T, bufs, tr, None if cache is None else (cache, i))
            if tr is not None:
                trace = tr
        return (h, trace) if trace_layer is not None else h

    @torch.no_grad()
    def forward(self, input_ids, return_logits=True, last_token_only=False):
        """input_ids: LongTensor [B, T] on the engine's device.  Returns logits [B, T, V] (or the final hidden state)."""
        B, T = input_ids.shape
        h = torch.nn.functional.embedding(input_ids, self.embed)
        if self.cfg.normalize_embed:
            h = h * (self.H ** 0.5)
        h = h.reshape(B * T, self.H).contiguous()
        h = self.backbone(h, B, T).view(B, T, self.H)
        if not return_logits:
            return h
        if last_token_only:
            h = h[:, -1:, :]
        hn = self.final_norm(h)
        return torch.nn.functional.linear(hn, self.lm_head, self.lm_head_bias)

    __call__ = forward

    # ---- decode against an int8 KV cache (reference: SimModel.generate, sim_model.py:181-235; capp/src/llm.cpp:545-653) --
    def new_cache(self, B, Tmax):
        return KVCache(len(self.layers), B, self.nkv, Tmax, self.hd, self.device)

    def _embed(self, ids):
        h = torch.nn.functional.embedding(ids, self.embed)
        if self.cfg.normalize_embed:
            h = h * (self.H ** 0.5)
        return h

    def _head(self, h):
        hn = self.final_norm(h)
        if hn.dim() == 2 and hn.shape[0] <= 16:        # decode step: HBM-bound fp32 GEMV instead of a library SGEMM
            out = K.fgemv(hn.contiguous(), self.lm_head)
            return out if self.lm_head_bias is None else out.add_(self.lm_head_bias)
        return torch.nn.functional.linear(hn, self.lm_head, self.lm_head_bias)

    @torch.no_grad()
    def prefill(self, input_ids, cache):
        """Context encoding: full integer forward over input_ids [B, T] that also fills the cache; returns the logits of
        the last position [B, V] (sim_model.py:195-212)."""
        B, T = input_ids.shape
        assert B == cache.B and T <= cache.Tmax
        h = self._embed(input_ids).reshape(B * T, self.H).contiguous()
        h = self.backbone(h, B, T, cache=cache).view(B, T, self.H)
        cache.length = T
        cache.pos_dev.fill_(T)
        return self._head(h[:, -1, :])

    def _decode_bufs(self, B):
        key = ("dec", B)
        if key not in self._bufs:
            dev = self.device
            u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
            nmax = max(L[k]["N"] for L in self.layers for k in ("qkv", "o", "w13", "w2"))
            self._bufs[key] = dict(x=u8(B, self.H), rs=torch.empty(B, dtype=torch.int32, device=dev), qkv=u8(B, (self.nh + 2 * self.nkv) * self.hd),
                                   attn=u8(B, self.nh * self.hd), rs_attn=torch.zeros(B, dtype=torch.int32, device=dev), act=u8(B, self.Ipad),
                                   rs_act=torch.zeros(B, dtype=torch.int32, device=dev), acc=torch.zeros(B, nmax, dtype=torch.int32, device=dev))
        return self._bufs[key]

    def _norm_dec(self, h, n, bufs, pending):
        """Row norm of the decode step; `pending` = (GEMM, row sums) whose residual epilogue is applied to h first, in the
        same launch."""
        if pending is None:
            return K.qnorm(h, n["qin"], n["w_fq"], n["bias"], n["qout"], self.layernorm, n["eps"], bufs["x"], bufs["rs"])
        g, rs = pending
        return K.qnorm_resid(h, n["qin"], n["w_fq"], n["bias"], n["qout"], self.layernorm, n["eps"], bufs["x"], bufs["rs"], bufs["acc"], g, rs)

    def _gemv(self, a, g, rowsum, mode, acc, **kw):
        if self.fused_gemv:
            return K.qgemv_fused(a, self._codes(g), acc, rowsum, g["sxw"], g["ow"], g["c0"], mode, bias=g["bias"], so=g["so"], oo=g["oo"],
                                 qmax=g["qmax"], qgroup=g["qgroup"], **kw)
        K.qgemv(a, self._codes(g), acc)
        return K.qgemv_epilogue(acc, a.shape[0], g["N"], rowsum, g["sxw"], g["ow"], g["c0"], mode, bias=g["bias"], so=g["so"], oo=g["oo"],
                                qmax=g["qmax"], qgroup=g["qgroup"], **kw)

    @torch.no_grad()
    def decode_hidden(self, h, cache, use_pos_dev=False):
        """One decode step on the fp32 residual rows h [B, H] of the new tokens (position cache.length); updated in place."""
        B = h.shape[0]
        bufs = self._decode_bufs(B)
        cos, sin = self._rope(cache.Tmax)
        pos = cache.length
        pkw = dict(pos_dev=cache.pos_dev, pos_bound=cache.Tmax - 1) if use_pos_dev else {}
        fuse = self.fused_resid_norm and not self.parallel
        pending = None                      # (GEMM, its row sums): a residual epilogue still to be applied to h
        for i, L in enumerate(self.layers):
            self._norm_dec(h, L["n1"], bufs, pending)
            # code-sum buffers are cleared by an epilogue that runs between their consumer and their next producer
            self._gemv(bufs["x"], L["qkv"], bufs["rs"], K.EPI_QUANT, bufs["acc"], out=bufs["qkv"], zero_out=bufs["rs_act"])
            K.qattn_decode(bufs["qkv"], B, self.nh, self.nkv, self.hd, self.rot, pos, L["rope_in"], L["rope_out"], cos, sin, cache.k[i], cache.v[i],
                           cache.rsk[i], L["attn"], L["attn_lut"], out=bufs["attn"], rowsum_out=bufs["rs_attn"], **pkw)
            if fuse:
                K.qgemv(bufs["attn"], self._codes(L["o"]), bufs["acc"])
                self._norm_dec(h, L["n2"], bufs, (L["o"], bufs["rs_attn"]))
            else:
                self._gemv(bufs["attn"], L["o"], bufs["rs_attn"], K.EPI_RESID, bufs["acc"], resid=h)
                if not self.parallel:                # (parallel block: the MLP reads the input norm's codes, still in bufs["x"])
                    self._norm_dec(h, L["n2"], bufs, None)
            w2in = L["w2_in"]
            self._gemv(bufs["x"], L["w13"], bufs["rs"], K.EPI_ACTMUL, bufs["acc"], out=bufs["act"], lut=L["act_lut"], s2=w2in[0], o2=w2in[1],
                       qmax2=w2in[2], rowsum_out=bufs["rs_act"], zero_out=bufs["rs_attn"])
            if fuse:
                K.qgemv(bufs["act"], self._codes(L["w2"]), bufs["acc"])
                pending = (L["w2"], bufs["rs_act"])
            else:
                self._gemv(bufs["act"], L["w2"], bufs["rs_act"], K.EPI_RESID, bufs["acc"], resid=h)
        if pending is not None:             # the last layer's w2 epilogue has no norm to ride on
            g, rs = pending
            K.qgemv_epilogue(bufs["acc"], B, g["N"], rs, g["sxw"], g["ow"], g["c0"], K.EPI_RESID, bias=g["bias"], so=g["so"], oo=g["oo"],
                             qmax=g["qmax"], qgroup=g["qgroup"], resid=h)
        return h

    @torch.no_grad()
    def decode_step(self, tokens, cache):
        """tokens: LongTensor [B] (the tokens at position cache.length).  Returns next-token logits [B, V]."""
        assert cache.length < cache.Tmax, "KV cache is full"
        h = self._embed(tokens.view(-1)).contiguous()
        self.decode_hidden(h, cache)
        cache.length += 1
        cache.pos_dev.fill_(cache.length)
        return self._head(h)

    @torch.no_grad()
    def capture_decode(self, cache):
        """CUDA graph of one greedy decode step: token buffer -> logits -> argmax -> token buffer, position += 1 on the
        device.  Returns (graph, tokens, logits); each replay advances every sequence of the cache by one token."""
        B = cache.B
        tokens = torch.zeros(B, dtype=torch.long, device=self.device)
        state = dict(logits=None)

        def step():
            h = self._embed(tokens).contiguous()
            self.decode_hidden(h, cache, use_pos_dev=True)
            state["logits"] = self._head(h)
            tokens.copy_(state["logits"].argmax(-1))
            cache.pos_dev.add_(1)

        # warm-up outside the capture (function attributes, workspaces), then restore the position and the scratch state
        keep = cache.length
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        cache.pos_dev.fill_(keep)
        tokens.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        cache.pos_dev.fill_(keep)
        return DecodeGraph(g, tokens, state["logits"], cache)

    @torch.no_grad()
    def generate(self, context_ids, max_new_tokens, eos_token_id=None, pad_token_id=0, do_sample=False, temperature=0.5, Tmax=None):
        """Greedy / sampled generation with the int8 KV cache; same contract as SimModel.generate (sim_model.py:181-235):
        context_ids LongTensor [B, context_len] -> LongTensor [B, context_len + generated] (stops once every sequence has
        emitted an eos token; finished sequences are padded with pad_token_id)."""
        context_ids = context_ids.to(self.device)
        B, T0 = context_ids.shape
        Tmax = Tmax or (T0 + max_new_tokens)
        assert T0 < Tmax and T0 + max_new_tokens <= Tmax
        eos = [eos_token_id] if isinstance(eos_token_id, int) else list(eos_token_id or [])
        cache = self.new_cache(B, Tmax)
        logits = self.prefill(context_ids, cache)
        out = [context_ids]
        done = torch.zeros(B, dtype=torch.bool, device=self.device)
        for i in range(max_new_tokens):
            if do_sample:
                nxt = torch.multinomial(torch.softmax(logits / temperature, dim=-1), num_samples=1).view(-1)
            else:
                nxt = logits.argmax(-1)
            nxt = torch.where(done, torch.full_like(nxt, pad_token_id), nxt)
            out.append(nxt.view(B, 1))
            for e in eos:
                done |= nxt == e
            if bool(done.all()) or i + 1 == max_new_tokens or cache.length >= cache.Tmax:
                break
            logits = self.decode_step(nxt, cache)
        return torch.cat(out, dim=1)


class DecodeGraph:
    """The captured greedy decode step + its static buffers.  replay() refuses to step past the cache capacity (the device
    position is incremented by the graph itself) and keeps cache.length in step with the device position.  Unpacks as
    (graph, tokens, logits) for callers that only need the buffers."""

    def __init__(self, graph, tokens, logits, cache):
        self.graph, self.tokens, self.logits, self.cache = graph, tokens, logits, cache

    def replay(self):
        if self.cache.length >= self.cache.Tmax:
            raise RuntimeError(f"KV cache is full ({self.cache.Tmax} positions): cannot decode another token")
        self.graph.replay()
        self.cache.length += 1

    def rewind(self, length):
        """Set the host and device position (benchmarks replay the same steps several times)."""
        self.cache.length = int(length)
        self.cache.pos_dev.fill_(int(length))

    def __iter__(self):
        return iter((self, self.tokens, self.logits))


class KVCache:
    """uint8 key / value codes of every layer, k and v [B, nkv, Tmax, hd] (the reference keeps [L, n_heads, T-1, head_dim],
    sim_model.py:118-119), plus the per-key code sums the zero-point correction of the scores needs."""

    def __init__(self, n_layers, B, nkv, Tmax, hd, device):
        self.B, self.Tmax, self.length = B, Tmax, 0
        u8 = lambda: torch.zeros(B, nkv, Tmax, hd, dtype=torch.uint8, device=device)
        self.k = [u8() for _ in range(n_layers)]
        self.v = [u8() for _ in range(n_layers)]
        self.rsk = [torch.zeros(B, nkv, Tmax, dtype=torch.int32, device=device) for _ in range(n_layers)]
        self.pos_dev = torch.zeros(1, dtype=torch.int32, device=device)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.k + self.v + self.rsk)
