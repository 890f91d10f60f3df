"""TEST INFRASTRUCTURE ONLY -- exact-integer CPU restatement of the statically-quantised forward (the arithmetic the
sm_100a engine kernels implement), written with numpy int64 for the integer parts and numpy float32 for the
requantisation steps (IEEE fp32 division, np.rint = round-half-to-even), each op in the same order as the kernels.

It restates the reference's fake-quant modules on their integer codes:
  qlinear_int   <-> QLinear.forward            mobilellm/quantization/qmodule.py:341-358
  (more functions are added next to each engine kernel: norm, rope, attention, act)
The fp32 fake-quant reference accumulates its GEMMs in fp32, so its codes can differ from the exact-integer result by
one LSB at rounding ties; tests/ measure that flip rate against the goldens, and check the kernels bit-exactly against
this file.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.
"""
import numpy as np

f32 = np.float32


def quant_codes(y, scale, offset, qmin, qmax):
    """clamp(rne(y / s) + o, qmin, qmax) in fp32 (qm:286-287); returns float32 array holding integers."""
    y = np.asarray(y, dtype=f32)
    q = np.rint(y / f32(scale) if np.ndim(scale) == 0 else y / np.asarray(scale, f32)).astype(f32) + np.asarray(offset, f32)
    return np.clip(q, f32(qmin), f32(qmax)).astype(f32)


def dequant(q, scale, offset):
    """(q - o) * s in fp32 (qm:290)"""
    return ((np.asarray(q, f32) - np.asarray(offset, f32)).astype(f32) * np.asarray(scale, f32)).astype(f32)


def int_acc(a_codes, b_codes, ox, ow):
    """sum_k (A - ox)(B - ow[n]) exactly, int64.  a: [M,K], b: [N,K], ow: [N] or scalar."""
    a = a_codes.astype(np.int64) - np.int64(ox)
    b = b_codes.astype(np.int64) - np.asarray(ow, np.int64).reshape(-1, 1)
    return a @ b.T


def qlinear_y(a_codes, sx, ox, b_codes, sw, ow, bias=None):
    """Pre-requant output of the integer linear: y = float(I) * (sx*sw[n]) (+ bias[n]), all fp32 roundings explicit."""
    I = int_acc(a_codes, b_codes, ox, ow)
    sxw = (f32(sx) * np.asarray(sw, f32)).astype(f32).reshape(1, -1)
    y = (I.astype(f32) * sxw).astype(f32)
    if bias is not None:
        y = (y + np.asarray(bias, f32).reshape(1, -1)).astype(f32)
    else:
        y = (y + f32(0)).astype(f32)
    return y


def qlinear_int(a_codes, sx, ox, b_codes, sw, ow, so, oo, qmax, bias=None):
    """Integer QLinear: returns output codes (int64)."""
    y = qlinear_y(a_codes, sx, ox, b_codes, sw, ow, bias)
    return quant_codes(y, np.asarray(so, f32).reshape(1, -1) if np.ndim(so) else so,
                       np.asarray(oo, f32).reshape(1, -1) if np.ndim(oo) else oo, 0, qmax).astype(np.int64)


# =================================================================================================================
# The rest of the integer forward.  Shared conventions: codes are int64 arrays, (s, o, qmax) triples describe static
# per-tensor activation quantizers (o integral), every fp32 op is written out in kernel order.
# =================================================================================================================
def _mm_exact(a, b):
    """Exact integer matmul through float64 BLAS (all |partial sums| < 2^53 for 8/16-bit codes and K <= 2^20)."""
    return np.rint(a.astype(np.float64) @ b.astype(np.float64)).astype(np.int64)


def weight_quant(w, bits, symmetric, per_channel):
    """Static weight quantizer on its first forward (qm:262-290): codes, scale[g], offset[g] in fp32 arithmetic."""
    w = np.asarray(w, f32)
    w2 = w.reshape(1, -1) if w.ndim == 1 else w
    if per_channel:
        mn, mx = w2.min(axis=1, keepdims=True), w2.max(axis=1, keepdims=True)
    else:
        mn, mx = w2.min().reshape(1, 1), w2.max().reshape(1, 1)
    if symmetric:
        qmin, qmax = -2 ** (bits - 1), 2 ** (bits - 1) - 1
        alpha = np.maximum(np.abs(mn), np.abs(mx)).astype(f32)
        scale = np.clip((alpha / f32(qmax)).astype(f32), f32(1e-5), f32(1e6)).astype(f32)
        offset = np.zeros_like(scale)
    else:
        qmin, qmax = 0, 2 ** bits - 1
        alpha = (mx - mn).astype(f32)
        scale = np.clip((alpha / f32(qmax)).astype(f32), f32(1e-5), f32(1e6)).astype(f32)
        offset = (-np.rint((mn / scale).astype(f32))).astype(f32) + f32(0)
    codes = np.clip(np.rint((w2 / scale).astype(f32)).astype(f32) + offset, f32(qmin), f32(qmax)).astype(f32)
    return codes.astype(np.int64), scale.reshape(-1), offset.reshape(-1), qmin, qmax


def qnorm_int(x, qin, w_fq, bias, qout, layernorm=False, eps=1e-5):
    """mq_qnorm: x fp32 [M,H] -> u8 codes.  qin/qout = (s, o, qmax)."""
    x = np.asarray(x, f32)
    s_in, o_in, qmax_in = f32(qin[0]), f32(qin[1]), f32(qin[2])
    q = quant_codes(x, s_in, o_in, 0, qmax_in)
    r = (q - o_in).astype(f32)
    ri = r.astype(np.int64)
    H = x.shape[-1]
    xh = (r * s_in).astype(f32)
    if layernorm:
        s1 = ri.sum(-1, keepdims=True).astype(np.float64); s2 = (ri * ri).sum(-1, keepdims=True).astype(np.float64)
        m = s1 / H
        var = s2 / H - m * m
        mean = (m * np.float64(s_in)).astype(f32)
        rstd = (1.0 / np.sqrt(var * np.float64(s_in) * np.float64(s_in) + np.float64(f32(eps)))).astype(f32)
        t = (((xh - mean).astype(f32) * rstd).astype(f32) * np.asarray(w_fq, f32)).astype(f32)
    else:
        s2 = (ri * ri).sum(-1, keepdims=True)
        nrm = (np.sqrt(s2.astype(np.uint64).astype(f32)).astype(f32) * s_in).astype(f32)
        den = np.maximum(nrm, f32(1e-12))
        alpha = f32(np.sqrt(np.float64(H)))
        t = (np.asarray(w_fq, f32) * (alpha * (xh / den).astype(f32)).astype(f32)).astype(f32)
    if bias is not None:
        t = (t + np.asarray(bias, f32)).astype(f32)
    return quant_codes(t, qout[0], qout[1], 0, qout[2]).astype(np.int64)


def rope_tables(T, rot, base=10000.0):
    """cos/sin [T, rot] exactly as hm:308-318 computes them in fp32 (torch is used so the bits match the product's
    tables when both are built on the CPU; the kernels take the tables as input)."""
    import torch
    inv_freq = 1.0 / (base ** (torch.arange(0, rot, 2, dtype=torch.int64).float() / rot))
    pos = torch.arange(T)[None]
    freqs = (inv_freq[None, :, None].float().expand(1, -1, 1) @ pos[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)[0]
    return emb.cos().numpy().astype(f32), emb.sin().numpy().astype(f32)


def qrope_int(qkv_codes, B, T, nh, nkv, hd, rot, qin, qout, cos, sin):
    """mq_qrope.  qin = [(s,o)]*3 projection output quantizers, qout = [(s,o)]*3 bmm input quantizers (8 bit).
    Returns q [B,nh,T,hd], k [B,nkv,T,hd], v [B,nkv,T,hd] codes."""
    c = np.asarray(qkv_codes).astype(f32).reshape(B, T, -1)
    segs = [(0, nh), (nh * hd, nkv), (nh * hd + nkv * hd, nkv)]
    outs = []
    for i, (off, n) in enumerate(segs):
        x = dequant(c[:, :, off:off + n * hd].reshape(B, T, n, hd), qin[i][0], qin[i][1])
        if i < 2 and rot > 0:
            half = rot // 2
            xr = x[..., :rot]
            rh = np.concatenate([-xr[..., half:], xr[..., :half]], axis=-1)
            cs, sn = cos[None, :, None, :], sin[None, :, None, :]
            rotd = ((xr * cs).astype(f32) + (rh * sn).astype(f32)).astype(f32)
            x = np.concatenate([rotd, x[..., rot:]], axis=-1)
        outs.append(quant_codes(x, qout[i][0], qout[i][1], 0, 255).astype(np.int64).transpose(0, 2, 1, 3))
    return outs


def exp_tables(s_s, hd):
    """Two-level exp table of the quantised softmax: A[i] = rne(2^31 exp(-256 i a)), B[j] = rne(2^31 exp(-j a)),
    a = s_s / sqrt(hd) in float64; E(k) = (A[k >> 8] * B[k & 255]) >> 31 ~= 2^31 exp(-k a) for 16-bit k.
    Returns one uint32[512] array (A then B) -- the `lut` argument of mq_qattn."""
    a = np.float64(f32(s_s)) / np.sqrt(np.float64(hd))
    i = np.arange(256, dtype=np.float64)
    A = np.rint(np.exp(-256.0 * i * a) * 2.0 ** 31).astype(np.uint32)
    B = np.rint(np.exp(-i * a) * 2.0 ** 31).astype(np.uint32)
    return np.concatenate([A, B])


def exp_eval(tab, k):
    """E(k) for an int64 array k in [0, 65535] (uint64)."""
    k = np.asarray(k, np.int64)
    return (tab[k >> 8].astype(np.uint64) * tab[256 + (k & 255)].astype(np.uint64)) >> np.uint64(31)


def qattn_int(q, k, v, nh, nkv, qq, qk, qv, qs, qp, qo, lut=None, heads=None, q_start=0):
    """mq_qattn.  q [B,nh,Tq,hd], k/v [B,nkv,T,hd] codes; qq/qk/qv = (s,o); qs = (s,o,qmax) score quantizer;
    qp = (s, o(=0), qmax) prob quantizer; qo = (s,o) output quantizer (8 bit).  Returns codes [B*Tq, nh*hd].
    heads: optional iterable of (b, h) pairs -- only those are computed (the rest of the output stays 0; full-size
    parity tests check a few heads against this oracle and the remaining ones against another kernel).
    q_start: absolute position of query row 0 (sequence-sharded prefill: keys 0 .. q_start + Tq - 1 are visible)."""
    B, _, Tq, hd = q.shape
    T = k.shape[2]
    rep = nh // nkv
    if lut is None:
        lut = exp_tables(qs[0], hd)
    # de-offset prob codes: clamp(rne(p/s)+o_p, 0, qmax) - o_p == clamp(rne(p/s), 0, qmax - o_p) for p >= 0, o_p >= 0
    assert int(qp[1]) >= 0
    qp = (qp[0], 0, f32(qp[2]) - f32(qp[1]))
    sqk = f32(f32(qq[0]) * f32(qk[0])); spv = f32(f32(qp[0]) * f32(qv[0]))
    out = np.zeros((B, Tq, nh, hd), np.int64)
    causal = np.arange(T)[None, :] <= (q_start + np.arange(Tq))[:, None]
    todo = set((b, h) for b in range(B) for h in range(nh)) if heads is None else set(heads)
    for b in range(B):
        for h in range(nh):
            if (b, h) not in todo:
                continue
            kv = h // rep
            I = _mm_exact(q[b, h] - int(qq[1]), (k[b, kv] - int(qk[1])).T)
            c = quant_codes((I.astype(f32) * sqk).astype(f32), qs[0], qs[1], 0, qs[2]).astype(np.int64)
            cmax = np.where(causal, c, -1).max(axis=1, keepdims=True)
            E = np.where(causal, exp_eval(lut, np.clip(cmax - c, 0, int(qs[2]))), np.uint64(0))
            S = E.sum(axis=1, keepdims=True, dtype=np.uint64)
            p = (E.astype(f32) / S.astype(f32)).astype(f32)
            cp = np.where(causal, quant_codes(p, qp[0], 0, 0, qp[2]).astype(np.int64), 0)
            A = _mm_exact(cp, v[b, kv] - int(qv[1]))
            out[b, :, h, :] = quant_codes((A.astype(f32) * spv).astype(f32), qo[0], qo[1], 0, 255).astype(np.int64)
    return out.reshape(B * Tq, nh * hd)


def qattn_decode_int(q, k, v, nh, nkv, qq, qk, qv, qs, qp, qo, lut=None):
    """mq_qattn_decode (attention part): the new token's row against every cached key, SimAttention.forward with k_cache /
    v_cache (mobilellm/model/sim_model.py:271-330; the causal mask row of a decode step hides nothing, :222-223).
    q [B,nh,hd] codes of the new token, k / v [B,nkv,Tk,hd] codes of positions 0..Tk-1 (the new token included).
    Same arithmetic as qattn_int restricted to one query row.  Returns codes [B, nh*hd]."""
    B, _, hd = q.shape
    rep = nh // nkv
    if lut is None:
        lut = exp_tables(qs[0], hd)
    assert int(qp[1]) >= 0
    qp = (qp[0], 0, f32(qp[2]) - f32(qp[1]))
    sqk = f32(f32(qq[0]) * f32(qk[0])); spv = f32(f32(qp[0]) * f32(qv[0]))
    out = np.zeros((B, nh, hd), np.int64)
    for b in range(B):
        for h in range(nh):
            kv = h // rep
            I = _mm_exact((q[b, h] - int(qq[1])).reshape(1, hd), (k[b, kv] - int(qk[1])).T)
            c = quant_codes((I.astype(f32) * sqk).astype(f32), qs[0], qs[1], 0, qs[2]).astype(np.int64)
            E = exp_eval(lut, np.clip(c.max() - c, 0, int(qs[2])))
            S = E.sum(dtype=np.uint64)
            p = (E.astype(f32) / f32(S)).astype(f32)
            cp = quant_codes(p, qp[0], 0, 0, qp[2]).astype(np.int64)
            A = _mm_exact(cp, v[b, kv] - int(qv[1]))
            out[b, h] = quant_codes((A.astype(f32) * spv).astype(f32), qo[0], qo[1], 0, 255).astype(np.int64)[0]
    return out.reshape(B, nh * hd)


def act_lut(kind, q_w1out, q_in2, q_out):
    """256-entry table: w1-output code -> fq_out(act(x^)) as fp32 (QSiLU qm:739-753 / QGELU qm:790-799).  The
    transcendental is evaluated in float64 and rounded once to fp32."""
    from math import erf
    c = np.arange(256, dtype=f32)
    x = dequant(c, q_w1out[0], q_w1out[1])
    if kind == "silu":
        sg = (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(f32)
        sgq = dequant(quant_codes(sg, q_in2[0], q_in2[1], 0, q_in2[2]), q_in2[0], q_in2[1])
        a = (x * sgq).astype(f32)
    else:
        xd = x.astype(np.float64)
        a = (0.5 * xd * (1.0 + np.vectorize(erf)(xd / np.sqrt(2.0)))).astype(f32)
    return dequant(quant_codes(a, q_out[0], q_out[1], 0, q_out[2]), q_out[0], q_out[1])


# =================================================================================================================
# Whole-model integer forward (mirror of mobilequant_b200/engine/int_forward.py, numpy on the CPU)
# =================================================================================================================
def _sq(act_dict, recipe, name, slot):
    bits, sym, _ = recipe[name][slot]
    assert not sym
    mn, mx = (0.0, 1.0) if (slot == "input2" and slot not in act_dict[name]) else act_dict[name][slot]
    qmax = 2 ** bits - 1
    mn, mx = f32(mn), f32(mx)
    s = np.clip(f32((mx - mn) / f32(qmax)), f32(1e-5), f32(1e6))
    o = -np.rint(f32(mn / s)) + f32(0)
    return f32(s), f32(o), f32(qmax)


class IntModel:
    def __init__(self, sd, cfg, recipe, act_dict):
        self.sd = {k: (v.detach().cpu().numpy().astype(f32) if hasattr(v, "detach") else np.asarray(v, f32)) for k, v in sd.items()}
        self.cfg, self.recipe, self.act = cfg, recipe, act_dict
        self.nh, self.nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
        self.hd = cfg.get("head_dim") or cfg["hidden_size"] // self.nh
        self.rot = int(cfg.get("partial_rotary_factor", 1.0) * self.hd)
        self.H = cfg["hidden_size"]
        self.layernorm = cfg.get("norm_class", "rmsnorm") == "layernorm"

    def _w(self, name):
        bits, sym, pc = self.recipe[name]["weight"]
        return weight_quant(self.sd[name + ".weight"], bits, sym, pc)

    def _bias(self, name):
        b = self.sd.get(name + ".bias")
        return None if b is None or np.abs(b).max() == 0 else b

    def _lin_y(self, a_codes, qx, name):
        codes, sw, ow, _, _ = self._w(name)
        N = codes.shape[0]
        sw = np.broadcast_to(sw, (N,)) if sw.size == 1 else sw
        ow = np.broadcast_to(ow, (N,)) if ow.size == 1 else ow
        return qlinear_y(a_codes, qx[0], int(qx[1]), codes, sw, ow.astype(np.int64), self._bias(name))

    def norm(self, h, name):
        bits, sym, pc = self.recipe[name]["weight"]
        codes, sw, ow, _, _ = weight_quant(self.sd[name + ".weight"], bits, sym, pc)
        w_fq = dequant(codes.astype(f32), sw.reshape(1, 1), ow.reshape(1, 1)).reshape(-1)
        return qnorm_int(h, _sq(self.act, self.recipe, name, "input"), w_fq, self._bias(name), _sq(self.act, self.recipe, name, "output"),
                         self.layernorm, self.cfg.get("layer_norm_eps", 1e-5))

    def _qkv(self, h, i):
        """input norm + fused q|k|v projection codes of block i; returns (x1, qkv, qin, qout)."""
        p = f"model.layers.{i}."
        sq = lambda n, s: _sq(self.act, self.recipe, p + n, s)
        x1 = self.norm(h, p + "input_layernorm")
        qx = sq("input_layernorm", "output")
        outs = []
        for n in ("q_proj", "k_proj", "v_proj"):
            y = self._lin_y(x1, qx, p + "self_attn." + n)
            qo = sq("self_attn." + n, "output")
            outs.append(quant_codes(y, qo[0], qo[1], 0, qo[2]).astype(np.int64))
        qkv = np.concatenate(outs, axis=1)
        qin = [sq("self_attn." + n, "output")[:2] for n in ("q_proj", "k_proj", "v_proj")]
        qout = [sq("self_attn.qk_bmm", "input")[:2], sq("self_attn.qk_bmm", "input2")[:2], sq("self_attn.pv_bmm", "input2")[:2]]
        return x1, qkv, qin, qout

    def block(self, h, i, B, T, cos, sin, trace=None, cache=None):
        """cache (optional): dict layer -> [k, v] code arrays [B, nkv, T, hd], filled for the decode steps that follow."""
        p = f"model.layers.{i}."
        sq = lambda n, s: _sq(self.act, self.recipe, p + n, s)
        x1, qkv, qin, qout = self._qkv(h, i)
        q, k, v = qrope_int(qkv, B, T, self.nh, self.nkv, self.hd, self.rot, qin, qout, cos, sin)
        if cache is not None:
            cache[i] = [k, v]
        attn = qattn_int(q, k, v, self.nh, self.nkv, qout[0], qout[1], qout[2], sq("self_attn.qk_bmm", "output"),
                         sq("self_attn.pv_bmm", "input"), sq("self_attn.pv_bmm", "output")[:2])
        return self._tail(h, attn, i, x1, trace, dict(x1=x1, qkv=qkv, q=q, k=k, v=v))

    def decode_block(self, h, i, pos, cache, cos, sin):
        """One-token step of block i against the uint8 KV cache (SimBlock / SimAttention with k_cache, v_cache,
        mobilellm/model/sim_model.py:271-330; capp/src/llm.cpp:545-653): h [B, H] rows of the new tokens at position pos;
        the rotated k / v codes are appended to cache[i]."""
        p = f"model.layers.{i}."
        sq = lambda n, s: _sq(self.act, self.recipe, p + n, s)
        B = h.shape[0]
        x1, qkv, qin, qout = self._qkv(h, i)
        q, k, v = qrope_int(qkv, B, 1, self.nh, self.nkv, self.hd, self.rot, qin, qout, cos[pos:pos + 1], sin[pos:pos + 1])
        cache[i] = [np.concatenate([cache[i][0], k], axis=2), np.concatenate([cache[i][1], v], axis=2)]     # sim_model.py:229-231
        attn = qattn_decode_int(q[:, :, 0], cache[i][0], cache[i][1], self.nh, self.nkv, qout[0], qout[1], qout[2], sq("self_attn.qk_bmm", "output"),
                                sq("self_attn.pv_bmm", "input"), sq("self_attn.pv_bmm", "output")[:2])
        return self._tail(h, attn, i, x1)

    def decode(self, h, pos, cache, cos, sin):
        for i in range(self.cfg["num_hidden_layers"]):
            h = self.decode_block(h, i, pos, cache, cos, sin)
        return h

    def _tail(self, h, attn, i, x1=None, trace=None, tr0=None):
        """o_proj + residual, post-attention norm, MLP + residual (everything of the block after attention).  Block variants
        of hm:1257-1263: sequential (MLP reads the norm of the updated residual) and parallel with shared norm (phi: the MLP
        reads input_layernorm's codes x1); two-linear MLP (num_linears_per_mlp == 2): no gate."""
        p = f"model.layers.{i}."
        A, R = self.act, self.recipe
        sq = lambda n, s: _sq(A, R, p + n, s)
        qa = sq("self_attn.pv_bmm", "output")
        y = self._lin_y(attn, qa, p + "self_attn.o_proj")
        qo = sq("self_attn.o_proj", "output")
        h = (h + dequant(quant_codes(y, qo[0], qo[1], 0, qo[2]), qo[0], qo[1])).astype(f32)
        h_mid = h
        parallel = bool(self.cfg.get("parallel_residual"))
        assert parallel == bool(self.cfg.get("shared_attention_norm")), "mixed parallel / shared-norm blocks are not integer paths"
        if parallel:
            x2, qx2 = x1, sq("input_layernorm", "output")
        else:
            x2 = self.norm(h, p + "post_attention_layernorm")
            qx2 = sq("post_attention_layernorm", "output")
        q1 = sq("mlp.w1", "output")
        c1 = quant_codes(self._lin_y(x2, qx2, p + "mlp.w1"), q1[0], q1[1], 0, q1[2]).astype(np.int64)
        kind = "silu" if "input2" in R[p + "mlp.act_fn"] else "gelu"
        lut = act_lut(kind, q1, sq("mlp.act_fn", "input2") if kind == "silu" else None, sq("mlp.act_fn", "output"))
        if self.cfg.get("num_linears_per_mlp", 3) == 3:
            q3 = sq("mlp.w3", "output")
            c3 = quant_codes(self._lin_y(x2, qx2, p + "mlp.w3"), q3[0], q3[1], 0, q3[2])
            prod = (lut[c1] * dequant(c3, q3[0], q3[1])).astype(f32)
        else:
            prod = lut[c1].astype(f32)
        qw2 = sq("mlp.w2", "input")
        act = quant_codes(prod, qw2[0], qw2[1], 0, qw2[2]).astype(np.int64)
        y = self._lin_y(act, qw2, p + "mlp.w2")
        qo = sq("mlp.w2", "output")
        h = (h + dequant(quant_codes(y, qo[0], qo[1], 0, qo[2]), qo[0], qo[1])).astype(f32)
        if trace is not None:
            trace.update(tr0 or {})
            trace.update(attn=attn, h_mid=h_mid, x2=x2, act=act)
        return h

    def backbone(self, h, B, T, cos=None, sin=None, trace_layer=None, cache=None):
        if cos is None:
            cos, sin = rope_tables(T, self.rot, self.cfg.get("rope_theta", 10000.0))
        trace = None
        for i in range(self.cfg["num_hidden_layers"]):
            tr = {} if trace_layer == i else None
            h = self.block(h, i, B, T, cos, sin, tr, cache)
            if tr is not None:
                trace = tr
        return h, trace

    def embed(self, ids):
        h = self.sd["model.embed_tokens.weight"][np.asarray(ids)]
        if self.cfg.get("normalize_embed"):
            h = (h * f32(self.cfg["hidden_size"] ** 0.5)).astype(f32)
        return h.reshape(-1, self.H).astype(f32)
