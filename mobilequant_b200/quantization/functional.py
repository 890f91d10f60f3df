"""autograd.Function wrappers: forward AND backward are single fused CUDA kernels from libmqb200 (no torch elementwise
chains; the reference spends ~8 launches forward / ~10 backward per Quantizer call)."""
import torch
from .. import kernels as K


class StaticFakeQuantFn(torch.autograd.Function):
    """Quantizer.forward with cached scale/offset (qmodule.py:279-295).  scale/offset: 0-d (per tensor) or one entry
    per row of the last dimension (cached per-channel weight quantizer)."""

    @staticmethod
    def forward(ctx, x, scale, offset, qmin, qmax):
        group = 0 if scale.numel() == 1 else x.shape[-1]
        s = scale.detach().reshape(-1).float().contiguous()
        o = offset.detach().reshape(-1).float().contiguous()
        xc = x.detach().float().contiguous()
        y, _ = K.fq_fwd(xc, s, o, qmin, qmax, group=group)
        ctx.save_for_backward(xc, s, o)
        ctx.meta = (qmin, qmax, group, scale.shape, offset.shape, x.dtype)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        xc, s, o = ctx.saved_tensors
        qmin, qmax, group, sshape, oshape, xdtype = ctx.meta
        need_x, need_s, need_o = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if (need_s or need_o) and group != 0:
            raise NotImplementedError("learnable per-channel static scales are not part of any MobileQuant recipe")
        gx, gs, go = K.fq_bwd(xc, g.float().contiguous(), s, o, qmin, qmax, group=group, want_gx=need_x,
                              want_gparams=need_s or need_o)
        return (gx.to(xdtype) if need_x else None, gs.reshape(sshape) if need_s else None,
                go.reshape(oshape) if need_o else None, None, None)


class LetLwcWeightQuantFn(torch.autograd.Function):
    """LET transform (algorithm.py:60-96) + dynamic / LWC Quantizer.forward (qmodule.py:262-290) in one pass over the
    weight.  Returns (w_fq, scale, offset); scale/offset are not differentiable outputs (the reference re-derives them
    every forward)."""

    @staticmethod
    def forward(ctx, w, col_fac, col_mode, row_fac, row_mode, sig_up, sig_low, bits, symmetric, per_channel, out=None):
        w2 = w.detach().float().contiguous()
        cf = None if col_fac is None else col_fac.detach().reshape(-1).float().contiguous()
        rf = None if row_fac is None else row_fac.detach().reshape(-1).float().contiguous()
        su = None if sig_up is None else sig_up.detach().reshape(-1).float().contiguous()
        sl = None if sig_low is None else sig_low.detach().reshape(-1).float().contiguous()
        out = K.wprep_fwd(w2, bits, symmetric, per_channel, cf, col_mode, rf, row_mode, su, sl, out=out)
        ctx.save_for_backward(w2, cf, rf, su, sl, out["minmax"])
        ctx.meta = (col_mode, row_mode, bits, symmetric, per_channel,
                    None if col_fac is None else col_fac.shape, None if row_fac is None else row_fac.shape,
                    None if sig_up is None else sig_up.shape, None if sig_low is None else sig_low.shape, w.dtype)
        ctx.mark_non_differentiable(out["scale"], out["offset"])
        return out["w_fq"].to(w.dtype), out["scale"], out["offset"]

    @staticmethod
    def backward(ctx, g, _gs, _go):
        w2, cf, rf, su, sl, mm = ctx.saved_tensors
        col_mode, row_mode, bits, symmetric, per_channel, cshape, rshape, ushape, lshape, wdtype = ctx.meta
        n = ctx.needs_input_grad
        need_w = n[0]
        if need_w and (col_mode or row_mode):
            raise NotImplementedError("dL/dW through a fused LET transform is never needed (weights are frozen)")
        res = K.wprep_bwd(w2, g.float().contiguous(), bits, symmetric, per_channel, cf, col_mode, rf, row_mode, su, sl,
                          need_col=n[1], need_row=n[3], need_sig=n[5] or n[6], need_wt=need_w, minmax=mm)
        g_col, g_row, g_up, g_low = res[:4]
        g_w = res[4].to(wdtype) if need_w else None
        return (g_w, g_col.reshape(cshape) if (n[1] and g_col is not None) else None, None,
                g_row.reshape(rshape) if (n[3] and g_row is not None) else None, None,
                g_up.reshape(ushape) if (n[5] and g_up is not None) else None,
                g_low.reshape(lshape) if (n[6] and g_low is not None) else None, None, None, None, None)


class AttnProbsFn(torch.autograd.Function):
    """fq2(softmax(fq1(S) / sqrt(hd) + causal mask)): the element-wise attention core between qk_bmm and pv_bmm (hm:514-534 with
    qk_bmm.output_quantizer and pv_bmm.input_quantizer, qm:453-466) as one kernel forward and one backward
    (csrc/calib_attn.cu).  Saves S and two floats per row; scale / offset gradients of both quantizers come out of the same
    backward pass.  q1 / q2 are (scale, offset, qmin, qmax) or scale None for a disabled quantizer."""

    @staticmethod
    def forward(ctx, S, mul, s1, o1, qmin1, qmax1, s2, o2, qmin2, qmax2):
        Sc = S.detach().contiguous()
        f = lambda t: None if t is None else t.detach().reshape(()).float().contiguous()
        q1 = None if s1 is None else (f(s1), f(o1), qmin1, qmax1)
        q2 = None if s2 is None else (f(s2), f(o2), qmin2, qmax2)
        Tq = Sc.shape[-2]
        P, stats = K.attn_probs_fwd(Sc, Tq, True, mul, q1, q2)
        ctx.save_for_backward(Sc, stats, *(t for q in (q1, q2) if q is not None for t in q[:2]))
        ctx.meta = (mul, Tq, None if q1 is None else (qmin1, qmax1), None if q2 is None else (qmin2, qmax2),
                    None if s1 is None else (s1.shape, o1.shape), None if s2 is None else (s2.shape, o2.shape))
        return P

    @staticmethod
    def backward(ctx, g):
        Sc, stats, *qs = ctx.saved_tensors
        mul, Tq, r1, r2, sh1, sh2 = ctx.meta
        q1 = q2 = None
        if r1 is not None:
            q1 = (qs[0], qs[1], r1[0], r1[1]); qs = qs[2:]
        if r2 is not None:
            q2 = (qs[0], qs[1], r2[0], r2[1])
        n = ctx.needs_input_grad
        want = n[2] or n[3] or n[6] or n[7]
        dS, gp = K.attn_probs_bwd(Sc, stats, g.float().contiguous(), Tq, True, mul, q1, q2, want_gparams=want)
        pick = lambda need, i, shape: gp[i].reshape(shape) if (need and shape is not None) else None
        return (dS if n[0] else None, None,
                pick(n[2], 0, sh1 and sh1[0]), pick(n[3], 1, sh1 and sh1[1]), None, None,
                pick(n[6], 2, sh2 and sh2[0]), pick(n[7], 3, sh2 and sh2[1]), None, None)


def _pack_q(params):
    """flat [s, o, qmin, qmax] * n (s None = disabled) -> kernel-side list + the tensors to save + shapes for the gradients."""
    qs, saved, shapes = [], [], []
    for i in range(0, len(params), 4):
        s, o, lo, hi = params[i:i + 4]
        if s is None:
            qs.append(None); shapes.append(None)
            continue
        sd, od = s.detach().reshape(()).float().contiguous(), o.detach().reshape(()).float().contiguous()
        qs.append((sd, od, lo, hi)); saved += [sd, od]; shapes.append((s.shape, o.shape))
    return qs, saved, shapes


def _unpack_q(saved, bounds):
    qs, k = [], 0
    for b in bounds:
        if b is None:
            qs.append(None)
        else:
            qs.append((saved[k], saved[k + 1], b[0], b[1])); k += 2
    return qs


class SiluGateFn(torch.autograd.Function):
    """fq_w(fq_o(A * fq_s(sigmoid(A))) * B) with A = fq_a(ya), B = fq_b(yb): the output quantizers of w1 / w3, QSiLU (qm:691-753),
    the gate product (hm:1059) and w2.input_quantizer as one kernel forward and one backward (csrc/calib_act.cu).  `y` holds
    ya | yb side by side ([.., 2*I], the result of ONE GEMM over the concatenated w1 / w3 weights).  params = (scale, offset,
    qmin, qmax) of fq_a, fq_b, fq_s, fq_o, fq_w flattened; scale None disables a quantizer."""

    @staticmethod
    def forward(ctx, y, *params):
        yc = y.detach().contiguous()
        qs, saved, shapes = _pack_q(params)
        out = K.silu_gate_fwd(qs, y=yc)
        ctx.save_for_backward(yc, *saved)
        ctx.meta = ([None if q is None else (q[2], q[3]) for q in qs], shapes)
        return out

    @staticmethod
    def backward(ctx, g):
        yc, *saved = ctx.saved_tensors
        bounds, shapes = ctx.meta
        qs = _unpack_q(saved, bounds)
        n = ctx.needs_input_grad
        want = any(n[1 + 4 * i] or n[2 + 4 * i] for i in range(5))
        dy, _, gp = K.silu_gate_bwd(g.float().contiguous(), qs, y=yc, want_gparams=want)
        out = [dy if n[0] else None]
        for i in range(5):
            sh = shapes[i]
            out += [gp[2 * i].reshape(sh[0]) if (sh is not None and n[1 + 4 * i]) else None,
                    gp[2 * i + 1].reshape(sh[1]) if (sh is not None and n[2 + 4 * i]) else None, None, None]
        return tuple(out)


class RmsNormL2Fn(torch.autograd.Function):
    """QRMSNorm.forward in its L2-norm form (qm:515-531, hm:187-195): input quantizer, F.normalize, alpha, weight, bias and
    output quantizer as one kernel forward, one backward plus the fixed-order fold of dL/dweight (csrc/calib_act.cu).
    params = (scale, offset, qmin, qmax) of the input and the output quantizer, flattened."""

    @staticmethod
    def forward(ctx, x, w, bias, alpha, eps, *params):
        xc = x.detach().contiguous()
        wc = w.detach().reshape(-1).float().contiguous()
        bc = None if bias is None else bias.detach().reshape(-1).float().contiguous()
        qs, saved, shapes = _pack_q(params)
        out, nrm = K.rmsnorm_l2_fwd(xc, wc, bc, alpha, eps, qs)
        ctx.save_for_backward(xc, wc, nrm, *([bc] if bc is not None else []), *saved)
        ctx.meta = (alpha, eps, bc is not None, [None if q is None else (q[2], q[3]) for q in qs], shapes, w.shape,
                    None if bias is None else bias.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        alpha, eps, has_b, bounds, shapes, wshape, bshape = ctx.meta
        xc, wc, nrm, *rest = ctx.saved_tensors
        bc = rest.pop(0) if has_b else None
        qs = _unpack_q(rest, bounds)
        n = ctx.needs_input_grad
        want = any(n[5 + 4 * i] or n[6 + 4 * i] for i in range(2))
        dx, dw, db, gp = K.rmsnorm_l2_bwd(xc, wc, bc, nrm, g.float().contiguous(), alpha, eps, qs, want_dbias=has_b and n[2],
                                          want_gparams=want)
        out = [dx if n[0] else None, dw.reshape(wshape) if n[1] else None, db.reshape(bshape) if (has_b and n[2]) else None, None, None]
        for i in range(2):
            sh = shapes[i]
            out += [gp[2 * i].reshape(sh[0]) if (sh is not None and n[5 + 4 * i]) else None,
                    gp[2 * i + 1].reshape(sh[1]) if (sh is not None and n[6 + 4 * i]) else None, None, None]
        return tuple(out)


class QkvRopeFn(torch.autograd.Function):
    """Everything between the (concatenated) q/k/v projection GEMM and the attention matmuls (hm:470-512): the three QLinear
    output quantizers, the head split / transpose, RoPE, and the QMatMul input quantizers of q, k and v, as one kernel forward
    and one backward (csrc/calib_act.cu).  y: [B, T, (nh + 2 nkv) hd]; returns q [B, nh, T, hd], k, v [B, nkv, T, hd].
    params = (scale, offset, qmin, qmax) x 6 flattened (q_proj.out, k_proj.out, v_proj.out, qk.input, qk.input2, pv.input2)."""

    @staticmethod
    def forward(ctx, y, cos, sin, nh, nkv, hd, rot, *params):
        yc = y.detach().contiguous()
        B, T = yc.shape[0], yc.shape[1]
        qs, saved, shapes = _pack_q(params)
        q, k, v = K.qkv_rope_fwd(yc, B, T, nh, nkv, hd, rot, cos, sin, qs)
        ctx.save_for_backward(yc, cos, sin, *saved)
        ctx.meta = (B, T, nh, nkv, hd, rot, [None if t is None else (t[2], t[3]) for t in qs], shapes)
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        yc, cos, sin, *saved = ctx.saved_tensors
        B, T, nh, nkv, hd, rot, bounds, shapes = ctx.meta
        qs = _unpack_q(saved, bounds)
        n = ctx.needs_input_grad
        want = any(n[7 + 4 * i] or n[8 + 4 * i] for i in range(6))
        dy, gp = K.qkv_rope_bwd(yc, B, T, nh, nkv, hd, rot, cos, sin, dq.float().contiguous(), dk.float().contiguous(),
                                dv.float().contiguous(), qs, want_gparams=want)
        out = [dy if n[0] else None, None, None, None, None, None, None]
        for i in range(6):
            sh = shapes[i]
            out += [gp[2 * i].reshape(sh[0]) if (sh is not None and n[7 + 4 * i]) else None,
                    gp[2 * i + 1].reshape(sh[1]) if (sh is not None and n[8 + 4 * i]) else None, None, None]
        return tuple(out)


class GroupedWeightFn(torch.autograd.Function):
    """The row-wise concatenation of several fake-quantised weights WITHOUT the copy: each part was written by its weight pass
    straight into its row slice of `buf` (wprep_fwd(out=...)), so the forward just hands out the buffer and the backward splits
    the gradient back into per-weight row blocks (contiguous views)."""

    @staticmethod
    def forward(ctx, buf, *parts):
        ctx.rows = [p.shape[0] for p in parts]
        return buf.view_as(buf)

    @staticmethod
    def backward(ctx, g):
        return (None,) + tuple(g.split(ctx.rows, dim=0))
