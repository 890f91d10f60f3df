"""Tensor-level wrappers over the C ABI (include/mqb200.h).  These are the only callers of libmqb200 in the package."""
import ctypes
import math
from ctypes import c_int, c_int32, c_int64, c_float, c_void_p
import torch
from . import _lib
from ._lib import mq_qcfg, ptr, check, stream_ptr, MQError

_P = c_void_p
_protos_done = False


def _protos():
    global _protos_done
    if _protos_done:
        return _lib.load()
    lib = _lib.load()
    lib.mq_fq_fwd.argtypes = [_P, _P, _P, _P, c_int64, _P, _P, c_int64, c_float, c_float, _P]
    lib.mq_fq_bwd.argtypes = [_P, _P, _P, _P, c_int64, _P, _P, c_int64, c_float, c_float, _P, _P, _P]
    lib.mq_attn_probs_supported.argtypes = [c_int]
    lib.mq_attn_probs_fwd.argtypes = [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_float, _P, _P, c_float, c_float, _P, _P, c_float,
                                      c_float, _P]
    lib.mq_attn_probs_bwd.argtypes = [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_float, _P, _P, c_float, c_float, _P, _P,
                                      c_float, c_float, _P, _P]
    lib.mq_silu_gate_fwd.argtypes = [_P, _P, _P, c_int64, _P, c_int64, c_int, _P, _P, _P, _P, _P]
    lib.mq_silu_gate_bwd.argtypes = [_P, _P, _P, c_int64, _P, _P, _P, c_int64, c_int64, c_int, _P, _P, _P, _P, _P, _P]
    lib.mq_rmsnorm_l2_supported.argtypes = [c_int]
    lib.mq_rmsnorm_l2_fwd.argtypes = [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_float, _P, _P, _P, _P, _P]
    lib.mq_rmsnorm_l2_bwd.argtypes = [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_float, _P, _P, _P, _P, _P, _P]
    lib.mq_qkv_rope_fwd.argtypes = [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P]
    lib.mq_qkv_rope_bwd.argtypes = [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]
    lib.mq_minmax.argtypes = [_P, _P, c_int64, _P, c_int, _P]
    lib.mq_minmax_2d.argtypes = [_P, _P, c_int64, c_int64, c_int, _P, _P, c_int, _P]
    lib.mq_wprep_fwd.argtypes = [_P, _P, c_int64, c_int64, _P, c_int, _P, c_int, _P, _P, c_int, mq_qcfg,
                                 _P, _P, c_int, _P, _P, _P, _P, _P, _P]
    lib.mq_wprep_bwd.argtypes = [_P, _P, _P, c_int64, c_int64, _P, c_int, _P, c_int, _P, _P, c_int, mq_qcfg,
                                 _P, _P, _P, _P, _P, _P, _P, _P]
    lib.mq_qgemm.argtypes = [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, c_float,
                             c_int, _P, c_int64, _P, _P, c_float, c_float, c_float, _P, c_int, _P]
    lib.mq_qgemm_w4a8.argtypes = [_P, _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, c_float,
                                  c_int, _P, c_int64, _P, _P, c_float, c_float, c_float, _P, c_int, _P]
    lib.mq_qnorm.argtypes = [_P, _P, c_int64, c_int, c_int, c_float, c_float, c_float, _P, _P, c_float, c_float, c_float,
                             c_float, c_float, _P, _P, _P]
    lib.mq_qnorm_resid.argtypes = [_P, _P, c_int, c_int, c_int, c_float, c_float, c_float, _P, _P, c_float, c_float, c_float, c_float, c_float,
                                   _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, c_float, c_int, _P]
    lib.mq_qrope.argtypes = [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]
    lib.mq_qattn.argtypes = [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]
    lib.mq_qattn_shard.argtypes = [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]
    lib.mq_qgemv.argtypes = [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P]
    lib.mq_qgemv_epilogue.argtypes = [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, c_float, _P, c_int64, _P, _P,
                                      c_float, c_float, c_float, _P, c_int, _P, _P]
    lib.mq_qgemv_fused.argtypes = [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, c_float,
                                   _P, c_int64, _P, _P, c_float, c_float, c_float, _P, c_int, _P, _P]
    lib.mq_unpack4.argtypes = [_P, _P, c_int64, c_int, _P, _P]
    lib.mq_fgemv.argtypes = [_P, _P, _P, _P, c_int, c_int, c_int, _P]
    lib.mq_qattn_decode.argtypes = [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P,
                                    _P, _P, _P, _P, _P]
    lib.mq_selftest_div.argtypes = [_P, c_int64, ctypes.c_uint64, c_int, c_float, _P, _P]
    lib.mq_adamw_step.argtypes = [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P, c_float, c_float, c_float, c_float, _P, _P]
    _protos_done = True
    return lib


def _h(t):
    return _lib.ctx(t.device.index)


# ---- launch accounting (bench.py: "gpu_launches", per-kernel-class device time) ----------------------------------
_launches = 0
_timing = None


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count():
    return _launches


def enable_event_timing(on):
    global _timing
    _timing = {} if on else None


def collect_event_timing():
    """{name: {"ms": total device ms, "n": calls}}; call after torch.cuda.synchronize()."""
    out = {}
    for name, evs in (_timing or {}).items():
        out[name] = {"ms": sum(a.elapsed_time(b) for a, b in evs), "n": len(evs)}
    return out


def _launch(name, fn, *args):
    global _launches
    _launches += 1
    if _timing is None:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    _timing.setdefault(name, []).append((e0, e1))
    return rc


F32 = torch.float32


# ---- K1 ---------------------------------------------------------------------------------------------------------
def fq_fwd(x, scale, offset, qmin, qmax, group=0, want_y=True, want_codes=False):
    lib = _protos()
    x = x.contiguous()
    y = torch.empty_like(x) if want_y else None
    codes = torch.empty(x.shape, dtype=torch.int32, device=x.device) if want_codes else None
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("fq_fwd", lib.mq_fq_fwd, h, ptr(x, F32), ptr(y), ptr(codes), x.numel(), ptr(scale, F32), ptr(offset, F32),
                            int(group), float(qmin), float(qmax), stream_ptr()), h)
    return y, codes


def fq_bwd(x, g, scale, offset, qmin, qmax, group=0, want_gx=True, want_gparams=True):
    lib = _protos()
    x = x.contiguous(); g = g.contiguous()
    gx = torch.empty_like(x) if want_gx else None
    gs = torch.empty((), dtype=F32, device=x.device) if want_gparams else None
    go = torch.empty((), dtype=F32, device=x.device) if want_gparams else None
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("fq_bwd", lib.mq_fq_bwd, h, ptr(x, F32), ptr(g, F32), ptr(gx), x.numel(), ptr(scale, F32), ptr(offset, F32),
                            int(group), float(qmin), float(qmax), ptr(gs), ptr(go), stream_ptr()), h)
    return gx, gs, go


def attn_probs_supported(T):
    return bool(_protos().mq_attn_probs_supported(int(T)))


def attn_probs_fwd(S, Tq, causal, mul, q1, q2):
    """P = fq2(softmax(fq1(S) * mul + causal mask)) over the last dim of S [..., Tq, T] (contiguous fp32).  q1 / q2: None or
    (scale, offset, qmin, qmax) with 0-d CUDA scale / offset.  Returns (P, stats)."""
    lib = _protos()
    T = S.shape[-1]
    rows = S.numel() // T
    P = torch.empty_like(S)
    stats = torch.empty((rows, 2), dtype=F32, device=S.device)
    h = _h(S)
    s1, o1, lo1, hi1 = q1 if q1 is not None else (None, None, 0.0, 0.0)
    s2, o2, lo2, hi2 = q2 if q2 is not None else (None, None, 0.0, 0.0)
    with torch.cuda.device(S.device):
        check(_launch("attn_probs_fwd", lib.mq_attn_probs_fwd, h, ptr(S, F32), ptr(P), ptr(stats), rows, T, int(Tq), int(causal), float(mul),
                      ptr(s1), ptr(o1), float(lo1), float(hi1), ptr(s2), ptr(o2), float(lo2), float(hi2), stream_ptr()), h)
    return P, stats


def attn_probs_bwd(S, stats, g, Tq, causal, mul, q1, q2, want_gparams=True):
    lib = _protos()
    T = S.shape[-1]
    rows = S.numel() // T
    dS = torch.empty_like(S)
    gp = torch.empty(4, dtype=F32, device=S.device) if want_gparams else None
    h = _h(S)
    s1, o1, lo1, hi1 = q1 if q1 is not None else (None, None, 0.0, 0.0)
    s2, o2, lo2, hi2 = q2 if q2 is not None else (None, None, 0.0, 0.0)
    with torch.cuda.device(S.device):
        check(_launch("attn_probs_bwd", lib.mq_attn_probs_bwd, h, ptr(S, F32), ptr(stats, F32), ptr(g, F32), ptr(dS), rows, T, int(Tq),
                      int(causal), float(mul), ptr(s1), ptr(o1), float(lo1), float(hi1), ptr(s2), ptr(o2), float(lo2), float(hi2),
                      ptr(gp), stream_ptr()), h)
    return dS, gp


def _qarrays(qs):
    """[(scale, offset, qmin, qmax) | None, ...] -> host arrays of device pointers / bounds for the fused calibration kernels."""
    n = len(qs)
    sc = (c_void_p * n)(*[ptr(q[0], F32) if q is not None else c_void_p(0) for q in qs])
    of = (c_void_p * n)(*[ptr(q[1], F32) if q is not None else c_void_p(0) for q in qs])
    lo = (c_float * n)(*[float(q[2]) if q is not None else 0.0 for q in qs])
    hi = (c_float * n)(*[float(q[3]) if q is not None else 0.0 for q in qs])
    return sc, of, lo, hi


def _gate_operands(y, a, b):
    """(pointer a, pointer b, row stride, rows, cols) of the gate operands: the two halves of y [rows, 2*cols], or a / b [.., cols]."""
    if y is not None:
        cols = y.shape[-1] // 2
        rows = y.numel() // (2 * cols)
        base = ptr(y, F32)
        return base, c_void_p(y.data_ptr() + 4 * cols), 2 * cols, rows, cols
    cols = a.shape[-1]
    return ptr(a, F32), ptr(b, F32), cols, a.numel() // cols, cols


def silu_gate_fwd(qs, y=None, a=None, b=None):
    """fq_w(fq_o(A * fq_s(sigmoid(A))) * B), A = fq_a(ya), B = fq_b(yb); qs = [fq_a, fq_b, fq_s, fq_o, fq_w], each None or
    (scale, offset, qmin, qmax).  Operands: y [.., 2*I] (ya | yb side by side, one GEMM result) or separate a, b [.., I]."""
    lib = _protos()
    pa, pb, ld, rows, cols = _gate_operands(y, a, b)
    src = y if y is not None else a
    out = torch.empty(src.shape[:-1] + (cols,), dtype=F32, device=src.device)
    sc, of, lo, hi = _qarrays(qs)
    h = _h(src)
    with torch.cuda.device(src.device):
        check(_launch("silu_gate_fwd", lib.mq_silu_gate_fwd, h, pa, pb, ld, ptr(out), rows, cols, sc, of, lo, hi, stream_ptr()), h)
    return out


def silu_gate_bwd(g, qs, y=None, a=None, b=None, want_gparams=True):
    """Returns (dy, None, gparams) for the side-by-side layout, (da, db, gparams) otherwise."""
    lib = _protos()
    pa, pb, ld, rows, cols = _gate_operands(y, a, b)
    src = y if y is not None else a
    if y is not None:
        dy = torch.empty_like(y)
        pda, pdb, ldd = ptr(dy), c_void_p(dy.data_ptr() + 4 * cols), 2 * cols
        res = (dy, None)
    else:
        da, db = torch.empty_like(a), torch.empty_like(a)
        pda, pdb, ldd = ptr(da), ptr(db), cols
        res = (da, db)
    gp = torch.empty(10, dtype=F32, device=src.device) if want_gparams else None
    sc, of, lo, hi = _qarrays(qs)
    h = _h(src)
    with torch.cuda.device(src.device):
        check(_launch("silu_gate_bwd", lib.mq_silu_gate_bwd, h, pa, pb, ld, ptr(g, F32), pda, pdb, ldd, rows, cols, sc, of, lo, hi, ptr(gp),
                      stream_ptr()), h)
    return res + (gp,)


def rmsnorm_l2_supported(H):
    return bool(_protos().mq_rmsnorm_l2_supported(int(H)))


def rmsnorm_l2_fwd(x, w, bias, alpha, eps, qs):
    """fq_out(w * (alpha * fq_in(x) / max(||fq_in(x)||, eps)) + bias) over the last dim; qs = [fq_in, fq_out].  Returns (out, nrm)."""
    lib = _protos()
    H = x.shape[-1]
    rows = x.numel() // H
    out = torch.empty_like(x)
    nrm = torch.empty(rows, dtype=F32, device=x.device)
    sc, of, lo, hi = _qarrays(qs)
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("rmsnorm_l2_fwd", lib.mq_rmsnorm_l2_fwd, h, ptr(x, F32), ptr(w, F32), ptr(bias), ptr(out), ptr(nrm), rows, H, float(alpha),
                      float(eps), sc, of, lo, hi, stream_ptr()), h)
    return out, nrm


def rmsnorm_l2_bwd(x, w, bias, nrm, g, alpha, eps, qs, want_dbias=False, want_gparams=True):
    lib = _protos()
    H = x.shape[-1]
    rows = x.numel() // H
    dx = torch.empty_like(x)
    dw = torch.empty(H, dtype=F32, device=x.device)
    dbias = torch.empty(H, dtype=F32, device=x.device) if want_dbias else None
    gp = torch.empty(4, dtype=F32, device=x.device) if want_gparams else None
    sc, of, lo, hi = _qarrays(qs)
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("rmsnorm_l2_bwd", lib.mq_rmsnorm_l2_bwd, h, ptr(x, F32), ptr(w, F32), ptr(bias), ptr(nrm, F32), ptr(g, F32), ptr(dx), ptr(dw),
                      ptr(dbias), rows, H, float(alpha), float(eps), sc, of, lo, hi, ptr(gp), stream_ptr()), h)
    return dx, dw, dbias, gp


def qkv_rope_fwd(y, B, T, nh, nkv, hd, rot, cos, sin, qs):
    """y [B, T, (nh + 2 nkv) hd] -> q [B, nh, T, hd], k, v [B, nkv, T, hd] (quantise, rotate, quantise; see include/mqb200.h).
    qs = [q_proj.out, k_proj.out, v_proj.out, qk.input, qk.input2, pv.input2]."""
    lib = _protos()
    dev = y.device
    q = torch.empty((B, nh, T, hd), dtype=F32, device=dev)
    k = torch.empty((B, nkv, T, hd), dtype=F32, device=dev)
    v = torch.empty((B, nkv, T, hd), dtype=F32, device=dev)
    sc, of, lo, hi = _qarrays(qs)
    h = _h(y)
    with torch.cuda.device(dev):
        check(_launch("qkv_rope_fwd", lib.mq_qkv_rope_fwd, h, ptr(y, F32), B * T, T, nh, nkv, hd, rot, ptr(cos), ptr(sin),
                      int(cos is not None and cos.shape[0] > 1), ptr(q), ptr(k), ptr(v), sc, of, lo, hi, stream_ptr()), h)
    return q, k, v


def qkv_rope_bwd(y, B, T, nh, nkv, hd, rot, cos, sin, dq, dk, dv, qs, want_gparams=True):
    lib = _protos()
    dy = torch.empty_like(y)
    gp = torch.empty(12, dtype=F32, device=y.device) if want_gparams else None
    sc, of, lo, hi = _qarrays(qs)
    h = _h(y)
    with torch.cuda.device(y.device):
        check(_launch("qkv_rope_bwd", lib.mq_qkv_rope_bwd, h, ptr(y, F32), B * T, T, nh, nkv, hd, rot, ptr(cos), ptr(sin),
                      int(cos is not None and cos.shape[0] > 1), ptr(dq, F32), ptr(dk, F32), ptr(dv, F32), ptr(dy), sc, of, lo, hi, ptr(gp),
                      stream_ptr()), h)
    return dy, gp


# ---- K8 ---------------------------------------------------------------------------------------------------------
def minmax(x, out=None, accumulate=False):
    """out: float32[2] CUDA tensor = [min, max] (running when accumulate)."""
    lib = _protos()
    x = x.contiguous()
    if out is None:
        out = torch.empty(2, dtype=F32, device=x.device); accumulate = False
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("minmax", lib.mq_minmax, h, ptr(x, F32), x.numel(), ptr(out, F32), int(accumulate), stream_ptr()), h)
    return out


def minmax_2d(x2d, per_row, out_min=None, out_max=None, accumulate=False):
    lib = _protos()
    x2d = x2d.contiguous()
    rows, cols = x2d.shape
    n = rows if per_row else cols
    if out_min is None:
        out_min = torch.empty(n, dtype=F32, device=x2d.device); out_max = torch.empty_like(out_min); accumulate = False
    h = _h(x2d)
    with torch.cuda.device(x2d.device):
        check(_launch("minmax_2d", lib.mq_minmax_2d, h, ptr(x2d, F32), rows, cols, int(per_row), ptr(out_min, F32), ptr(out_max, F32),
                               int(accumulate), stream_ptr()), h)
    return out_min, out_max


# ---- K2 ---------------------------------------------------------------------------------------------------------
MODE_NONE, MODE_DIV, MODE_MUL = 0, 1, 2


def wprep_fwd(w, bits, symmetric, per_channel, col_fac=None, col_mode=0, row_fac=None, row_mode=0, sig_up=None,
              sig_low=None, want_fq=True, want_codes=False, pack4=False, want_wt=False, out=None):
    """Returns dict(w_fq, codes, scale, offset, colsum, wt).  w is [rows, cols] (a norm weight is [1, H]).  `out`: a contiguous
    fp32 tensor of w's shape that receives w_fq (e.g. a row slice of a buffer several weights share)."""
    lib = _protos()
    w = w.contiguous()
    rows, cols = w.shape
    dev = w.device
    groups = rows if per_channel else 1
    if out is not None and (out.shape != w.shape or out.dtype != F32 or out.device != w.device or not out.is_contiguous()):
        raise MQError("wprep_fwd: `out` must be a contiguous fp32 tensor of the weight's shape on its device")
    w_fq = out if (out is not None and want_fq) else (torch.empty_like(w) if want_fq else None)
    out = dict(w_fq=w_fq, codes=None, colsum=None,
               scale=torch.empty(groups, dtype=F32, device=dev), offset=torch.empty(groups, dtype=F32, device=dev),
               wt=torch.empty_like(w) if want_wt else None, minmax=torch.empty(2 * groups, dtype=F32, device=dev))
    if want_codes:
        ncode = rows * cols // 2 if pack4 else rows * cols
        out["codes"] = torch.empty(ncode, dtype=torch.int8 if symmetric else torch.uint8, device=dev)
        out["codes"] = out["codes"].view(rows, -1)
        out["colsum"] = torch.empty(rows, dtype=torch.int32, device=dev)
    h = _h(w)
    cfg = mq_qcfg(int(bits), int(bool(symmetric)))
    with torch.cuda.device(dev):
        check(_launch("wprep_fwd", lib.mq_wprep_fwd, h, ptr(w, F32), rows, cols, ptr(col_fac), int(col_mode), ptr(row_fac), int(row_mode),
                               ptr(sig_up), ptr(sig_low), int(bool(per_channel)), cfg, ptr(out["w_fq"]),
                               ptr(out["codes"]), int(bool(pack4)), ptr(out["scale"]), ptr(out["offset"]),
                               ptr(out["colsum"]), ptr(out["wt"]), ptr(out["minmax"]), stream_ptr()), h)
    return out


def wprep_bwd(w, g, bits, symmetric, per_channel, col_fac=None, col_mode=0, row_fac=None, row_mode=0, sig_up=None,
              sig_low=None, need_col=True, need_row=True, need_sig=True, need_wt=False, minmax=None):
    """`minmax`: the forward's out["minmax"] (group min / max of the transformed weight); without it they are recomputed."""
    lib = _protos()
    w = w.contiguous(); g = g.contiguous()
    rows, cols = w.shape
    dev = w.device
    groups = rows if per_channel else 1
    g_col = torch.empty(cols, dtype=F32, device=dev) if (need_col and col_mode) else None
    g_row = torch.empty(rows, dtype=F32, device=dev) if (need_row and row_mode) else None
    g_up = torch.empty(groups, dtype=F32, device=dev) if (need_sig and sig_up is not None) else None
    g_low = torch.empty(groups, dtype=F32, device=dev) if (need_sig and sig_low is not None) else None
    scratch = torch.empty_like(w) if g_col is not None else None
    g_wt = torch.empty_like(w) if need_wt else None
    h = _h(w)
    cfg = mq_qcfg(int(bits), int(bool(symmetric)))
    with torch.cuda.device(dev):
        check(_launch("wprep_bwd", lib.mq_wprep_bwd, h, ptr(w, F32), ptr(g, F32), rows, cols, ptr(col_fac), int(col_mode), ptr(row_fac),
                               int(row_mode), ptr(sig_up), ptr(sig_low), int(bool(per_channel)), cfg, ptr(g_col),
                               ptr(g_row), ptr(g_up), ptr(g_low), ptr(g_wt), ptr(scratch), ptr(minmax), stream_ptr()), h)
    if need_wt:
        return g_col, g_row, g_up, g_low, g_wt
    return g_col, g_row, g_up, g_low


# ---- K3/K7 ------------------------------------------------------------------------------------------------------
EPI_QUANT, EPI_ACTMUL, EPI_RESID, EPI_F32, EPI_I32 = 0, 1, 2, 3, 4


def qgemm(a, b, rowsum, sxw, ow, c0, mode, bias=None, so=None, oo=None, qmax=255.0, out_bits=8, out=None, ldo=None,
          rowsum_out=None, lut=None, s2=1.0, o2=0.0, qmax2=255.0, resid=None, qgroup=None, packed4=False):
    """a: [M,K] uint8/int8 codes, b: [N,K] uint8/int8 codes.  so/oo: one entry per `qgroup` columns (default: one
    entry for the whole N when they have a single element).  See include/mqb200.h:mq_qgemm.
    packed4: b is [N, K/2] uint8 holding two unsigned 4-bit codes per byte (mq_qgemm_w4a8: expanded inside the kernel)."""
    lib = _protos()
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == (K // 2 if packed4 else K)
    dev = a.device
    if out is None and mode != EPI_RESID:
        if mode == EPI_QUANT:
            out = torch.empty(M, N, dtype=torch.uint8 if out_bits == 8 else torch.int16, device=dev)
        elif mode == EPI_ACTMUL:
            out = torch.empty(M, N // 2, dtype=torch.uint8, device=dev)
        elif mode == EPI_F32:
            out = torch.empty(M, N, dtype=F32, device=dev)
        else:
            out = torch.empty(M, N, dtype=torch.int32, device=dev)
    if ldo is None:
        ldo = resid.shape[-1] if mode == EPI_RESID else out.shape[-1]
    if qgroup is None:
        qgroup = (N + 31) // 32 * 32 if (so is None or so.numel() == 1) else 0
        if mode == EPI_ACTMUL and so is not None and so.numel() == N // 128:
            qgroup = 128
    if so is not None and mode in (EPI_QUANT, EPI_ACTMUL, EPI_RESID):
        assert qgroup > 0 and so.numel() == (N + qgroup - 1) // qgroup == oo.numel(), "so/oo need one entry per qgroup columns"
    h = _h(a)
    tail = (ptr(rowsum, torch.int32), ptr(sxw, F32), ptr(ow, torch.int32), ptr(c0, torch.int32), ptr(bias), int(mode), ptr(so), ptr(oo),
            float(qmax), int(out_bits), ptr(out), int(ldo), ptr(rowsum_out), ptr(lut), float(s2), float(o2), float(qmax2), ptr(resid),
            int(qgroup), stream_ptr())
    with torch.cuda.device(dev):
        if packed4:
            check(_launch("qgemm", lib.mq_qgemm_w4a8, h, ptr(a), int(a.dtype == torch.int8), ptr(b, torch.uint8), M, N, K, *tail), h)
        else:
            check(_launch("qgemm", lib.mq_qgemm, h, ptr(a), int(a.dtype == torch.int8), ptr(b), int(b.dtype == torch.int8), M, N, K, *tail), h)
    return resid if mode == EPI_RESID else out


# ---- K4 / K5 / K6 (integer engine) ---------------------------------------------------------------------------------
def _host_floats(vals):
    arr = (c_float * len(vals))(*[float(v) for v in vals])
    return arr


def qnorm(x, qin, w_fq, bias, qout, layernorm=False, eps=1e-5, codes=None, rowsum=None):
    """x fp32 [rows, H]; qin/qout = (scale, offset, qmax) python floats.  Returns (codes u8, rowsum int32)."""
    lib = _protos()
    rows, H = x.shape
    if codes is None:
        codes = torch.empty(rows, H, dtype=torch.uint8, device=x.device)
    if rowsum is None:
        rowsum = torch.empty(rows, dtype=torch.int32, device=x.device)
    h = _h(x)
    import math
    with torch.cuda.device(x.device):
        check(_launch("qnorm", lib.mq_qnorm, h, ptr(x, F32), rows, H, int(layernorm), float(qin[0]), float(qin[1]), float(qin[2]), ptr(w_fq, F32),
                           ptr(bias), float(math.sqrt(H)), float(eps), float(qout[0]), float(qout[1]), float(qout[2]),
                           ptr(codes), ptr(rowsum), stream_ptr()), h)
    return codes, rowsum


def qnorm_resid(x, qin, w_fq, bias, qout, layernorm, eps, codes, rowsum, acc, g, g_rowsum):
    """Decode step: x[rows, H] += dequant(Q_out(acc ...)) of the preceding skinny GEMM `g` (engine weight dict: sxw / ow / c0 /
    bias / so / oo / qmax / qgroup), then qnorm of the updated rows -- one launch."""
    lib = _protos()
    rows, H = x.shape
    assert g["N"] == H
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("qnorm_resid", lib.mq_qnorm_resid, h, ptr(x, F32), rows, H, int(layernorm), float(qin[0]), float(qin[1]), float(qin[2]),
                           ptr(w_fq, F32), ptr(bias), float(math.sqrt(H)), float(eps), float(qout[0]), float(qout[1]), float(qout[2]),
                           ptr(codes), ptr(rowsum), ptr(acc, torch.int32), int(acc.stride(0)), ptr(g_rowsum, torch.int32), ptr(g["sxw"], F32),
                           ptr(g["ow"], torch.int32), ptr(g["c0"], torch.int32), ptr(g["bias"]), ptr(g["so"]), ptr(g["oo"]),
                           float(g["qmax"]), int(g["qgroup"]), stream_ptr()), h)
    return codes, rowsum


def qrope(qkv, B, T, nh, nkv, hd, rot, qin, qout, cos, sin, bufs=None):
    """qkv u8 [B*T, ldq]; qin/qout: 3 x (scale, offset).  Returns dict(q, k, vt, rsq, rsk)."""
    lib = _protos()
    dev = qkv.device
    if bufs is None:
        bufs = dict(q=torch.empty(B, nh, T, hd, dtype=torch.uint8, device=dev), k=torch.empty(B, nkv, T, hd, dtype=torch.uint8, device=dev),
                    vt=torch.empty(B, nkv, hd, T, dtype=torch.uint8, device=dev),
                    rsq=torch.empty(B, nh, T, dtype=torch.int32, device=dev), rsk=torch.empty(B, nkv, T, dtype=torch.int32, device=dev))
    h = _h(qkv)
    pin = _host_floats([v for so in qin for v in so]); pout = _host_floats([v for so in qout for v in so])
    with torch.cuda.device(dev):
        check(_launch("qrope", lib.mq_qrope, h, ptr(qkv), qkv.shape[-1], B, T, nh, nkv, hd, rot, ctypes.cast(pin, _P), ctypes.cast(pout, _P),
                           ptr(cos, F32), ptr(sin, F32), ptr(bufs["q"]), ptr(bufs["k"]), ptr(bufs["vt"]), ptr(bufs["rsq"]),
                           ptr(bufs["rsk"]), stream_ptr()), h)
    return bufs


def qattn(bufs, B, T, nh, nkv, hd, qparams, lut, out=None, rowsum_out=None):
    """qparams = [o_q, o_k, o_v, s_q*s_k, s_s, o_s, qmax_s, s_p, qmax_p, s_p*s_v, s_out, o_out] (python floats)."""
    lib = _protos()
    dev = bufs["q"].device
    if out is None:
        out = torch.empty(B * T, nh * hd, dtype=torch.uint8, device=dev)
    h = _h(out)
    pq = _host_floats(qparams)
    with torch.cuda.device(dev):
        check(_launch("qattn", lib.mq_qattn, h, ptr(bufs["q"]), ptr(bufs["k"]), ptr(bufs["vt"]), ptr(bufs["rsq"]), ptr(bufs["rsk"]), B, T, nh, nkv, hd,
                           ctypes.cast(pq, _P), ptr(lut), ptr(out), ptr(rowsum_out), stream_ptr()), h)
    return out


def qattn_shard(q, rsq, k, vt, rsk, B, Tq, T, q_start, nh, nkv, hd, qparams, lut, out=None, rowsum_out=None):
    """Attention of a query shard (absolute positions q_start .. q_start + Tq - 1) against all T keys; include/mqb200.h:mq_qattn_shard."""
    lib = _protos()
    dev = q.device
    if out is None:
        out = torch.empty(B * Tq, nh * hd, dtype=torch.uint8, device=dev)
    h = _h(out)
    pq = _host_floats(qparams)
    with torch.cuda.device(dev):
        check(_launch("qattn", lib.mq_qattn_shard, h, ptr(q), ptr(k), ptr(vt), ptr(rsq, torch.int32), ptr(rsk, torch.int32), B, Tq, T, q_start,
                      nh, nkv, hd, ctypes.cast(pq, _P), ptr(lut), ptr(out), ptr(rowsum_out), stream_ptr()), h)
    return out


# ---- decode step (skinny GEMM + epilogue, RoPE/append/attention against the int8 KV cache) ---------------------------
def qgemv(x, w, acc, ksplit=0):
    """acc[B, ldacc] (int32, zero on entry) += x[B, K] codes @ w[N, K]^T codes, B <= 128."""
    lib = _protos()
    B, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and acc.dtype == torch.int32 and acc.shape[0] >= B and acc.stride(0) >= N
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("qgemv", lib.mq_qgemv, h, ptr(x), int(x.dtype == torch.int8), ptr(w), int(w.dtype == torch.int8), B, N, K,
                           ptr(acc, torch.int32), int(acc.stride(0)), int(ksplit), stream_ptr()), h)
    return acc


def unpack4(packed, out):
    """packed uint8 [N, K/2] (two 4-bit codes per byte) -> out int8/uint8 [N, K]; out.dtype decides sign extension."""
    lib = _protos()
    n = packed.numel() * 2
    assert out.numel() == n
    h = _h(packed)
    with torch.cuda.device(packed.device):
        check(_launch("unpack4", lib.mq_unpack4, h, ptr(packed), n, int(out.dtype == torch.int8), ptr(out), stream_ptr()), h)
    return out


def fgemv(x, w, out=None):
    """fp32 out[B, V] = x[B, K] @ w[V, K]^T for B <= 16 (decode-step lm_head)."""
    lib = _protos()
    B, Kd = x.shape
    V = w.shape[0]
    if out is None:
        out = torch.empty(B, V, dtype=F32, device=x.device)
    h = _h(x)
    with torch.cuda.device(x.device):
        check(_launch("fgemv", lib.mq_fgemv, h, ptr(x, F32), ptr(w, F32), ptr(out, F32), B, V, Kd, stream_ptr()), h)
    return out


def qgemv_epilogue(acc, B, N, rowsum, sxw, ow, c0, mode, bias=None, so=None, oo=None, qmax=255.0, out=None, ldo=None, rowsum_out=None,
                   lut=None, s2=1.0, o2=0.0, qmax2=255.0, resid=None, qgroup=None, zero_out=None):
    """mq_qgemm's epilogue `mode` (EPI_QUANT 8-bit / EPI_ACTMUL / EPI_RESID) on the accumulator of qgemv; zeroes acc."""
    lib = _protos()
    dev = acc.device
    if out is None and mode != EPI_RESID:
        out = torch.empty(B, N // 2 if mode == EPI_ACTMUL else N, dtype=torch.uint8, device=dev)
    if ldo is None:
        ldo = resid.stride(0) if mode == EPI_RESID else out.stride(0)
    if qgroup is None:
        qgroup = 128 if (mode == EPI_ACTMUL and so.numel() == N // 128) else (N + 31) // 32 * 32
    assert so.numel() == (N + qgroup - 1) // qgroup == oo.numel(), "so/oo need one entry per qgroup columns"
    h = _h(acc)
    with torch.cuda.device(dev):
        check(_launch("qgemv_epi", lib.mq_qgemv_epilogue, h, ptr(acc, torch.int32), int(acc.stride(0)), B, N, ptr(rowsum, torch.int32),
                           ptr(sxw, F32), ptr(ow, torch.int32), ptr(c0, torch.int32), ptr(bias), int(mode), ptr(so), ptr(oo), float(qmax),
                           ptr(out), int(ldo), ptr(rowsum_out), ptr(lut), float(s2), float(o2), float(qmax2), ptr(resid), int(qgroup),
                           ptr(zero_out), stream_ptr()), h)
    return resid if mode == EPI_RESID else out


def qgemv_fused(x, w, acc, rowsum, sxw, ow, c0, mode, bias=None, so=None, oo=None, qmax=255.0, out=None, ldo=None, rowsum_out=None,
                lut=None, s2=1.0, o2=0.0, qmax2=255.0, resid=None, qgroup=None, zero_out=None, ksplit=0):
    """qgemv + qgemv_epilogue in one launch (the last CTA of each 128-column group runs the epilogue)."""
    lib = _protos()
    B, Kd = x.shape
    N = w.shape[0]
    dev = x.device
    if out is None and mode != EPI_RESID:
        out = torch.empty(B, N // 2 if mode == EPI_ACTMUL else N, dtype=torch.uint8, device=dev)
    if ldo is None:
        ldo = resid.stride(0) if mode == EPI_RESID else out.stride(0)
    if qgroup is None:
        qgroup = 128 if (mode == EPI_ACTMUL and so.numel() == N // 128) else (N + 31) // 32 * 32
    assert so.numel() == (N + qgroup - 1) // qgroup == oo.numel(), "so/oo need one entry per qgroup columns"
    h = _h(x)
    with torch.cuda.device(dev):
        check(_launch("qgemv_fused", lib.mq_qgemv_fused, h, ptr(x), int(x.dtype == torch.int8), ptr(w), int(w.dtype == torch.int8), B, N, Kd,
                           ptr(acc, torch.int32), int(acc.stride(0)), int(ksplit), ptr(rowsum, torch.int32), ptr(sxw, F32), ptr(ow, torch.int32),
                           ptr(c0, torch.int32), ptr(bias), int(mode), ptr(so), ptr(oo), float(qmax), ptr(out), int(ldo), ptr(rowsum_out),
                           ptr(lut), float(s2), float(o2), float(qmax2), ptr(resid), int(qgroup), ptr(zero_out), stream_ptr()), h)
    return resid if mode == EPI_RESID else out


def qattn_decode(qkv, B, nh, nkv, hd, rot, pos, qin, qout, cos, sin, k_cache, v_cache, rsk_cache, qparams, lut, out=None,
                 rowsum_out=None, pos_dev=None, pos_bound=None):
    """qkv u8 [B, ldq]; caches u8 [B, nkv, Tmax, hd] / int32 [B, nkv, Tmax] (row `pos` is written)."""
    lib = _protos()
    dev = qkv.device
    if out is None:
        out = torch.empty(B, nh * hd, dtype=torch.uint8, device=dev)
    Tmax = k_cache.shape[2]
    h = _h(qkv)
    pin = _host_floats([v for so in qin for v in so]); pout = _host_floats([v for so in qout for v in so])
    pq = _host_floats(qparams)
    with torch.cuda.device(dev):
        check(_launch("qattn_decode", lib.mq_qattn_decode, h, ptr(qkv), int(qkv.stride(0)), B, nh, nkv, hd, rot, Tmax, int(pos),
                           ptr(pos_dev), int(pos if pos_bound is None else pos_bound), ctypes.cast(pin, _P), ctypes.cast(pout, _P),
                           ptr(cos, F32), ptr(sin, F32), ptr(k_cache), ptr(v_cache), ptr(rsk_cache, torch.int32), ctypes.cast(pq, _P),
                           ptr(lut), ptr(out), ptr(rowsum_out), stream_ptr()), h)
    return out


def adamw_step(params, grads, exp_avg, exp_avg_sq, seg_end, lr, state, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """Global grad norm + skip-on-non-finite + AdamW over flat fp32 buffers (include/mqb200.h:mq_adamw_step).
    seg_end: python ints (ascending, last == params.numel()); lr: float32 CUDA tensor [len(seg_end)]; state: float32[8]."""
    lib = _protos()
    n = params.numel()
    seg = (c_int64 * len(seg_end))(*[int(v) for v in seg_end])
    h = _h(params)
    with torch.cuda.device(params.device):
        check(_launch("adamw_step", lib.mq_adamw_step, h, ptr(params, F32), ptr(grads, F32), ptr(exp_avg, F32), ptr(exp_avg_sq, F32), n,
                      len(seg_end), ctypes.cast(seg, _P), ptr(lr, F32), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                      ptr(state, F32), stream_ptr()), h)
    return state


def selftest_div(n, seed=1, mode=0, fixed_scale=1.0, device=None):
    """Number of disagreements between the branch-free exact requantisation and IEEE division + rint (must be 0)."""
    lib = _protos()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.zeros(1, dtype=torch.int64, device=dev)
    h = _lib.ctx(dev.index)
    with torch.cuda.device(dev):
        check(_launch("selftest_div", lib.mq_selftest_div, h, int(n), int(seed), int(mode), float(fixed_scale), ptr(out), stream_ptr()), h)
    return int(out.item())
