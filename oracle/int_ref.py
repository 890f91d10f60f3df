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
