"""TEST INFRASTRUCTURE ONLY -- CPU oracle: a restatement of the reference's fake-quant arithmetic.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
Nothing under mobilequant_b200/ imports it and the product has no CPU path.

The reference is PyTorch, so the restatement uses torch CPU fp32 ops (the reference's own arithmetic library):
IEEE fp32 division, torch.round (half-to-even), clamp.  Each function cites the reference lines it restates
("qm" = mobilellm/quantization/qmodule.py, "alg" = mobilellm/quantization/algorithm.py,
"hm" = mobilellm/model/hf_model.py).  Parity pinning: oracle/make_golden.py runs the UNMODIFIED reference in the
build container and checks these functions against it (bit-exact) before writing tests/golden/*.pt; the reference
itself ships no golden vectors.
"""
import math
import torch
import torch.nn.functional as F

CLIPMIN, CLIPMAX = 1e-5, 1e6   # qm:11-12


# ---------------------------------------------------------------------------------------------------------------
# Quantizer arithmetic
# ---------------------------------------------------------------------------------------------------------------
def qrange(bits, symmetric):
    """qm:45-54"""
    if symmetric:
        return -2 ** (bits - 1), 2 ** (bits - 1) - 1
    return 0, 2 ** bits - 1


def scale_offset_from_minmax(mn, mx, bits, symmetric):
    """qm:40-61.  mn/mx: tensors (any shape) or python floats.  Returns (scale, offset, qmin, qmax)."""
    mn = mn if torch.is_tensor(mn) else torch.tensor(mn)
    mx = mx if torch.is_tensor(mx) else torch.tensor(mx)
    qmin, qmax = qrange(bits, symmetric)
    if symmetric:
        alpha = torch.maximum(mn.abs(), mx.abs())
        beta = 0
    else:
        alpha = mx - mn
        beta = mn
    scale = (alpha / qmax).clamp(min=CLIPMIN, max=CLIPMAX)
    offset = -(beta / scale).round()
    return scale, offset, qmin, qmax


def minmax_from_scale_offset(scale, offset, bits, symmetric):
    """qm:66-76"""
    _, qmax = qrange(bits, symmetric)
    scale = scale.clamp(min=CLIPMIN, max=CLIPMAX)
    alpha = scale * qmax
    beta = -offset * scale
    mx = alpha + beta
    mn = beta if not symmetric else -mx
    return mn, mx


def round_ste(x):
    """qm:17-21"""
    return (x.round() - x).detach() + x


def quant_codes(x, scale, offset, qmin, qmax):
    """qm:286-287 -- the pre-dequant integer code (float tensor holding integers when offset is integral)."""
    return (round_ste(x / scale) + offset).clamp(qmin, qmax)


def fake_quant(x, scale, offset, qmin, qmax):
    """qm:286-290"""
    return (quant_codes(x, scale, offset, qmin, qmax) - offset) * scale


def tensor_minmax(x, per_channel):
    """qm:26-34 (group_size == -1)"""
    if per_channel:
        return torch.amin(x, dim=-1, keepdim=True), torch.amax(x, dim=-1, keepdim=True)
    y = x.contiguous().view(-1)
    return torch.amin(y, dim=-1), torch.amax(y, dim=-1)


def dynamic_fake_quant(x, bits, symmetric, per_channel, sig_up=None, sig_low=None, return_params=False):
    """Quantizer.forward on the dynamic / LWC branch, qm:262-290.  sig_up/sig_low = sigmoid(bound factors)."""
    mn, mx = tensor_minmax(x, per_channel)
    if sig_up is not None:
        mx = sig_up * mx      # qm:271
        mn = sig_low * mn     # qm:272
    scale, offset, qmin, qmax = scale_offset_from_minmax(mn, mx, bits, symmetric)
    y = fake_quant(x, scale, offset, qmin, qmax)
    if return_params:
        return y, scale, offset, qmin, qmax
    return y


# ---------------------------------------------------------------------------------------------------------------
# LET (learnable equivalent transformation) on weights, alg:47-96
# ---------------------------------------------------------------------------------------------------------------
MODE_NONE, MODE_DIV, MODE_MUL = 0, 1, 2


def let_weight(w, col_fac=None, col_mode=0, row_fac=None, row_mode=0):
    """W' = (W {*,/} col_fac[k]) {/,*} row_fac[n].
    fc.weight * scales.view(1,-1)  alg:68,87 ; ln.weight / scales  alg:60 ;
    temp_weight / scales.view(-1,1)  alg:77,93 ; temp_weight * scales.view(-1,1)  alg:95."""
    t = w
    if col_mode == MODE_MUL:
        t = t * col_fac.view(1, -1)
    elif col_mode == MODE_DIV:
        t = t / col_fac.view(1, -1)
    if row_mode == MODE_DIV:
        t = t / row_fac.view(-1, 1)
    elif row_mode == MODE_MUL:
        t = t * row_fac.view(-1, 1)
    return t


def weight_codes(w, bits, symmetric, per_channel):
    """Integer codes + params of a static weight quantizer on its first forward (qm:262-277): int64 codes."""
    mn, mx = tensor_minmax(w, per_channel)
    scale, offset, qmin, qmax = scale_offset_from_minmax(mn, mx, bits, symmetric)
    codes = quant_codes(w, scale, offset, qmin, qmax)
    return codes.to(torch.int64), scale.reshape(-1), offset.reshape(-1)
