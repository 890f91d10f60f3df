"""K1 / K8 / K2 CUDA kernels (through the C ABI) against the CPU oracle and the reference-generated goldens.
Bar: forward values and integer codes bit-exact; gradients to fp32 summation round-off (tolerance stated)."""
import os
import torch
import pytest
from oracle import fakequant_ref as fr

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_fq_static_golden(cuda, golden_dir):
    from mobilequant_b200 import kernels as K
    for c in _load(golden_dir, "quantizer.pt"):
        if c["lwc"]:
            continue
        x = c["x"].to(cuda); s = c["scale"].to(cuda).reshape(()); o = c["offset"].to(cuda).reshape(())
        y, codes = K.fq_fwd(x, s, o, c["qmin"], c["qmax"], want_codes=True)
        assert torch.equal(y.cpu(), c["y"])
        ref_codes = fr.quant_codes(c["x"], c["scale"], c["offset"], c["qmin"], c["qmax"]).to(torch.int32)
        assert torch.equal(codes.cpu(), ref_codes)
        gx, gs, go = K.fq_bwd(x, c["gy"].to(cuda), s, o, c["qmin"], c["qmax"])
        assert torch.equal(gx.cpu(), c["gx"])
        # sums over up to ~10k elements: fp32 order-of-summation tolerance
        assert torch.allclose(gs.cpu(), c["g_scale"], rtol=2e-5, atol=1e-4)
        assert torch.allclose(go.cpu(), c["g_offset"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096, 1 << 20, (1 << 22) + 5])
def test_fq_sizes_and_unaligned(cuda, n):
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g) * 3
    s, o, qmin, qmax = fr.scale_offset_from_minmax(-2.5, 4.0, 8, False)
    y, codes = K.fq_fwd(x.to(cuda), s.to(cuda), o.to(cuda), qmin, qmax, want_codes=True)
    assert torch.equal(y.cpu(), fr.fake_quant(x, s, o, qmin, qmax))
    assert torch.equal(codes.cpu().float(), fr.quant_codes(x, s, o, qmin, qmax))
    # a misaligned view takes the scalar path
    if n > 8:
        xv = x.to(cuda)[1:]
        y2, _ = K.fq_fwd(xv, s.to(cuda), o.to(cuda), qmin, qmax)
        assert torch.equal(y2.cpu(), fr.fake_quant(x[1:], s, o, qmin, qmax))


def test_fq_fractional_offset_lrl(cuda):
    """After LRL steps scale/offset are arbitrary floats: same arithmetic, non-integer codes."""
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(7)
    x = torch.randn(5, 300, 64, generator=g)
    s = torch.tensor(0.0123457); o = torch.tensor(127.00037)
    xr = x.clone().requires_grad_(True); sr = s.clone().requires_grad_(True); orr = o.clone().requires_grad_(True)
    yr = fr.fake_quant(xr, sr, orr, 0, 255)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    y, _ = K.fq_fwd(x.to(cuda), s.to(cuda), o.to(cuda), 0, 255)
    assert torch.equal(y.cpu(), yr.detach())
    gx, gs, go = K.fq_bwd(x.to(cuda), gy.to(cuda), s.to(cuda), o.to(cuda), 0, 255)
    assert torch.equal(gx.cpu(), xr.grad)
    assert torch.allclose(gs.cpu(), sr.grad, rtol=1e-4, atol=1e-3)
    assert torch.allclose(go.cpu(), orr.grad, rtol=1e-4, atol=1e-6)


def test_fq_per_channel_groups(cuda):
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    w = torch.randn(64, 96, generator=g) * 0.05
    mn, mx = fr.tensor_minmax(w, True)
    s, o, qmin, qmax = fr.scale_offset_from_minmax(mn, mx, 4, True)
    y, codes = K.fq_fwd(w.to(cuda), s.reshape(-1).to(cuda), o.reshape(-1).to(cuda), qmin, qmax, group=96,
                        want_codes=True)
    assert torch.equal(y.cpu(), fr.fake_quant(w, s, o, qmin, qmax))


@pytest.mark.parametrize("n", [1, 5, 4096, (1 << 21) + 3])
def test_minmax(cuda, n):
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g)
    out = K.minmax(x.to(cuda))
    assert out.cpu().tolist() == [x.min().item(), x.max().item()]
    x2 = torch.randn(n, generator=g) * 2
    K.minmax(x2.to(cuda), out, accumulate=True)      # running update, generate_act_range.py:69
    assert out.cpu().tolist() == [min(x.min().item(), x2.min().item()), max(x.max().item(), x2.max().item())]


def test_minmax_2d(cuda):
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(11)
    x = torch.randn(777, 130, generator=g)
    mn, mx = K.minmax_2d(x.to(cuda), per_row=True)
    assert torch.equal(mn.cpu(), x.amin(1)) and torch.equal(mx.cpu(), x.amax(1))
    mn, mx = K.minmax_2d(x.to(cuda), per_row=False)
    assert torch.equal(mn.cpu(), x.amin(0)) and torch.equal(mx.cpu(), x.amax(0))


def test_wprep_golden_lwc(cuda, golden_dir):
    """LWC weight quantizer forward (bit-exact vs the reference Quantizer) and bound-factor gradients."""
    from mobilequant_b200 import kernels as K
    for c in _load(golden_dir, "quantizer.pt"):
        if not c["lwc"]:
            continue
        w = c["x"].to(cuda)
        up = c["up"].clone().requires_grad_(True); low = c["low"].clone().requires_grad_(True)
        su, sl = torch.sigmoid(up), torch.sigmoid(low)
        out = K.wprep_fwd(w, c["bits"], c["sym"], c["per_channel"], sig_up=su.detach().reshape(-1).to(cuda),
                          sig_low=sl.detach().reshape(-1).to(cuda), want_codes=c["bits"] <= 8)
        assert torch.equal(out["w_fq"].cpu(), c["y"]), (c["bits"], c["sym"], c["per_channel"])
        assert torch.equal(out["scale"].cpu(), c["scale"].reshape(-1))
        assert torch.equal(out["offset"].cpu(), c["offset"].reshape(-1) + 0.0)
        if c["bits"] <= 8:
            ref_codes = fr.quant_codes(c["x"], c["scale"], c["offset"], c["qmin"], c["qmax"]).to(torch.int64)
            assert torch.equal(out["codes"].cpu().to(torch.int64), ref_codes)
            assert torch.equal(out["colsum"].cpu().to(torch.int64), ref_codes.sum(1))
        _, _, g_su, g_sl = K.wprep_bwd(w, c["gy"].to(cuda), c["bits"], c["sym"], c["per_channel"],
                                       sig_up=su.detach().reshape(-1).to(cuda), sig_low=sl.detach().reshape(-1).to(cuda))
        su.backward(g_su.cpu().reshape(su.shape)); sl.backward(g_sl.cpu().reshape(sl.shape))
        assert torch.allclose(up.grad, c["g_up"], rtol=1e-4, atol=1e-5), (up.grad - c["g_up"]).abs().max()
        assert torch.allclose(low.grad, c["g_low"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("bits,sym,per_ch,col_mode,row_mode", [
    (8, False, False, 2, 0),   # q/k/v/w1 under ln smoothing (alg:68)
    (8, False, False, 2, 1),   # w3: * fc1 scale, / fc2 scale (alg:68,77)
    (8, False, True, 2, 0),    # w2 per-channel (alg:87)
    (4, True, True, 2, 2),     # k_proj under q-k smoothing (alg:95), W4 symmetric
    (16, False, False, 1, 0),  # norm weight / scale (alg:60)
])
def test_wprep_let_fwd_bwd(cuda, bits, sym, per_ch, col_mode, row_mode):
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(bits * 100 + col_mode * 10 + row_mode)
    rows, cols = (1, 256) if bits == 16 else (96, 160)
    w = torch.randn(rows, cols, generator=g) * 0.02
    cf = (1 + 0.2 * torch.randn(cols, generator=g)).requires_grad_(True)
    rf = (1 + 0.2 * torch.randn(rows, generator=g)).requires_grad_(True)
    groups = rows if per_ch else 1
    up = (4 + 0.5 * torch.randn(groups, 1, generator=g)).requires_grad_(True)
    low = (4 + 0.5 * torch.randn(groups, 1, generator=g)).requires_grad_(True)
    su, sl = torch.sigmoid(up), torch.sigmoid(low)
    wt = fr.let_weight(w, cf, col_mode, rf if row_mode else None, row_mode)
    y = fr.dynamic_fake_quant(wt, bits, sym, per_ch, su if per_ch else su.reshape(-1)[0], sl if per_ch else sl.reshape(-1)[0])
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy, retain_graph=True)
    d = lambda t: t.detach().reshape(-1).to(cuda).contiguous()
    out = K.wprep_fwd(w.to(cuda), bits, sym, per_ch, d(cf), col_mode, d(rf) if row_mode else None, row_mode, d(su), d(sl),
                      want_wt=True)
    assert torch.equal(out["wt"].cpu(), wt.detach())
    assert torch.equal(out["w_fq"].cpu(), y.detach())
    g_col, g_row, g_up, g_low = K.wprep_bwd(w.to(cuda), gy.to(cuda), bits, sym, per_ch, d(cf), col_mode,
                                            d(rf) if row_mode else None, row_mode, d(su), d(sl))
    tol = dict(rtol=2e-4, atol=2e-5)
    assert torch.allclose(g_col.cpu(), cf.grad, **tol), (g_col.cpu() - cf.grad).abs().max()
    if row_mode:
        assert torch.allclose(g_row.cpu(), rf.grad, **tol), (g_row.cpu() - rf.grad).abs().max()
    gsu = torch.autograd.grad(su, up, g_up.cpu().reshape(su.shape))[0]
    gsl = torch.autograd.grad(sl, low, g_low.cpu().reshape(sl.shape))[0]
    assert torch.allclose(gsu, up.grad, **tol), (gsu - up.grad).abs().max()
    assert torch.allclose(gsl, low.grad, **tol)


def test_wprep_pack4(cuda):
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(5)
    w = torch.randn(32, 64, generator=g) * 0.02
    out = K.wprep_fwd(w.to(cuda), 4, True, True, want_codes=True, pack4=True)
    codes, _, _ = fr.weight_codes(w, 4, True, True)
    packed = out["codes"].cpu().view(torch.uint8).to(torch.int64)
    lo = packed & 0xF; hi = packed >> 4
    sext = lambda v: torch.where(v >= 8, v - 16, v)
    un = torch.stack([sext(lo), sext(hi)], dim=-1).reshape(32, 64)
    assert torch.equal(un, codes)


@pytest.mark.parametrize("mode,scale", [(0, 1.0), (2, 1.0), (1, 0.0123), (1, 9.1553e-4), (1, 1.0 / 65535), (1, 0.99999994), (1, 3.05e-5)])
def test_div_rn_exact(cuda, mode, scale):
    """The branch-free exact requantisation of the integer-engine kernels == IEEE division + rint (qm:286) on 2^28
    pseudo-random operand pairs per case, including all-ones significands (the Markstein exception)."""
    from mobilequant_b200 import kernels as K
    assert K.selftest_div(1 << 28, seed=7 + mode, mode=mode, fixed_scale=scale) == 0


def test_flat_adamw_matches_torch_adamw(cuda):
    """mq_adamw_step (global grad norm + skip-on-non-finite + AdamW over a flat buffer, three learning-rate groups) against
    torch.optim.AdamW on the CPU (what the reference constructs, alg:513,716-722) and the scaler's norm (optim.py:5-25)."""
    from mobilequant_b200.utils.optim import FlatAdamW, ampscaler_get_grad_norm
    g = torch.Generator().manual_seed(11)
    shapes = [(), (), (37,), (64, 1), (1,), (130,), ()]
    lrs = [1e-3, 1e-2, 1e-6]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    dev_p = [torch.nn.Parameter(p.detach().clone().to(cuda)) for p in ref_p]
    grouping = [[0, 2, 5], [3, 4], [1, 6]]
    ref = torch.optim.AdamW([{"params": [ref_p[i] for i in grp], "lr": lr} for grp, lr in zip(grouping, lrs)], weight_decay=0.01)
    opt = FlatAdamW([{"params": [dev_p[i] for i in grp], "lr": lr} for grp, lr in zip(grouping, lrs)], weight_decay=0.01, device=cuda)
    for p, q in zip(ref_p, dev_p):
        assert torch.equal(p.detach(), q.detach().cpu())                 # re-pointing kept the values
    for it in range(6):
        grads = [torch.randn(s, generator=g) * (10.0 ** (it - 3)) for s in shapes]
        opt.zero_grad()
        for p, q, gr in zip(ref_p, dev_p, grads):
            p.grad = gr.clone()
            q.grad.add_(gr.to(cuda))                                      # autograd accumulates in place into the flat views
        for i, lr in enumerate(lrs):
            ref.param_groups[i]["lr"] = lr * (1 + it)
            opt.set_lr(i, lr * (1 + it))
        opt.sync_lr()
        norm_ref = ampscaler_get_grad_norm(ref_p)
        ref.step()
        norm = opt.step()
        assert norm.item() == pytest.approx(norm_ref.item(), rel=1e-6)
        for p, q in zip(ref_p, dev_p):
            assert torch.allclose(q.detach().cpu(), p.detach(), rtol=1e-5, atol=1e-7), it
    # a non-finite gradient skips the step entirely (GradScaler.step semantics)
    before = [q.detach().clone() for q in dev_p]
    step_before = opt.state[0].item()
    opt.zero_grad()
    dev_p[2].grad[3] = float("inf")
    norm = opt.step()
    assert not torch.isfinite(norm)
    assert all(torch.equal(a, b.detach()) for a, b in zip(before, dev_p)) and opt.state[0].item() == step_before and opt.skipped_steps == 1


def _probs_chain(S, hd, q1, q2):
    """The op-by-op attention core of HFAttention.forward (hm:514-534) with StaticFakeQuantFn as the quantizers."""
    import math
    from mobilequant_b200.quantization.functional import StaticFakeQuantFn
    from mobilequant_b200.model.hf_model import causal_mask_4d
    T = S.shape[-1]
    a = S if q1 is None else StaticFakeQuantFn.apply(S, *q1)
    a = a / math.sqrt(hd) + causal_mask_4d(S.shape[0], T, torch.float32, S.device)
    p = torch.nn.functional.softmax(a, dim=-1, dtype=torch.float32)
    return p if q2 is None else StaticFakeQuantFn.apply(p, *q2)


@pytest.mark.parametrize("T,hd,bits1,pmin", [(64, 64, 16, 0.0), (256, 80, 8, 0.0), (1024, 64, 16, 0.0), (2048, 256, 16, 0.0),
                                             (512, 64, 16, 0.05), (36, 64, 16, 0.0), (1024, 64, None, 0.0)])
def test_attn_probs_fused_matches_chain(cuda, T, hd, bits1, pmin):
    """csrc/calib_attn.cu (forward + backward incl. the LRL scale / offset gradients) against the element-wise chain it
    replaces.  pmin > 0 puts fq2's offset outside its code range (masked columns then contribute to the gradients)."""
    import math
    from mobilequant_b200.quantization.functional import AttnProbsFn
    from mobilequant_b200.quantization.qmodule import compute_scale_offset_from_min_max
    torch.manual_seed(T + hd)
    B, nh = (1, 2) if T >= 1024 else (2, 3)
    S0 = (torch.randn(B, nh, T, T, device=cuda) * 6.0)
    W = torch.randn(B, nh, T, T, device=cuda)

    def qparams(mn, mx, bits):
        if bits is None:
            return None
        s, o, _, _, lo, hi = compute_scale_offset_from_min_max(mn, mx, bits, False)
        return [torch.nn.Parameter(s.to(cuda)), torch.nn.Parameter(o.to(cuda)), lo, hi]

    outs = []
    for fused in (False, True):
        S = S0.clone().requires_grad_(True)
        q1, q2 = qparams(-20.0, 18.0, bits1), qparams(pmin, 1.0, 16)
        if fused:
            mul = (torch.ones(()) / torch.tensor(math.sqrt(hd))).item()
            f1 = q1 if q1 is not None else [None, None, 0.0, 0.0]
            out = AttnProbsFn.apply(S, mul, *f1, *q2)
        else:
            out = _probs_chain(S, hd, q1, q2)
        (out * W).sum().backward()
        outs.append((out.detach(), S.grad, q1, q2))
    (o0, g0, a1, a2), (o1, g1, b1, b2) = outs
    lsb = a2[0].item()
    d = (o0 - o1).abs()
    assert d.max().item() <= 1.01 * lsb                       # at most one code apart (exp / summation order), and rarely
    assert (d > 0).float().mean().item() < 1e-3
    gs = g0.abs().max().item()
    assert (g0 - g1).abs().max().item() < 2e-3 * gs
    assert (g0 - g1).abs().mean().item() < 1e-5 * gs
    assert (g1[..., 0, 1:] == 0).all()                         # masked columns get an exact zero
    # fq2's scale gradient is a cancelling sum of g * (rounding residual): every probability that lands on the other side of a
    # rounding boundary (exp / summation order; counted above) moves it by ~|g|, so the bound grows with sqrt(#flips)
    nflip = int((d > 0).sum().item())
    for ref, got in ((a1, b1), (a2, b2)):
        if ref is None:
            continue
        for i in (0, 1):
            r, g = ref[i].grad.item(), got[i].grad.item()
            tol = 0.02 * abs(r) + 1e-3 * W.numel() ** 0.5 * (lsb if ref is a2 else ref[0].item())
            if ref is a2 and i == 0:
                tol += 4.0 * (nflip + 1) ** 0.5
            assert abs(r - g) <= tol, (i, r, g, nflip)


@pytest.mark.parametrize("on", [(True,) * 5, (True, True, False, True, True), (False, False, True, False, False), (False,) * 5])
def test_silu_gate_fused_matches_chain(cuda, on):
    """csrc/calib_act.cu silu_gate (forward, backward, the five quantizers' LRL gradients) against the module chain it
    replaces: w2.input_quantizer(QSiLU(fq_a(ya)) * fq_b(yb)), ya | yb side by side in one GEMM result."""
    from mobilequant_b200.quantization.functional import SiluGateFn, StaticFakeQuantFn
    from mobilequant_b200.quantization.qmodule import compute_scale_offset_from_min_max
    torch.manual_seed(5)
    I = 68
    y0 = torch.randn(3, 257, 2 * I, device=cuda) * torch.tensor([3.0] * I + [2.0] * I, device=cuda)
    W = torch.randn(3, 257, I, device=cuda)

    def qparams(mn, mx, bits, enabled):
        if not enabled:
            return [None, None, 0.0, 0.0]
        s, o, _, _, lo, hi = compute_scale_offset_from_min_max(mn, mx, bits, False)
        return [torch.nn.Parameter(s.to(cuda)), torch.nn.Parameter(o.to(cuda)), lo, hi]

    res = []
    for fused in (False, True):
        y = y0.clone().requires_grad_(True)
        qs = [qparams(-8.0, 9.0, 8, on[0]), qparams(-5.0, 6.0, 8, on[1]), qparams(0.0, 1.0, 8, on[2]), qparams(-0.3, 6.0, 8, on[3]),
              qparams(-9.0, 11.0, 8, on[4])]
        if fused:
            out = SiluGateFn.apply(y, *[v for q in qs for v in q])
        else:
            fq = lambda x, q: x if q[0] is None else StaticFakeQuantFn.apply(x.contiguous(), *q)
            a, b = fq(y[..., :I], qs[0]), fq(y[..., I:], qs[1])
            out = fq(fq(a * fq(torch.sigmoid(a), qs[2]), qs[3]) * b, qs[4])
        (out * W).sum().backward()
        res.append((out.detach(), y.grad, qs))
    (o0, g0, q0), (o1, g1, q1) = res
    assert torch.equal(o0, o1)
    assert torch.allclose(g0, g1, rtol=1e-5, atol=1e-6)
    for r, g in zip(q0, q1):
        if r[0] is None:
            continue
        for i in (0, 1):
            ref, got = r[i].grad.item(), g[i].grad.item()
            # the chain evaluates d/dscale as g*t5 - (g*s*m)*(u/s) (rounding noise |g|*|u|*2^-24 per element), the fused kernel
            # as g*(t5 - m*u)
            assert abs(ref - got) <= 1e-3 * abs(ref) + 2e-2, (i, ref, got)


@pytest.mark.parametrize("H,bias,on", [(2048, False, (True, True)), (2560, True, (True, True)), (64, False, (False, True)),
                                       (1024, True, (False, False))])
def test_rmsnorm_l2_fused_matches_chain(cuda, H, bias, on):
    """csrc/calib_act.cu rmsnorm_l2 (forward, dx, dweight, dbias, LRL gradients) against QRMSNorm's op-by-op forward."""
    from mobilequant_b200.quantization import qmodule as Q
    torch.manual_seed(H)
    x0 = torch.randn(2, 37, H, device=cuda) * 2.0
    Wt = torch.randn_like(x0)
    res = []
    for fused in (False, True):
        os.environ["MQB200_FUSED_NORM"] = "1" if fused else "0"
        try:
            m = Q.QRMSNorm(dict(dim=H, eps=1e-6, device=cuda, dtype=torch.float32, l2norm_as_rmsnorm=True),
                           Q.QuantConfig(bitwidth=16), None, Q.QuantConfig(bitwidth=8))
            torch.manual_seed(1); m.weight.data = torch.randn(H, device=cuda)
            if bias:
                m.bias = torch.nn.Parameter(torch.randn(H, device=cuda) * 0.1)
            m.set_scale_offset({"input": [-7.0, 8.0], "output": [-60.0, 70.0]}, "parameter")
            if not on[0]:
                m.input_quantizer = None
            if not on[1]:
                m.output_quantizer.enable = False
            x = x0.clone().requires_grad_(True)
            out = m(x)
            (out * Wt).sum().backward()
            res.append((out.detach(), x.grad, m.weight.grad, None if not bias else m.bias.grad,
                        [0.0 if q.grad is None else q.grad.item() for q in m.parameters() if q.dim() == 0]))
        finally:
            os.environ.pop("MQB200_FUSED_NORM", None)
    (o0, gx0, gw0, gb0, gq0), (o1, gx1, gw1, gb1, gq1) = res
    lsb = 130.0 / 255
    d = (o0 - o1).abs()
    assert d.max().item() <= (1.01 * lsb if on[1] else 1e-4)      # ||x|| summation order: at most one output code apart, rarely
    assert (d > 1e-5).float().mean().item() < 2e-3
    gs = gx0.abs().max().item()
    assert (gx0 - gx1).abs().max().item() < 0.05 * gs and (gx0 - gx1).abs().mean().item() < 2e-4 * gs
    assert torch.allclose(gw0, gw1, rtol=2e-3, atol=2e-3 * gw0.abs().max().item())
    if bias:
        assert torch.allclose(gb0, gb1, rtol=2e-3, atol=2e-3 * gb0.abs().max().item())
    assert len(gq0) == len(gq1)
    for r, g in zip(gq0, gq1):
        assert abs(r - g) <= 0.03 * abs(r) + 0.05 * max(abs(v) for v in gq0), (gq0, gq1)


@pytest.mark.parametrize("nh,nkv,hd,rot,on", [(8, 2, 64, 64, True), (4, 4, 64, 16, True), (6, 3, 80, 32, True), (8, 2, 64, 64, False)])
def test_qkv_rope_fused_matches_chain(cuda, nh, nkv, hd, rot, on):
    """csrc/calib_act.cu qkv_rope (forward, dy, the six quantizers' LRL gradients) against the module graph it replaces:
    output quantizers -> head split / transpose -> RoPE (hm:338-367, partial rotary hm:489-501) -> matmul input quantizers."""
    from mobilequant_b200.quantization.functional import QkvRopeFn, StaticFakeQuantFn
    from mobilequant_b200.quantization.qmodule import compute_scale_offset_from_min_max
    from mobilequant_b200.model.hf_model import rope_cos_sin, apply_rotary_pos_emb
    torch.manual_seed(nh * hd + rot)
    B, T = 2, 37
    y0 = torch.randn(B, T, (nh + 2 * nkv) * hd, device=cuda) * 2.0
    Wq, Wk, Wv = torch.randn(B, nh, T, hd, device=cuda), torch.randn(B, nkv, T, hd, device=cuda), torch.randn(B, nkv, T, hd, device=cuda)
    pos = torch.arange(T, device=cuda).unsqueeze(0) + torch.tensor([[0], [3]], device=cuda)
    cos, sin = rope_cos_sin(pos, rot, 10000.0, cuda, torch.float32)

    def qparams(mn, mx, bits):
        if not on:
            return [None, None, 0.0, 0.0]
        s, o, _, _, lo, hi = compute_scale_offset_from_min_max(mn, mx, bits, False)
        return [torch.nn.Parameter(s.to(cuda)), torch.nn.Parameter(o.to(cuda)), lo, hi]

    res = []
    for fused in (False, True):
        y = y0.clone().requires_grad_(True)
        qs = [qparams(-6.0, 7.0, 8), qparams(-5.0, 5.5, 8), qparams(-7.0, 6.0, 8), qparams(-8.0, 8.5, 8), qparams(-7.5, 7.0, 8),
              qparams(-6.5, 6.0, 8)]
        if fused:
            q, k, v = QkvRopeFn.apply(y, cos.contiguous(), sin.contiguous(), nh, nkv, hd, rot, *[t for p in qs for t in p])
        else:
            fq = lambda x, p: x if p[0] is None else StaticFakeQuantFn.apply(x.contiguous(), *p)
            yq, yk, yv = y.split([nh * hd, nkv * hd, nkv * hd], dim=-1)
            q = fq(yq, qs[0]).view(B, T, nh, hd).transpose(1, 2)
            k = fq(yk, qs[1]).view(B, T, nkv, hd).transpose(1, 2)
            v = fq(yv, qs[2]).view(B, T, nkv, hd).transpose(1, 2)
            if rot == hd:
                q, k = apply_rotary_pos_emb(q, k, cos, sin)
            else:
                qr, kr = apply_rotary_pos_emb(q[..., :rot], k[..., :rot], cos, sin)
                q = torch.cat((qr, q[..., rot:]), dim=-1)
                k = torch.cat((kr, k[..., rot:]), dim=-1)
            q, k, v = fq(q, qs[3]), fq(k, qs[4]), fq(v, qs[5])
        ((q * Wq).sum() + (k * Wk).sum() + (v * Wv).sum()).backward()
        res.append((q.detach(), k.detach(), v.detach(), y.grad, qs))
    (q0, k0, v0, g0, p0), (q1, k1, v1, g1, p1) = res
    assert torch.equal(q0, q1) and torch.equal(k0, k1) and torch.equal(v0, v1)
    assert torch.allclose(g0, g1, rtol=1e-5, atol=1e-6)
    for r, g in zip(p0, p1):
        if r[0] is None:
            continue
        for i in (0, 1):
            ref, got = r[i].grad.item(), g[i].grad.item()
            assert abs(ref - got) <= 1e-3 * abs(ref) + 2e-2, (i, ref, got)
