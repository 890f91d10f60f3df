"""Decode step against the int8 KV cache (mq_qgemv / mq_qgemv_epilogue / mq_qattn_decode, IntEngine.prefill / decode_step /
generate).  Oracle: the full-sequence integer forward (oracle/int_ref.py) -- with static quantizers row `pos` of a full
forward IS the decode step -- so every check is bit-exact."""
import numpy as np
import pytest
import torch
from oracle import int_ref as ir
from oracle import model_ref as mr
from helpers import load_golden, MODEL_GOLDENS, product_model

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.mark.parametrize("B,N,K,signed", [(1, 256, 128, False), (8, 2560, 2048, False), (17, 200, 352, True), (64, 512, 5632, False),
                                          (100, 384, 256, True), (128, 128, 2048, False)])
def test_qgemv_exact_integer(cuda, B, N, K, signed):
    from mobilequant_b200 import kernels as Kn
    rng = np.random.default_rng(B + N + K)
    x = rng.integers(0, 256, size=(B, K)).astype(np.uint8)
    w = rng.integers(-128 if signed else 0, 128 if signed else 256, size=(N, K)).astype(np.int8 if signed else np.uint8)
    ref = x.astype(np.int64) @ w.astype(np.int64).T
    acc = torch.zeros(B, (N + 3) // 4 * 4, dtype=torch.int32, device=cuda)
    for ksplit in (0, 1, 3):
        acc.zero_()
        Kn.qgemv(torch.from_numpy(x).to(cuda), torch.from_numpy(w).to(cuda), acc, ksplit=ksplit)
        assert np.array_equal(acc[:, :N].cpu().numpy().astype(np.int64), ref), f"ksplit={ksplit}"


@pytest.mark.parametrize("B", [1, 8, 40])
def test_qgemv_epilogues_match_qgemm(cuda, B):
    """Skinny GEMM + epilogue == the big-M kernel (itself oracle-checked in test_qgemm_gpu.py) in all three modes."""
    from mobilequant_b200 import kernels as Kn
    N, K = 1280, 2048
    rng = np.random.default_rng(B)
    a = torch.from_numpy(rng.integers(0, 256, size=(B, K)).astype(np.uint8)).to(cuda)
    b = torch.from_numpy(rng.integers(0, 256, size=(N, K)).astype(np.uint8)).to(cuda)
    ox = 117
    ow = torch.from_numpy(rng.integers(100, 156, size=N).astype(np.int32)).to(cuda)
    rowsum = a.to(torch.int32).sum(1).to(torch.int32)
    c0 = (K * ox * ow.long() - ox * b.long().sum(1)).to(torch.int32)
    sxw = torch.from_numpy((f32(0.02) * rng.uniform(1e-4, 3e-4, size=N).astype(f32)).astype(f32)).to(cuda)
    bias = torch.from_numpy(rng.normal(0, 0.3, size=N).astype(f32)).to(cuda)
    G = N // 128
    so = torch.from_numpy(rng.uniform(0.015, 0.03, size=G).astype(f32)).to(cuda); oo = torch.full((G,), 128.0, device=cuda)
    lut = torch.randn(256, generator=torch.Generator().manual_seed(1)).to(cuda)
    h0 = torch.randn(B, N, generator=torch.Generator().manual_seed(2)).to(cuda)
    acc = torch.zeros(B, N, dtype=torch.int32, device=cuda)
    # QUANT
    rs1 = torch.zeros(B, dtype=torch.int32, device=cuda); rs2 = torch.zeros_like(rs1)
    ref = Kn.qgemm(a, b, rowsum, sxw, ow, c0, Kn.EPI_QUANT, bias=bias, so=so, oo=oo, qgroup=128, qmax=255, out_bits=8, rowsum_out=rs1)
    Kn.qgemv(a, b, acc)
    got = Kn.qgemv_epilogue(acc, B, N, rowsum, sxw, ow, c0, Kn.EPI_QUANT, bias=bias, so=so, oo=oo, qgroup=128, qmax=255, rowsum_out=rs2)
    assert torch.equal(ref, got) and torch.equal(rs1, rs2) and int(acc.abs().max()) == 0
    rs2.zero_(); z = torch.full((B,), 7, dtype=torch.int32, device=cuda)
    got = Kn.qgemv_fused(a, b, acc, rowsum, sxw, ow, c0, Kn.EPI_QUANT, bias=bias, so=so, oo=oo, qgroup=128, qmax=255, rowsum_out=rs2, zero_out=z)
    assert torch.equal(ref, got) and torch.equal(rs1, rs2) and int(acc.abs().max()) == 0 and int(z.abs().max()) == 0
    # ACTMUL
    rs1.zero_(); rs2.zero_()
    ref = Kn.qgemm(a, b, rowsum, sxw, ow, c0, Kn.EPI_ACTMUL, so=so, oo=oo, qgroup=128, qmax=255, lut=lut, s2=0.01, o2=128.0, qmax2=255, rowsum_out=rs1)
    Kn.qgemv(a, b, acc)
    got = Kn.qgemv_epilogue(acc, B, N, rowsum, sxw, ow, c0, Kn.EPI_ACTMUL, so=so, oo=oo, qgroup=128, qmax=255, lut=lut, s2=0.01, o2=128.0, qmax2=255,
                            rowsum_out=rs2)
    assert torch.equal(ref, got) and torch.equal(rs1, rs2) and int(acc.abs().max()) == 0
    rs2.zero_()
    got = Kn.qgemv_fused(a, b, acc, rowsum, sxw, ow, c0, Kn.EPI_ACTMUL, so=so, oo=oo, qgroup=128, qmax=255, lut=lut, s2=0.01, o2=128.0, qmax2=255,
                         rowsum_out=rs2)
    assert torch.equal(ref, got) and torch.equal(rs1, rs2) and int(acc.abs().max()) == 0
    # RESID
    h1, h2 = h0.clone(), h0.clone()
    Kn.qgemm(a, b, rowsum, sxw, ow, c0, Kn.EPI_RESID, so=so[:1], oo=oo[:1], qgroup=N, qmax=65535, resid=h1)
    Kn.qgemv(a, b, acc)
    Kn.qgemv_epilogue(acc, B, N, rowsum, sxw, ow, c0, Kn.EPI_RESID, so=so[:1], oo=oo[:1], qgroup=N, qmax=65535, resid=h2)
    assert torch.equal(h1, h2) and int(acc.abs().max()) == 0
    h3 = h0.clone()
    for _ in range(3):                                   # repeated launches: the arrival counters reset themselves
        h3.copy_(h0)
        Kn.qgemv_fused(a, b, acc, rowsum, sxw, ow, c0, Kn.EPI_RESID, so=so[:1], oo=oo[:1], qgroup=N, qmax=65535, resid=h3)
        assert torch.equal(h1, h3) and int(acc.abs().max()) == 0


@pytest.mark.parametrize("B,T,nh,nkv,hd,rot", [(2, 40, 4, 2, 32, 32), (1, 70, 8, 8, 64, 16), (2, 33, 8, 1, 64, 64), (1, 50, 2, 1, 128, 128),
                                               (1, 37, 8, 1, 256, 256), (1, 300, 8, 2, 64, 64), (1, 1100, 4, 1, 64, 64)])
def test_qattn_decode_matches_oracle_rows(cuda, B, T, nh, nkv, hd, rot):
    """Token by token: RoPE + append + attention of the new row == row t of the oracle's qrope + causal attention."""
    from mobilequant_b200 import kernels as K
    rng = np.random.default_rng(T + hd + nh)
    N = (nh + 2 * nkv) * hd
    qkv = rng.integers(0, 256, size=(B * T, N)).astype(np.uint8)
    qin = [(f32(0.031), f32(120)), (f32(0.027), f32(131)), (f32(0.011), f32(127))]
    qout = [(f32(0.033), f32(125)), (f32(0.029), f32(128)), (f32(0.012), f32(126))]
    cos, sin = ir.rope_tables(T, rot)
    q, k, v = ir.qrope_int(qkv, B, T, nh, nkv, hd, rot, qin, qout, cos, sin)
    smax = 255 * 255 * hd * 0.033 * 0.029 * 0.12
    qs = (f32(2 * smax / 65535), f32(32768), f32(65535)); qp = (f32(1.0 / 65535), f32(0), f32(65535)); qo = (f32(0.7 / 255), f32(128), f32(255))
    ref = ir.qattn_int(q, k, v, nh, nkv, qout[0], qout[1], qout[2], qs, qp, qo).reshape(B, T, nh * hd)
    # the oracle's own one-row (cache) form agrees with its causal form
    assert np.array_equal(ir.qattn_decode_int(q[:, :, T - 1], k, v, nh, nkv, qout[0], qout[1], qout[2], qs, qp, qo), ref[:, T - 1])
    lut = torch.from_numpy(ir.exp_tables(qs[0], hd).view(np.int32)).to(cuda)
    params = [qout[0][1], qout[1][1], qout[2][1], f32(qout[0][0]) * f32(qout[1][0]), qs[0], qs[1], qs[2], qp[0], qp[2],
              f32(qp[0]) * f32(qout[2][0]), qo[0], qo[1]]
    Tmax = T + 3 if T < 1000 else 2100        # the long case also runs 8-CTA clusters (device-position launches size by Tmax)
    kc = torch.zeros(B, nkv, Tmax, hd, dtype=torch.uint8, device=cuda); vc = torch.zeros_like(kc)
    rsk = torch.zeros(B, nkv, Tmax, dtype=torch.int32, device=cuda)
    dcos, dsin = torch.from_numpy(cos).to(cuda), torch.from_numpy(sin).to(cuda)
    dq = torch.from_numpy(qkv.reshape(B, T, N)).to(cuda)
    pos_dev = torch.zeros(1, dtype=torch.int32, device=cuda)
    for t in range(T):
        rs = torch.zeros(B, dtype=torch.int32, device=cuda)
        kw = dict(pos_dev=pos_dev, pos_bound=Tmax - 1) if t % 2 else {}      # alternate host / device position
        pos_dev.fill_(t)
        out = K.qattn_decode(dq[:, t].contiguous(), B, nh, nkv, hd, rot, t, qin, qout, dcos, dsin, kc, vc, rsk, params, lut, rowsum_out=rs, **kw)
        got = out.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, ref[:, t]), f"t={t}: {(got != ref[:, t]).mean():.4f} mismatching"
        assert np.array_equal(rs.cpu().numpy().astype(np.int64), ref[:, t].sum(-1))
    assert np.array_equal(kc[:, :, :T].cpu().numpy().astype(np.int64), k)
    assert np.array_equal(vc[:, :, :T].cpu().numpy().astype(np.int64), v)
    assert np.array_equal(rsk[:, :, :T].cpu().numpy().astype(np.int64), k.sum(-1))


@pytest.mark.parametrize("fused_norm", [True, False])
@pytest.mark.parametrize("tag", MODEL_GOLDENS)
def test_engine_prefill_decode_equals_full_forward(cuda, tag, fused_norm):
    """prefill(T0) + token-by-token decode reproduces, bit for bit, the residual stream of the CPU oracle's full-sequence
    forward at every decoded position, and the logits of the engine's own full forward."""
    from mobilequant_b200.engine import IntEngine
    g = load_golden(f"model_{tag}.pt")
    eng = IntEngine(product_model(g), g["qcfg"], g["act_dict"], cuda)
    eng.fused_resid_norm = fused_norm          # residual epilogues inside the following row norm, or as their own launches
    ids = torch.cat(g["samples"][:2], dim=0).to(cuda)
    B, T = ids.shape
    T0 = T // 2
    im = ir.IntModel(g["state_dict"], g["cfg"], mr.recipe_from_qcfg_json(g["qcfg"]), g["act_dict"])
    cos, sin = ir.rope_tables(T, im.rot, g["cfg"].get("rope_theta", 10000.0))
    for tt in (T, T0):
        eng.set_rope_tables(tt, torch.from_numpy(cos[:tt]), torch.from_numpy(sin[:tt]))
    h_ref = im.backbone(im.embed(ids.cpu().numpy()), B, T, cos, sin)[0].reshape(B, T, -1)
    logits_full = eng(ids)
    cache = eng.new_cache(B, T)
    # the integer path is bit-exact; lm_head stays an fp32 library GEMM (qm:843-845) whose summation order depends on
    # the number of rows, hence the tolerance on logits only
    close = lambda a, b: torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(logits_full.abs().max()))
    logits = eng.prefill(ids[:, :T0], cache)
    assert close(logits, logits_full[:, T0 - 1])
    for t in range(T0, T):
        h = eng._embed(ids[:, t]).contiguous()
        eng.decode_hidden(h, cache)
        cache.length += 1
        assert np.array_equal(h.cpu().numpy(), h_ref[:, t]), f"position {t}"
        assert close(eng._head(h), logits_full[:, t])


def test_engine_generate_and_graph_replay(cuda):
    """generate() (greedy) == argmax chain of full forwards; the CUDA-graph replay of the decode step emits the same tokens."""
    from mobilequant_b200.engine import IntEngine
    g = load_golden("model_llama_w8_e2e.pt")
    eng = IntEngine(product_model(g), g["qcfg"], g["act_dict"], cuda)
    ctx = torch.cat(g["samples"][:2], dim=0).to(cuda)[:, :12]
    new = 6
    out = eng.generate(ctx, new)
    assert out.shape == (2, 12 + new)
    ids = ctx
    for _ in range(new):
        nxt = eng(ids)[:, -1].argmax(-1)
        ids = torch.cat([ids, nxt.view(-1, 1)], dim=1)
    assert torch.equal(out, ids)
    # graph replay
    cache = eng.new_cache(2, 12 + new)
    logits = eng.prefill(ctx, cache)
    graph, tokens, _ = eng.capture_decode(cache)
    tokens.copy_(logits.argmax(-1))
    got = [tokens.clone()]
    for _ in range(new - 1):
        graph.replay()
        got.append(tokens.clone())
    assert torch.equal(torch.stack(got, dim=1), ids[:, 12:])
    # capacity guard: the wrapper refuses to step past Tmax and keeps cache.length in step with the device position;
    # a raw replay past the capacity is a no-op for the caches (the kernel exits on pos > pos_bound)
    assert cache.length == 12 + new - 1 and int(cache.pos_dev.item()) == cache.length
    graph.replay()
    assert cache.length == cache.Tmax
    with pytest.raises(RuntimeError, match="KV cache is full"):
        graph.replay()
    snap = [t.clone() for t in cache.k + cache.v + cache.rsk]
    graph.graph.replay()                      # what a caller bypassing the wrapper would do
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(snap, cache.k + cache.v + cache.rsk))


def test_fgemv_lm_head(cuda):
    from mobilequant_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    for B, V, Kd in ((1, 1000, 256), (8, 32000, 2048), (13, 777, 512)):
        x = torch.randn(B, Kd, generator=g).to(cuda); w = (torch.randn(V, Kd, generator=g) * 0.02).to(cuda)
        ref = (x.double() @ w.double().t()).float()
        got = K.fgemv(x, w)
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)
        assert torch.equal(got, K.fgemv(x, w))          # deterministic
