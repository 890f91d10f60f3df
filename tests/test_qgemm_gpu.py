"""tcgen05 integer GEMM + fused epilogues (mq_qgemm) against the exact-integer oracle (oracle/int_ref.py)."""
import numpy as np
import pytest
import torch
from oracle import int_ref as ir

pytestmark = pytest.mark.gpu


def _prep(M, N, K, seed, a_signed=False, b_signed=False, per_channel=False):
    rng = np.random.default_rng(seed)
    a = rng.integers(-128 if a_signed else 0, 128 if a_signed else 256, size=(M, K)).astype(np.int8 if a_signed else np.uint8)
    b = rng.integers(-128 if b_signed else 0, 128 if b_signed else 256, size=(N, K)).astype(np.int8 if b_signed else np.uint8)
    ox = 0 if a_signed else int(rng.integers(0, 256))
    ow = np.zeros(N, np.int64) if b_signed else (rng.integers(0, 256, size=N) if per_channel else np.full(N, rng.integers(0, 256)))
    sx = np.float32(rng.uniform(0.01, 0.05))
    sw = (rng.uniform(1e-4, 3e-4, size=N) if per_channel else np.full(N, rng.uniform(1e-4, 3e-4))).astype(np.float32)
    return a, b, ox, ow.astype(np.int64), sx, sw


def _dev(a, b, ox, ow, sx, sw, K, cuda):
    ta, tb = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
    rowsum = ta.to(torch.int32).sum(1).to(torch.int32)
    colsum = torch.from_numpy(b.astype(np.int64).sum(1))
    c0 = (K * ox * torch.from_numpy(ow) - ox * colsum).to(torch.int32).to(cuda)
    sxw = torch.from_numpy((np.float32(sx) * sw).astype(np.float32)).to(cuda)
    return ta, tb, rowsum, sxw, torch.from_numpy(ow).to(torch.int32).to(cuda), c0


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (256, 512, 2048), (100, 96, 352), (1000, 2560, 2048), (384, 2048, 5632)])
@pytest.mark.parametrize("signed", [False, True])
def test_qgemm_exact_integer(cuda, M, N, K, signed):
    from mobilequant_b200 import kernels as Kn
    a, b, ox, ow, sx, sw = _prep(M, N, K, M + N + K, b_signed=signed)
    ta, tb, rowsum, sxw, tow, c0 = _dev(a, b, ox, ow, sx, sw, K, cuda)
    out = Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_I32)
    ref = ir.int_acc(a, b, ox, ow)
    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)


@pytest.mark.parametrize("bits", [8, 16])
def test_qgemm_quant_epilogue(cuda, bits):
    from mobilequant_b200 import kernels as Kn
    M, N, K = 300, 768, 1024
    a, b, ox, ow, sx, sw = _prep(M, N, K, 7 + bits, per_channel=True)
    ta, tb, rowsum, sxw, tow, c0 = _dev(a, b, ox, ow, sx, sw, K, cuda)
    rng = np.random.default_rng(3)
    bias = rng.normal(0, 0.5, size=N).astype(np.float32)
    y = ir.qlinear_y(a, sx, ox, b, sw, ow, bias)
    qmax = 2 ** bits - 1
    so = np.full(N, (y.max() - y.min()) / qmax * 0.9, np.float32); so[N // 2:] *= np.float32(1.3)   # two output quantizers
    oo = np.rint(-y.min() / so).astype(np.float32)
    ref = ir.quant_codes(y, so.reshape(1, -1), oo.reshape(1, -1), 0, qmax).astype(np.int64)
    rs_out = torch.zeros(M, dtype=torch.int32, device=cuda)
    out = Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_QUANT, bias=torch.from_numpy(bias).to(cuda),
                   so=torch.from_numpy(so[::N // 2].copy()).to(cuda), oo=torch.from_numpy(oo[::N // 2].copy()).to(cuda), qgroup=N // 2,
                   qmax=qmax, out_bits=bits, rowsum_out=rs_out)
    got = out.cpu().numpy()
    got = got.astype(np.int64) if bits == 8 else got.view(np.uint16).astype(np.int64)
    assert np.array_equal(got, ref)
    assert np.array_equal(rs_out.cpu().numpy().astype(np.int64), ref.sum(1))


def test_qgemm_resid_epilogue(cuda):
    from mobilequant_b200 import kernels as Kn
    M, N, K = 260, 512, 768
    a, b, ox, ow, sx, sw = _prep(M, N, K, 11)
    ta, tb, rowsum, sxw, tow, c0 = _dev(a, b, ox, ow, sx, sw, K, cuda)
    y = ir.qlinear_y(a, sx, ox, b, sw, ow)
    so = np.float32((y.max() - y.min()) / 65535); oo = np.float32(np.rint(-y.min() / so))
    rng = np.random.default_rng(5)
    h = rng.normal(0, 1, size=(M, N)).astype(np.float32)
    ref = (h + ir.dequant(ir.quant_codes(y, so, oo, 0, 65535), so, oo)).astype(np.float32)
    th = torch.from_numpy(h.copy()).to(cuda)
    Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_RESID, so=torch.full((1,), float(so), device=cuda),
             oo=torch.full((1,), float(oo), device=cuda), qmax=65535, resid=th)
    assert np.array_equal(th.cpu().numpy(), ref)


def test_qgemm_actmul_epilogue(cuda):
    """Fused w1||w3 GEMM + activation LUT * gate + w2 input quantizer."""
    from mobilequant_b200 import kernels as Kn
    M, I, K = 200, 384, 512
    rng = np.random.default_rng(21)
    a, b1, ox, ow1, sx, sw1 = _prep(M, I, K, 31)
    _, b3, _, ow3, _, sw3 = _prep(M, I, K, 32)
    y1 = ir.qlinear_y(a, sx, ox, b1, sw1, ow1); y3 = ir.qlinear_y(a, sx, ox, b3, sw3, ow3)
    so1 = np.float32((y1.max() - y1.min()) / 255); oo1 = np.float32(np.rint(-y1.min() / so1))
    so3 = np.float32((y3.max() - y3.min()) / 255); oo3 = np.float32(np.rint(-y3.min() / so3))
    lut = rng.normal(0, 1, size=256).astype(np.float32)
    q1 = ir.quant_codes(y1, so1, oo1, 0, 255).astype(np.int64)
    u = ir.dequant(ir.quant_codes(y3, so3, oo3, 0, 255), so3, oo3)
    prod = (lut[q1] * u).astype(np.float32)
    s2 = np.float32((prod.max() - prod.min()) / 255); o2 = np.float32(np.rint(-prod.min() / s2))
    ref = ir.quant_codes(prod, s2, o2, 0, 255).astype(np.int64)
    # interleave per 256-row tile: [128 rows w1 | 128 rows w3]
    nb = I // 128
    b = np.concatenate([np.concatenate([b1[i * 128:(i + 1) * 128], b3[i * 128:(i + 1) * 128]]) for i in range(nb)])
    il = lambda v1, v3: np.concatenate([np.concatenate([v1[i * 128:(i + 1) * 128], v3[i * 128:(i + 1) * 128]]) for i in range(nb)])
    ow = il(ow1, ow3); sw = il(sw1, sw3)
    ta, tb, rowsum, sxw, tow, c0 = _dev(a, b, ox, ow, sx, sw, K, cuda)
    so = torch.from_numpy(il(np.full(I, so1, np.float32), np.full(I, so3, np.float32))[::128].copy()).to(cuda)   # per 128-column half tile
    oo = torch.from_numpy(il(np.full(I, oo1, np.float32), np.full(I, oo3, np.float32))[::128].copy()).to(cuda)
    rs_out = torch.zeros(M, dtype=torch.int32, device=cuda)
    out = Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_ACTMUL, so=so, oo=oo, qgroup=128, qmax=255, lut=torch.from_numpy(lut).to(cuda),
                   s2=float(s2), o2=float(o2), qmax2=255, rowsum_out=rs_out)
    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(rs_out.cpu().numpy().astype(np.int64), ref.sum(1))


@pytest.mark.parametrize("ne", [8, 16])
@pytest.mark.parametrize("cl", [1, 2, 4])
@pytest.mark.parametrize("M,N,K", [(100, 96, 352), (1300, 1280, 2048), (2048, 512, 5632)])
def test_qgemm_cluster_variants(cuda, cl, ne, M, N, K, monkeypatch):
    """Every shape of the kernel (single CTA, CTA pair, two pairs with multicast B; 8 or 16 epilogue warps) gives the
    same exact integers, raw and through each fused epilogue; the small shape is checked against the oracle directly."""
    from mobilequant_b200 import kernels as Kn
    a, b, ox, ow, sx, sw = _prep(M, N, K, M + N + K + 5, per_channel=True)
    ta, tb, rowsum, sxw, tow, c0 = _dev(a, b, ox, ow, sx, sw, K, cuda)
    G = (N + 127) // 128
    so = torch.full((G,), 0.02, device=cuda); oo = torch.full((G,), 128.0, device=cuda)
    lut = torch.randn(256, generator=torch.Generator().manual_seed(1)).to(cuda)
    h0 = torch.randn(M, N, generator=torch.Generator().manual_seed(2)).to(cuda)

    def run_all():
        outs = [Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_I32)]
        rs = torch.zeros(M, dtype=torch.int32, device=cuda)
        outs.append(Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_QUANT, so=so, oo=oo, qgroup=128, qmax=255, out_bits=8, rowsum_out=rs))
        outs.append(rs)
        h = h0.clone()
        Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_RESID, so=so[:1], oo=oo[:1], qgroup=128 * G, qmax=65535, resid=h)
        outs.append(h)
        if N % 256 == 0:
            outs.append(Kn.qgemm(ta, tb, rowsum, sxw, tow, c0, Kn.EPI_ACTMUL, so=so, oo=oo, qgroup=128, qmax=255, lut=lut,
                                 s2=0.01, o2=128.0, qmax2=255))
        torch.cuda.synchronize()
        return outs

    monkeypatch.delenv("MQ_QGEMM_CL", raising=False)
    monkeypatch.delenv("MQ_QGEMM_NE", raising=False)
    base = run_all()
    if M <= 128:
        assert np.array_equal(base[0].cpu().numpy().astype(np.int64), ir.int_acc(a, b, ox, ow))
    monkeypatch.setenv("MQ_QGEMM_CL", str(cl))
    monkeypatch.setenv("MQ_QGEMM_NE", str(ne))
    got = run_all()
    for x, y in zip(base, got):
        assert torch.equal(x, y)


@pytest.mark.parametrize("M,N,K,mode", [(300, 512, 256, "quant"), (1024, 2560, 2048, "quant"), (257, 768, 1056, "i32"), (512, 2048, 5632, "resid"),
                                        (384, 1024, 2048, "actmul")])
def test_qgemm_w4a8_packed_equals_unpacked(cuda, M, N, K, mode):
    """mq_qgemm_w4a8 (nibbles expanded inside the kernel) == mq_qgemm on the one-code-per-byte matrix, bit for bit, for every
    epilogue, ragged M, K not a multiple of the 128-code k-slice, multi-tile N; plus the exact int64 accumulator."""
    from mobilequant_b200 import kernels as K_
    rng = np.random.default_rng(M + N + K)
    a = rng.integers(0, 256, size=(M, K)).astype(np.uint8)
    wu = rng.integers(0, 16, size=(N, K)).astype(np.uint8)                   # unsigned 4-bit codes (offset-binary for symmetric weights)
    packed = (wu[:, 0::2] | (wu[:, 1::2] << 4)).astype(np.uint8)
    ow = rng.integers(6, 10, size=N).astype(np.int32)
    ox = 117
    colsum = wu.astype(np.int64).sum(1)
    c0 = (K * ox * ow.astype(np.int64) - ox * colsum).astype(np.int32)
    sxw = (rng.uniform(0.5, 1.5, size=N) * 1e-3).astype(np.float32)
    dev = lambda x: torch.from_numpy(x).to(cuda)
    A, Wd, Pd = dev(a), dev(wu), dev(packed)
    rowsum = dev(a.astype(np.int64).sum(1).astype(np.int32))
    common = dict(rowsum=rowsum, sxw=dev(sxw), ow=dev(ow), c0=dev(c0))
    if mode == "i32":
        ref = K_.qgemm(A, Wd, mode=K_.EPI_I32, **common)
        got = K_.qgemm(A, Pd, mode=K_.EPI_I32, packed4=True, **common)
        exact = (a.astype(np.int64) - ox) @ (wu.astype(np.int64) - ow[:, None]).T
        assert np.array_equal(got.cpu().numpy().astype(np.int64), exact)
    elif mode == "quant":
        so, oo = torch.tensor([0.02], device=cuda), torch.tensor([131.0], device=cuda)
        rs1 = torch.zeros(M, dtype=torch.int32, device=cuda); rs2 = torch.zeros_like(rs1)
        ref = K_.qgemm(A, Wd, mode=K_.EPI_QUANT, so=so, oo=oo, rowsum_out=rs1, **common)
        got = K_.qgemm(A, Pd, mode=K_.EPI_QUANT, so=so, oo=oo, rowsum_out=rs2, packed4=True, **common)
        assert torch.equal(rs1, rs2)
    elif mode == "resid":
        so, oo = torch.tensor([3e-4], device=cuda), torch.tensor([32768.0], device=cuda)
        h0 = torch.randn(M, N, generator=torch.Generator().manual_seed(1)).to(cuda)
        ref = K_.qgemm(A, Wd, mode=K_.EPI_RESID, so=so, oo=oo, qmax=65535.0, resid=h0.clone(), **common)
        got = K_.qgemm(A, Pd, mode=K_.EPI_RESID, so=so, oo=oo, qmax=65535.0, resid=h0.clone(), packed4=True, **common)
    else:
        so = torch.full((N // 128,), 0.02, device=cuda); oo = torch.full((N // 128,), 128.0, device=cuda)
        lut = torch.linspace(-1, 3, 256, device=cuda)
        ref = K_.qgemm(A, Wd, mode=K_.EPI_ACTMUL, so=so, oo=oo, lut=lut, s2=0.05, o2=120.0, **common)
        got = K_.qgemm(A, Pd, mode=K_.EPI_ACTMUL, so=so, oo=oo, lut=lut, s2=0.05, o2=120.0, packed4=True, **common)
    assert torch.equal(got, ref)
