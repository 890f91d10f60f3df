"""Back-to-back launch timing of the decode-step kernels at model shapes (approximates in-graph cost)."""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from mobilequant_b200 import kernels as K
from oracle import int_ref as ir
cuda = torch.device("cuda:0"); f32 = np.float32
def timeit(name, fn, bytes_=None, iters=20):
    """GPU time per launch: `iters` launches captured in one CUDA graph (no host launch overhead), 5 replays."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * iters) * 1e3
    extra = f"  {bytes_/us/1e3:.0f} GB/s" if bytes_ else ""
    print(f"{name:44s} {us:8.2f} us{extra}", flush=True)
def gemv_suite(tag, B, H, I, nh, nkv, hd, V, T):
    print(f"== {tag} B={B}")
    Ipad = (I + 127) // 128 * 128
    shapes = [("qkv", (nh + 2 * nkv) * hd, H), ("o", H, nh * hd), ("w13", 2 * Ipad, H), ("w2", H, Ipad)]
    nmax = max(n for _, n, _ in shapes)
    acc = torch.zeros(B, nmax, dtype=torch.int32, device=cuda)
    for nm, N, Kd in shapes:
        # rotate through 8 weight copies so that the weights are not L2 resident (126 MB L2)
        ws = [torch.randint(0, 256, (N, Kd), dtype=torch.uint8, device=cuda) for _ in range(max(2, int(300e6 // (N * Kd))))]
        x = torch.randint(0, 256, (B, Kd), dtype=torch.uint8, device=cuda)
        it = [0]
        def f():
            K.qgemv(x, ws[it[0] % len(ws)], acc); it[0] += 1
        timeit(f"qgemv {nm} N={N} K={Kd}", f, N * Kd)
        rowsum = x.to(torch.int32).sum(1).to(torch.int32); sxw = torch.full((N,), 1e-5, device=cuda)
        ow = torch.full((N,), 128, dtype=torch.int32, device=cuda); c0 = torch.zeros(N, dtype=torch.int32, device=cuda)
        G = (N + 127) // 128; so = torch.full((G,), 0.05, device=cuda); oo = torch.full((G,), 128.0, device=cuda)
        if nm == "w13":
            lut = torch.randn(256, device=cuda); out = torch.empty(B, N // 2, dtype=torch.uint8, device=cuda)
            timeit(f"  epilogue ACTMUL", lambda: K.qgemv_epilogue(acc, B, N, rowsum, sxw, ow, c0, K.EPI_ACTMUL, so=so, oo=oo, qgroup=128, lut=lut, s2=0.01, o2=128.0, out=out))
        elif nm == "qkv":
            out = torch.empty(B, N, dtype=torch.uint8, device=cuda)
            timeit(f"  epilogue QUANT", lambda: K.qgemv_epilogue(acc, B, N, rowsum, sxw, ow, c0, K.EPI_QUANT, so=so, oo=oo, qgroup=128, out=out))
        else:
            h = torch.zeros(B, N, device=cuda)
            timeit(f"  epilogue RESID", lambda: K.qgemv_epilogue(acc, B, N, rowsum, sxw, ow, c0, K.EPI_RESID, so=so[:1], oo=oo[:1], qgroup=128 * G, qmax=65535, resid=h))
        del ws
    # qnorm
    h = torch.randn(B, H, device=cuda); w = torch.randn(H, device=cuda)
    qin = (f32(12.0 / 65535), f32(32768), f32(65535)); qout = (f32(8.0 / 255), f32(128), f32(255))
    codes = torch.empty(B, H, dtype=torch.uint8, device=cuda); rs = torch.empty(B, dtype=torch.int32, device=cuda)
    timeit("qnorm", lambda: K.qnorm(h, qin, w, None, qout, False, 1e-5, codes, rs))
    # attention decode
    N = (nh + 2 * nkv) * hd
    qkv = torch.randint(0, 256, (B, N), dtype=torch.uint8, device=cuda)
    kc = torch.randint(0, 256, (B, nkv, T, hd), dtype=torch.uint8, device=cuda); vc = torch.randint(0, 256, (B, nkv, T, hd), dtype=torch.uint8, device=cuda)
    rsk = kc.to(torch.int32).sum(-1).to(torch.int32)
    qi = [(f32(0.031), f32(120)), (f32(0.027), f32(131)), (f32(0.011), f32(127))]
    qo = [(f32(0.033), f32(125)), (f32(0.029), f32(128)), (f32(0.012), f32(126))]
    smax = 255 * 255 * hd * 0.033 * 0.029 * 0.12
    qs = (f32(2 * smax / 65535), f32(32768), f32(65535))
    lut = torch.from_numpy(ir.exp_tables(qs[0], hd).view(np.int32)).to(cuda)
    params = [qo[0][1], qo[1][1], qo[2][1], f32(qo[0][0]) * f32(qo[1][0]), qs[0], qs[1], qs[2], f32(1.0 / 65535), f32(65535),
              f32(1.0 / 65535) * f32(qo[2][0]), f32(0.7 / 255), f32(128)]
    cos, sin = ir.rope_tables(T, hd)
    dcos, dsin = torch.from_numpy(cos).to(cuda), torch.from_numpy(sin).to(cuda)
    out = torch.empty(B, nh * hd, dtype=torch.uint8, device=cuda); rso = torch.zeros(B, dtype=torch.int32, device=cuda)
    import os
    for cs in ("2", "4", "8"):
        os.environ["MQB200_DEC_CS"] = cs
        timeit(f"qattn_decode pos={T-1} max cluster {cs}", lambda: K.qattn_decode(qkv, B, nh, nkv, hd, hd, T - 1, qi, qo, dcos, dsin, kc, vc, rsk, params, lut, out=out, rowsum_out=rso),
               2 * B * nkv * T * hd)
    os.environ.pop("MQB200_DEC_CS", None)
    timeit(f"qattn_decode pos={T-1}", lambda: K.qattn_decode(qkv, B, nh, nkv, hd, hd, T - 1, qi, qo, dcos, dsin, kc, vc, rsk, params, lut, out=out, rowsum_out=rso),
           2 * B * nkv * T * hd)
    # lm_head
    x = torch.randn(B, H, device=cuda); wh = torch.randn(V, H, device=cuda) * 0.02
    timeit("fgemv lm_head", lambda: K.fgemv(x, wh), V * H * 4)
    timeit("torch linear lm_head", lambda: torch.nn.functional.linear(x, wh), V * H * 4)
gemv_suite("tinyllama", 8, 2048, 5632, 32, 4, 64, 32000, 1024)
gemv_suite("gemma-2b", 8, 2048, 16384, 8, 1, 256, 256000, 2048)
