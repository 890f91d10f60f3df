"""Weight-pass kernels (mq_wprep_fwd / mq_wprep_bwd) at TinyLlama weight shapes: device time per call from a CUDA graph of 20 calls."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobilequant_b200 import kernels as K
dev = torch.device("cuda")
torch.manual_seed(0)
def graph_time(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * n)
for rows, cols, per_ch in [(2048, 2048, False), (11264 // 2, 2048, False), (2048, 5632, True), (2560, 2048, False)]:
    w = torch.randn(rows, cols, device=dev) * 0.05
    g = torch.randn(rows, cols, device=dev) * 1e-3
    cf = torch.rand(cols, device=dev) + 0.5
    groups = rows if per_ch else 1
    su = torch.sigmoid(torch.full((groups,), 4.0, device=dev)); sl = su.clone()
    out = K.wprep_fwd(w, 8, False, per_ch, cf, 2, None, 0, su, sl)
    tf = graph_time(lambda: K.wprep_fwd(w, 8, False, per_ch, cf, 2, None, 0, su, sl))
    tb = graph_time(lambda: K.wprep_bwd(w, g, 8, False, per_ch, cf, 2, None, 0, su, sl, minmax=out["minmax"]))
    mb = rows * cols * 4 / 1e6
    print(f"[{rows} x {cols}] per_channel={per_ch}: fwd {tf:.1f} us ({3 * mb / tf:.2f} TB/s of 12 B/elem)  bwd {tb:.1f} us ({6 * mb / tb:.2f} TB/s of 24 B/elem)")
