"""Time of bl_gemm_f32 on the learner's shapes (c2 batch) next to torch.matmul fp32 / TF32: python tools/gemm_time.py [n_samples]"""
import sys
import torch
sys.path.insert(0, '.')
from boardlaw_b200.learner import gemm, amax

Nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(s, device=dev, generator=g)


def timed(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


x, w, w0, obs, wp, du = r(Nb, 256), r(256, 256), r(256, 162), r(Nb, 162), r(81, 256), r(Nb, 256)
ds = r(Nb, 81)
cases = [('fwd body   x.W^T', x, w, lambda: x @ w.t()), ('fwd in     obs.W0^T', obs, w0, lambda: obs @ w0.t()), ('fwd policy x.Wp^T', x, wp, lambda: x @ wp.t()),
         ('dgrad      du.W', du, w.t(), lambda: du @ w), ('dgrad pol  ds.Wp', ds, wp.t(), lambda: ds @ wp),
         ('wgrad body du^T.x', du.t(), x.t(), lambda: du.t() @ x), ('wgrad in   du^T.obs', du.t(), obs.t(), lambda: du.t() @ obs),
         ('wgrad pol  ds^T.x', ds.t(), x.t(), lambda: ds.t() @ x)]
for name, a, b, ref in cases:
    am, bm = amax(a), amax(b)
    t = timed(lambda: gemm(a, b, a_amax=am, b_amax=bm))
    torch.backends.cuda.matmul.allow_tf32 = False
    t32 = timed(ref)
    torch.backends.cuda.matmul.allow_tf32 = True
    ttf = timed(ref)
    fl = 2 * a.shape[0] * b.shape[0] * a.shape[1]
    print(f'{name:22s} M{a.shape[0]:6d} N{b.shape[0]:4d} K{a.shape[1]:6d}: tcgen05 split {t:7.1f} us ({fl / t / 1e6:6.1f} TFLOP/s useful)   cuBLAS fp32 {t32:7.1f} us   TF32 {ttf:7.1f} us')
