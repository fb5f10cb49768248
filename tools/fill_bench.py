"""Constant-texture fill against torch.fill_ on the same buffers: two buffers of each size, alternated, so that every
launch writes memory that is not in L2 (2 x 400 MB / 2 x 1.2 GB against 126 MB)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pbrt_b200 as pb

pb.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); pb.set_stream(stream.cuda_stream)
n = 100_000_000
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6545.0

def timed(fn, reps=20):
    for i in range(4): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for i in range(reps): fn(i)
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {"variant": os.environ.get("PBRT_B200_FILL_VARIANT", "0"), "ctas": os.environ.get("PBRT_B200_FILL_CTAS", "64")}
f = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
tex = pb.ConstantTexture(10.0)
ms = timed(lambda i: tex.evaluate_batch(n, out=f[i & 1])); out["f32_ours_gbs"] = n * 4 / ms / 1e6
ms = timed(lambda i: f[i & 1].fill_(10.0)); out["f32_torch_gbs"] = n * 4 / ms / 1e6
del f
g = [torch.empty((n, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
tex3 = pb.ConstantTexture((1.0, 0.0, 0.0))
ms = timed(lambda i: tex3.evaluate_batch(n, out=g[i & 1])); out["rgb_ours_gbs"] = n * 12 / ms / 1e6
assert bool((g[0][:5] == torch.tensor([1.0, 0.0, 0.0], device="cuda")).all()) and bool((g[1][-1] == torch.tensor([1.0, 0.0, 0.0], device="cuda")).all())
assert float(g[0].sum(dtype=torch.float64)) == n and float(g[1][:, 1:].abs().sum()) == 0.0
ms = timed(lambda i: g[i & 1].fill_(1.0)); out["rgb_torch_fill_scalar_gbs"] = n * 12 / ms / 1e6
out["peak_gbs"] = peak
print(json.dumps(out))
