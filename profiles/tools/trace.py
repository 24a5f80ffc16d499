"""GPU timeline of a few bench steps (torch.profiler / CUPTI): where does the device idle?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import __graft_entry__ as ge
import scenes
from torch.profiler import profile, ProfilerActivity

pkg = ge.load_package(); pkg.load()
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc = scenes.scene_b(N=N)
W, H, deg = sc["width"], sc["height"], sc["sh_degree"]
names = ("means", "quats", "scales", "opacities", "sh_coeffs")
P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(k in names) for k in
     ("means", "quats", "scales", "opacities", "sh_coeffs", "viewmats", "Ks", "background")}
P["target"] = torch.rand((1, H, W, 3), device=dev)

def step():
    for k in names: P[k].grad = None
    out = pkg.rasterize(P["means"], P["quats"], P["scales"], P["opacities"], P["sh_coeffs"], deg, P["viewmats"], P["Ks"],
                        W, H, bg_color=P["background"])
    loss = (out.render_colors - P["target"]).abs().mean()
    loss.backward()

for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4): step()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/trace.json")
tr = json.load(open("gpurun_out/trace.json"))
ev = [e for e in tr["traceEvents"] if e.get("ph") == "X"]
gpu = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
cpu = sorted([e for e in ev if e.get("cat") in ("cpu_op", "cuda_runtime", "user_annotation", "python_function")], key=lambda e: e["ts"])
t0, t1 = gpu[0]["ts"], gpu[-1]["ts"] + gpu[-1]["dur"]
busy = sum(e["dur"] for e in gpu)
print(f"span {(t1 - t0) / 4e3:.3f} ms/step, gpu busy {busy / 4e3:.3f} ms/step, {len(gpu) / 4:.0f} gpu activities/step")
# gaps of the LAST profiled step
gaps = []
for a, b in zip(gpu[:-1], gpu[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 4: gaps.append((g, a["name"][:50], b["name"][:50], a["ts"]))
third = t0 + (t1 - t0) * 0.5
print("gaps > 4us in the second half of the trace:")
tot = 0
for g, an, bn, ts in gaps:
    if ts > third:
        tot += g
        print(f"  {g:8.1f} us  after {an:50s} before {bn}")
print(f"  total {tot / 2e3:.3f} ms/step")
# per-name GPU time
agg = {}
for e in gpu: agg[e["name"][:60]] = agg.get(e["name"][:60], 0) + e["dur"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:30]: print(f"  {v / 4e3:8.4f} ms/step  {k}")
# CPU-side: top-level cpu ops durations
cagg = {}
for e in ev:
    if e.get("cat") == "cpu_op":
        cagg[e["name"][:60]] = cagg.get(e["name"][:60], [0, 0]); cagg[e["name"][:60]][0] += e["dur"]; cagg[e["name"][:60]][1] += 1
print("cpu ops (total us/step, calls/step):")
for k, v in sorted(cagg.items(), key=lambda kv: -kv[1][0])[:30]: print(f"  {v[0] / 4:8.1f} us {v[1] / 4:5.1f}  {k}")
os.remove("gpurun_out/trace.json")
