"""profiles/<tag>_bench.md from bench.py's JSON line(s): every table of the round's measurement story is re-derived from
the line the driver gets, nothing is typed by hand.
    python profiles/tools/bench_report.py r2 gpurun_out/r2h_bench.json [gpurun_out/r2h_scale_n*.json ...]"""
import json
import sys

tag, main_path, rest = sys.argv[1], sys.argv[2], sys.argv[3:]
d = json.loads([l for l in open(main_path).read().splitlines() if l.startswith("{")][-1])
out = []
w = out.append


def f(x, n=3):
    return "—" if x is None else (f"{x:.{n}f}" if isinstance(x, float) else str(x))


cfg = d["config"]
w(f"# Bench report {tag}\n")
w(f"Source: `{main_path}` (one `python bench.py` run on a B200 through `gpurun`; CUDA-event timing, max over ranks; "
  f"SM clock {d['clocks']['sm_mhz']:.0f} / {d['clocks']['sm_max_mhz']:.0f} MHz, throttle reasons {d['clocks']['reasons']}).\n")
w(f"Workload: {cfg['workload']} — {cfg['gaussians']} Gaussians, {cfg['visible']} visible, {cfg['intersections']} "
  f"intersections, longest tile list {cfg.get('max_tile_list')}.\n")
w("## Headline (config B, forward + backward, one view)\n")
w("| protocol | this repo ms/step | reference kernels ms/step | ratio |\n|---|---:|---:|---:|")
r = d.get("reference_cuda", {})
sp = r.get("speedup", {})
w(f"| operator API, device-resident (`value`) | {f(d['ms_per_step'])} | {f(r.get('ms_per_step'))} | {f(sp.get('device_resident'), 2)}× |")
w(f"| `e2e`: every input from pinned host memory, image + gradients back | {f(d['e2e']['ms_per_step'])} | {f(r.get('e2e_ms'))} | {f(sp.get('e2e'), 2)}× |")
w(f"| `e2e_resident`: parameters resident, camera + target in, loss out | {f(d['e2e_resident']['ms_per_step'])} | {f(r.get('e2e_resident_ms'))} | {f(sp.get('e2e_resident'), 2)}× |")
fu = d.get("fused", {})
if "ms_per_step" in fu:
    rr = r.get("raw_tensor_sequence", {})
    w(f"| raw SplatData in, raw gradients out (extended operator vs torch glue + operators), device-resident | "
      f"{f(fu['ms_per_step'])} | {f(rr.get('ms_per_step'))} | {f(rr.get('ms_per_step') / fu['ms_per_step'] if rr.get('ms_per_step') else None, 2)}× |")
    w(f"| the same, `e2e_resident` | {f(fu['e2e_resident_ms'])} | {f(rr.get('e2e_resident_ms'))} | "
      f"{f(sp.get('fused_vs_reference_training_step_e2e_resident'), 2)}× |")
    u = fu.get("operator_path_on_raw_tensors", {})
    w(f"| (this backend's eleven operators + torch activations on the raw tensors) | {f(u.get('ms_per_step'))} | | |")
w(f"\nValue: {d['value'] / 1e6:.1f} M Gaussians/s ({d['intersections_per_sec'] / 1e9:.2f} G intersections/s); distribution of "
  f"{d['distribution']['n']} steps: median {f(d['distribution']['median_ms'])}, p10 {f(d['distribution']['p10_ms'])}, "
  f"p90 {f(d['distribution']['p90_ms'])} ms; {d['gpu_launches']} own kernel launches in {d['steps']} steps.\n")
tr = d.get("train", {})
if "b200" in tr:
    w("## Training iteration (render + SSIM/L1 loss + backward + Adam; camera + target H2D, loss D2H each iteration)\n")
    w("| | ms/iteration | iterations/s |\n|---|---:|---:|")
    w(f"| this repo (extended operator + fused loss + one-launch Adam) | {f(tr['b200']['ms_per_iter'])} | {f(tr['b200']['iters_per_sec'], 1)} |")
    g = tr.get("b200_cuda_graph", {})
    if "ms_per_iter" in g:
        w(f"| the same iteration replayed from one CUDA graph | {f(g['ms_per_iter'])} | {f(g['iters_per_sec'], 1)} |")
    rc = tr.get("reference_cuda", {})
    if "ms_per_iter" in rc:
        w(f"| reference kernels (gsplat operators, ssim.cu, adam_kernels.cuh) glued by torch as its trainer does | {f(rc['ms_per_iter'])} | {f(rc['iters_per_sec'], 1)} |")
        w(f"\nSpeed-up {f(tr.get('speedup'), 2)}× (graph: {f(tr.get('speedup_cuda_graph'), 2)}×).\n")
w("## Per operator (CUDA events around each call of the L3 sequence)\n")
ops = d.get("ops_ms", {})
if ops:
    names = list(ops.get("b200", {}).keys())
    w("| operator | this repo ms | reference kernels ms |\n|---|---:|---:|")
    for n in names:
        w(f"| `{n}` | {f(ops['b200'].get(n), 4)} | {f(ops.get('reference_cuda', {}).get(n), 4)} |")
w("\n## Per kernel (library's own event profile, separate pass)\n")
w("| kernel scope | operator path ms | extended path ms | training iteration ms |\n|---|---:|---:|---:|")
k1, k2, k3 = d.get("kernels_ms", {}), fu.get("kernels_ms", {}), tr.get("b200", {}).get("kernels_ms", {})
for n in list(dict.fromkeys(list(k1) + list(k2) + list(k3))):
    w(f"| `{n}` | {f(k1.get(n), 4)} | {f(k2.get(n), 4)} | {f(k3.get(n), 4)} |")
ro = d.get("roofline")
if ro:
    w(f"\n## Roofline of the dominant kernel\n\n`{ro['kernel']}`: {ro['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic per launch / "
      f"{ro['avg_kernel_ms']:.4f} ms = {ro['achieved']:.0f} GB/s of {ro['peak']:.0f} GB/s ({ro['peak_source']}) = "
      f"**{ro['frac']:.3f}**; DRAM traffic per launch from ncu: {ro['traffic']} B ({ro['traffic_source']}).  {ro['note']}\n")
cb = d.get("cpu_baseline")
if cb:
    w(f"## CPU baseline\n\n{cb['value']:.0f} {cb['unit']} on {cb['cores']} cores ({cb['kind']}); sample: {cb['sample']}.\n")
cs = d.get("configs", {})
if cs:
    w("## Other configs (same run)\n")
    a = cs.get("A", {})
    if a:
        w(f"**A** ({a['workload']}): CPU oracle {f(a['cpu_oracle']['ms_per_step'], 1)} ms/step on {a['cpu_oracle']['cores']} cores "
          f"({f(a['cpu_oracle']['one_thread']['ms_per_step'], 1)} ms on one), this repo {f(a['b200']['ms_per_step'])} ms, reference "
          f"kernels {f(a.get('reference_cuda', {}).get('ms_per_step'))} ms (launch-latency bound: ~40 launches and a host "
          f"read-back per step).\n")
    D = cs.get("D", {})
    if D:
        w(f"**D** ({D['workload']}): {D['intersections']} intersections; operator path {f(D['ms_per_step'])} ms/step "
          f"({D['value'] / 1e6:.0f} M Gaussians/s), `e2e_resident` {f(D.get('e2e_resident_ms'))} ms, extended path "
          f"{f(D.get('fused', {}).get('ms_per_step'))} ms, reference kernels {f(D.get('reference_cuda', {}).get('ms_per_step'))} ms "
          f"({f(D.get('reference_cuda', {}).get('speedup_device_resident'), 2)}×).")
        t = D.get("train", {})
        if "b200" in t:
            w(f"Training iteration: {f(t['b200']['ms_per_iter'])} ms vs {f(t.get('reference_cuda', {}).get('ms_per_iter'))} ms "
              f"({f(t.get('speedup'), 2)}×).")
        w("\n| kernel scope (D) | ms |\n|---|---:|")
        for n, v in D.get("kernels_ms", {}).items():
            w(f"| `{n}` | {f(v, 4)} |")
if rest:
    w("\n## Multi-GPU (one view per rank, replicated parameters, gradient exchange)\n")
    w("| GPUs | ms/step | Gaussians/s (all ranks) | efficiency vs N x (1 GPU) | e2e_resident ms | exchange | all-reduce of 236 B/Gaussian |\n|---:|---:|---:|---:|---:|---|---|")
    lines = []
    for pth in rest:
        try:
            lines.append(json.loads([l for l in open(pth).read().splitlines() if l.startswith("{")][-1]))
        except Exception:
            pass
    lines.sort(key=lambda x: x["n_gpus"])
    base = next((x for x in lines if x["n_gpus"] == 1), None)
    for x in lines:
        eff = x["value"] / (x["n_gpus"] * base["value"]) if base else None
        ex = x.get("exchange", {})
        w(f"| {x['n_gpus']} | {f(x['ms_per_step'])} | {x['value'] / 1e6:.1f} M | {f(eff, 3)} | {f(x['e2e_resident']['ms_per_step'])} | "
          f"{ex.get('mode', '—')} | {f(ex.get('allreduce_236B_ms'))} ms, {f(ex.get('allreduce_bus_GBps'), 0)} GB/s bus |")
open(f"profiles/{tag}_bench.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
