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
fgs = d.get("fastgs", {})
if "b200" in fgs:
    w("\n## fastgs (EWA) path, SURVEY.md 8 f4: forward + backward through the reference caller's sequence, device-resident\n")
    w("| config | this repo ms/step | reference fastgs kernels ms/step | ratio | instances |\n|---|---:|---:|---:|---:|")
    w(f"| B | {f(fgs['b200']['ms_per_step'])} | {f(fgs.get('reference_cuda', {}).get('ms_per_step'))} | {f(fgs.get('speedup'), 2)}× | {fgs['b200'].get('instances')} |")
    fd = cs.get("D", {}).get("fastgs", {}) if cs else {}
    if "b200" in fd:
        w(f"| D | {f(fd['b200']['ms_per_step'])} | {f(fd.get('reference_cuda', {}).get('ms_per_step'))} | {f(fd.get('speedup'), 2)}× | {fd['b200'].get('instances')} |")
    w("\n| kernel scope (fastgs) | B ms | D ms |\n|---|---:|---:|")
    kd = fd.get("b200", {}).get("kernels_ms", {})
    for n, v in fgs["b200"].get("kernels_ms", {}).items():
        w(f"| `{n}` | {f(v, 4)} | {f(kd.get(n), 4)} |")
    w("\nThe reference's kernels are its own sources compiled unmodified as compute_90 PTX (oracle/build_ref.py explains why).")
    tf, tfd = d.get("train_fastgs", {}), (cs.get("D", {}).get("train_fastgs", {}) if cs else {})
    if "b200" in tf:
        w("\n### Training iteration on the fastgs path (render + background + SSIM/L1 loss + backward + Adam; camera + target H2D, loss D2H)\n")
        w("| config | this repo, API-exact forward (one read-back) | no read-back (capacity) | one CUDA graph | reference kernels | ratio eager / graph |\n|---|---:|---:|---:|---:|---:|")
        for nm, t in (("B", tf), ("D", tfd)):
            if "b200" in t:
                rc2 = t.get("reference_cuda", {})
                w(f"| {nm} | {f(t['b200']['ms_per_iter'])} ms ({f(t['b200']['iters_per_sec'], 1)} it/s) | "
                  f"{f(t.get('b200_capacity', {}).get('ms_per_iter'))} ms | {f(t.get('b200_cuda_graph', {}).get('ms_per_iter'))} ms "
                  f"({f(t.get('b200_cuda_graph', {}).get('iters_per_sec'), 1)} it/s) | {f(rc2.get('ms_per_iter'))} ms "
                  f"({f(rc2.get('iters_per_sec'), 1)} it/s) | {f(t.get('speedup'), 2)}× / {f(t.get('speedup_cuda_graph'), 2)}× |")

# ---- the results table, re-derived from the same line -------------------------------------------------
import os
import re
bl = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "BASELINE.md")
if os.path.exists(bl) and "ops_ms" in d and cs:
    ob, orf = ops.get("b200", {}), ops.get("reference_cuda", {})
    fwd_names = ("projection_ut_3dgs_fused", "spherical_harmonics_fwd", "intersect_tile", "intersect_offset",
                 "rasterize_to_pixels_from_world_3dgs_fwd")
    bwd_names = ("rasterize_to_pixels_from_world_3dgs_bwd", "spherical_harmonics_bwd")
    sm = lambda o, ns: sum(o.get(n, 0.0) for n in ns) if o else None
    a, D = cs.get("A", {}), cs.get("D", {})
    rows = ["| Config | Impl | GPUs | N | I | fwd ms | bwd ms | Gaussians/s (fwd+bwd) | iters/s | HBM GB/s (algorithmic) | % HBM roofline | binding pipe & % |",
            "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    if a:
        rows.append(f"| A | CPU oracle ({a['cpu_oracle']['cores']} cores; 1 core: {a['cpu_oracle']['one_thread']['ms_per_step']:.0f} ms) | 0 | 10 k | "
                    f"{a['b200'].get('intersections')} | — | — ({a['cpu_oracle']['ms_per_step']:.1f} ms fwd+bwd) | {a['cpu_oracle']['value']:.3g} | — | — | — | — |")
        rows.append(f"| A | reference gsplat sm_100 / new sm_100a | 1 | 10 k | {a['b200'].get('intersections')} | — | — "
                    f"({f(a.get('reference_cuda', {}).get('ms_per_step'))} / {f(a['b200']['ms_per_step'])} ms fwd+bwd, launch-bound) | "
                    f"{a.get('reference_cuda', {}).get('value', 0):.3g} / {a['b200']['value']:.3g} | — | — | — | — |")
    rows.append(f"| B | reference gsplat sm_100 | 1 | 1 M | {cfg['intersections']} | {f(sm(orf, fwd_names))} | {f(sm(orf, bwd_names))} | "
                f"{r.get('value', 0):.3g} ({f(r.get('ms_per_step'))} ms) | {f(tr.get('reference_cuda', {}).get('iters_per_sec'), 1)} | — | — | (its blend backward: 5.7 ms of atomics + per-pair VJP) |")
    rows.append(f"| B | new sm_100a | 1 | 1 M | {cfg['intersections']} | {f(sm(ob, fwd_names))} | {f(sm(ob, bwd_names))} | "
                f"{d['value']:.3g} ({f(d['ms_per_step'])} ms) | {f(tr.get('b200', {}).get('iters_per_sec'), 1)} (graph: "
                f"{f(tr.get('b200_cuda_graph', {}).get('iters_per_sec'), 1)}) | {ro['achieved']:.0f} (`{ro['kernel']}`) | {100 * ro['frac']:.1f} % | "
                f"instruction issue 79 % (blend backward), 74 % + XU 38 % (blend forward): profiles/ |")
    if "b200" in fgs:
        rows.append(f"| B (fastgs / EWA path) | reference fastgs / new sm_100a | 1 | 1 M | {fgs['b200'].get('instances')} instances | — | — | "
                    f"{fgs.get('reference_cuda', {}).get('value', 0):.3g} / {fgs['b200']['value']:.3g} "
                    f"({f(fgs.get('reference_cuda', {}).get('ms_per_step'))} / {f(fgs['b200']['ms_per_step'])} ms) | — | — | — | instruction issue 84 % (EWA blend backward) |")
    rows.append("| C | reference / new | 1 | — | — | — | — | — | not run: needs the reference's trainer, dataset loader and densification strategies (out of scope, SURVEY.md §8); the training-iteration rows of B and D are the per-iteration figures | — | — | — |")
    if D:
        dt = D.get("train", {})
        rows.append(f"| D | reference / new | 1 | 6 M | {D['intersections']} | — | — | {D.get('reference_cuda', {}).get('value', 0):.3g} / {D['value']:.3g} "
                    f"({f(D.get('reference_cuda', {}).get('ms_per_step'))} / {f(D['ms_per_step'])} ms) | "
                    f"{f(dt.get('reference_cuda', {}).get('iters_per_sec'), 1)} / {f(dt.get('b200', {}).get('iters_per_sec'), 1)} (render + loss + Adam; MCMC noise / relocation kernels exist, not in the loop) | — | — | — |")
    rows.append("| E | new | 2/4/8 | 1 M | — | — | — | see `profiles/r2_bench.md` \"Multi-GPU\" / the driver's SCALE record | — | — | — | — |")
    txt = open(bl).read()
    head = "### 2.4 Results table"
    i = txt.index(head)
    j = txt.index("\n", i)
    note = (f"\n\nFilled by `profiles/tools/bench_report.py` from `{main_path}` (B200, SM {d['clocks']['sm_mhz']:.0f} MHz, no throttle "
            f"reasons); fwd / bwd = sums of the per-operator CUDA-event times of the L3 sequence, the Gaussians/s column is the "
            f"whole step including the torch glue.\n\n")
    open(bl, "w").write(txt[:j] + note + "\n".join(rows) + "\n")

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
