"""Turn the ncu artefacts in gpurun_out/ (tag given as argv[1]) into markdown summaries under profiles/."""
import csv, collections, re, subprocess, sys, io, os
tag = sys.argv[1]
out = []
def launches(path=None, title=None):
    path = path or f'gpurun_out/launches_{tag}.csv'
    if not os.path.exists(path): return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv: continue
        full = r[ik]
        if 'gsb::' in full:
            name = re.search(r'gsb::\w+(<[\w, ]+>)?', full).group(0)
        elif 'fast_gs::' in full:
            name = 'fast_gs::' + re.search(r'(\w+)\(', full).group(1)
        elif 'cub::' in full:
            name = 'cub::' + re.search(r'cub::(?:\w+::)*(\w+)', full).group(1)
        else:
            name = re.sub(r'<.*', '', re.sub(r'^void ', '', full))[:70]
        agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += float(r[iv].replace(',', ''))
    tot = sum(v for _, v in agg.values())
    out.append(f"## {title or 'Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, `bench.py --steps 2 --warmup 1`)'}\n")
    out.append(f"{sum(c for c, _ in agg.values())} launches captured, {tot/1e6:.2f} ms of device time (cold-cache, serialised: compare SHARES).\n")
    out.append("| kernel | launches | total µs | share |\n|---|---:|---:|---:|")
    mine = 0.0
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        out.append(f"| `{k}` | {c} | {v/1e3:.1f} | {100*v/tot:.1f} % |")
    mine = sum(v for k, (c, v) in agg.items() if k.startswith('gsb::'))
    cub = sum(v for k, (c, v) in agg.items() if 'cub::' in k)
    out.append(f"\nOwn kernels (`gsb::`): {100*mine/tot:.1f} % of device time; CUB (scan / radix sort): {100*cub/tot:.1f} %; "
               f"torch glue of the L3 caller (elementwise, reductions, `inverse`): {100*(tot-mine-cub)/tot:.1f} %.\n")
WANT = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'), ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput % of peak'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
        ('smsp__inst_executed.sum', 'warp instructions'), ('smsp__inst_executed_op_global_red.sum', 'global RED instructions'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads / instruction'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
        ('launch__registers_per_thread', 'registers / thread'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]
def full(kernel, prefix='', note=''):
    p = f'gpurun_out/prof_{prefix}{kernel}_{tag}.ncu-rep'
    if not os.path.exists(p): return
    txt = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3: return
    hdr, units, vals = rows[0], rows[1], rows[-1]
    name = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
    out.append(f"## `{kernel}`{note} (`ncu --set full --clock-control none --import-source on`, 1 launch)\n")
    if name is not None: out.append(f"Captured instance: `{vals[name][:110]}`\n")
    out.append("| metric | value |\n|---|---:|")
    for m, label in WANT:
        if m in hdr:
            i = hdr.index(m); out.append(f"| {label} (`{m}`) | {vals[i]} {units[i]} |")
    st = [(h.split('issue_stalled_')[1].split('_per')[0], float(vals[i])) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    st.sort(key=lambda t: -t[1])
    out.append("\nTop warp-stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for n, v in st[:7]) + "\n")
launches()
for k in ('raster_bwd_kernel', 'raster_fwd_kernel', 'tile_bin_kernel', 'ssim_l1_kernel', 'fused_front_kernel',
          'fused_back_kernel'):
    full(k)
# fastgs path: profiles/tools/fastgs_prof.py
launches(f'gpurun_out/launches_fastgs_{tag}.csv', 'fastgs path, this backend: launch list of `profiles/tools/fastgs_prof.py --iters 2`')
launches(f'gpurun_out/launches_fastgs_ref_{tag}.csv', "fastgs path, the reference's own kernels: launch list of `fastgs_prof.py --iters 2 --reference`")
for k in ('fgs_front_kernel', 'raster_fwd_kernel', 'raster_bwd_kernel', 'fgs_back_kernel', 'tile_bin_kernel'):
    full(k, 'fastgs_', ' in the fastgs path')
open(f'profiles/{tag}_ncu_summary.md', 'w').write(f"# ncu summary {tag}\n\nSource artefacts: `gpurun_out/launches_{tag}.csv`, `gpurun_out/prof_*_{tag}.ncu-rep` "
     f"(scratch, not tracked); this file is the tracked digest.\n\n" + "\n".join(out) + "\n")
print("\n".join(out))
