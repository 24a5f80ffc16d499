"""Collect the parity numbers the GPU tests print (`pytest -m gpu -s`, gpurun_out/<tag>_gputest.log) into
profiles/<round>_parity.md: one line per (comparison, tensor) with relative L2, per-Gaussian outliers and max error."""
import re
import sys

tag = sys.argv[1]
out_name = sys.argv[2] if len(sys.argv) > 2 else "r2_parity.md"
log = open(f"gpurun_out/{tag}_gputest.log").read()
grad = re.compile(r"\[(?P<tag>[^\]]+)\] (?P<name>[\w]+): rel_l2 (?P<rel>[\d.e+-]+); Gaussians off by > (?P<tol>[\d.e+-]+): (?P<n>\d+) "
                  r"\((?P<frac>[\d.e+-]+) of (?P<N>\d+)\); max \|err\| / max \|ref\| (?P<mx>[\d.e+-]+)")
rows = {}
for m in grad.finditer(log):
    rows.setdefault(m["tag"], []).append(m)
img = re.findall(r"^\.?\[([^\]]+)\] (image rel_l2 [^\n]+)$", log, re.M)
other = re.findall(r"^\.?F?(\[(?:rs \d|config [ABD][^\]]*)\][^\n]*)$", log, re.M)
summary = re.findall(r"^(\d+ passed[^\n]*|\d+ failed[^\n]*)$", log, re.M)
L = [f"# Parity numbers, GPU session `{tag}`\n",
     f"Extracted from `gpurun_out/{tag}_gputest.log` (`pytest tests -m gpu -s` on the B200) by `profiles/tools/parity_report.py`.",
     f"Result: {summary[-1] if summary else 'n/a'}.\n",
     "How to read: `rel L2` = ‖got − want‖ / ‖want‖ over the tensor (bar 1e-3 for gradients, 1e-4 for images); `off by > 1 %` ="
     " Gaussians whose own relative error exceeds 1e-2 (bar: fraction < 2e-3); `max` = largest absolute error / largest reference"
     " element (bar 5e-2).  See `tests/parity.py`.\n",
     "## Gradients\n", "| comparison | tensor | rel L2 | off by > 1 % | of | max |", "|---|---|---:|---:|---:|---:|"]
for t, ms in rows.items():
    for m in ms:
        L.append(f"| {t} | `{m['name']}` | {m['rel']} | {m['n']} | {m['N']} | {m['mx']} |")
L += ["\n## Images\n"] + [f"* {t}: {v}" for t, v in img]
L += ["\n## Other printed comparisons\n"] + [f"* {o[:400]}" for o in other if "image rel_l2" not in o and ": rel_l2" not in o]
open(f"profiles/{out_name}", "w").write("\n".join(L) + "\n")
print(len(rows), "comparisons,", sum(len(v) for v in rows.values()), "gradient rows")
