"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), written to profiles/<tag>_sass_histogram.md.
Evidence for the instruction-level claims of DESIGN.md: packed fp32 (FFMA2 / FMUL2 / FADD2), bulk async copies (UBLKCP),
mbarrier (SYNCS), REDs, shuffles, MUFU, shared-memory atomics (ATOMS), MATCH / REDUX of the sort kernels."""
import collections
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = "gaussian-splatting-cuda_b200/lib/libgsb200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "RED", "ATOMS", "ATOMG",
       "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "MATCH", "REDUX", "VOTE", "BAR", "HMMA"]
out = [f"# SASS opcode histogram {tag}\n",
       f"`cuobjdump -sass {lib}` (sm_100a only), static instruction counts per kernel; produced by "
       "`python profiles/tools/sass_hist.py`.\n",
       "| kernel | total | " + " | ".join(KEY) + " |", "|---|---:|" + "---:|" * len(KEY)]
for k, c in hist.items():
    if not k.startswith("gsb::"):
        continue
    out.append(f"| `{k}` | {sum(c.values())} | " + " | ".join(str(c.get(x, 0)) for x in KEY) + " |")
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
out.append(f"\nArchitectures in the fat binary: {', '.join(arch)}.")
open(f"profiles/{tag}_sass_histogram.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:8]))
