"""Write profiles/ncu_traffic.json from the `ncu --set full` captures of one GPU session (tag = argv[1]): DRAM bytes
read + written per launch of the two blend kernels at config B, keyed to a hash of the kernel sources so that bench.py
reports `roofline.traffic` only while the capture still describes the shipped kernels (stale -> null)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

tag = sys.argv[1]
SOURCES = ["gsb_raster.cu", "gsb_raster.cuh", "gsb_common.cuh", "gsb_camera.cuh"]
out = {}
for scope, kernel in (("raster_bwd", "raster_bwd_kernel"), ("raster_fwd", "raster_fwd_kernel")):
    p = os.path.join(ROOT, "gpurun_out", f"prof_{kernel}_{tag}.ncu-rep")
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[-1]

    def get(metric):
        i = hdr.index(metric)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        return int(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u])
    out[scope] = {"gaussians": 1000000, "width": 1920, "height": 1080, "dram_bytes_read": get("dram__bytes_read.sum"),
                  "dram_bytes_write": get("dram__bytes_write.sum"),
                  "capture": f"ncu --set full --clock-control none, one launch, gpurun_out/prof_{kernel}_{tag}.ncu-rep "
                             f"(profiles/{tag}_ncu_summary.md)",
                  "sources": SOURCES, "source_sha": bench.source_hash(SOURCES)}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
