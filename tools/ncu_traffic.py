"""DRAM traffic of the solver kernels of one tick from an `ncu --set full` report -> profiles/ncu_traffic.json, keyed by the hash of the CUDA
sources (sapiogenesis_b200.build.source_hash) and by the workload. bench.py prints roofline.traffic from it only when both match the run.
usage: python tools/ncu_traffic.py <report.ncu-rep> <organisms> <world_age_ticks> <out.json>"""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sapiogenesis_b200.build import source_hash

rep, organisms, age, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
def to_ms(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
kern = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    k = kern.setdefault(name + " grid " + r[ix["launch__grid_size"]], {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
    k["launches"] += 1; k["dram_bytes"] += rd + wr; k["ms"] += to_ms(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
stage = lambda pred: sum(v["dram_bytes"] for n, v in kern.items() if pred(n))
doc = {
    "source_sha": source_hash(), "organisms": organisms, "world_age_ticks": age, "report": os.path.basename(rep),
    "note": "ncu --set full --clock-control none, one tick of the aged world; dram__bytes_read.sum + dram__bytes_write.sum per tick",
    "kernels": {"org_solve": stage(lambda n: n.startswith("k_org_solve")) + stage(lambda n: n.startswith("k_component")),
                "org_solve_classes_only": stage(lambda n: n.startswith("k_org_solve")),
                "components_only": stage(lambda n: n.startswith("k_component")),
                "coupled": stage(lambda n: n.startswith("k_coupled"))},
    "per_kernel": kern,
}
with open(out, "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps(doc["kernels"]))
