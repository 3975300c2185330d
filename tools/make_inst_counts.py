#!/usr/bin/env python
"""gpurun_out/counts.csv (tools/ncu_counts.sh) + count_manifest.json -> profiles/inst_counts.json
(warp instructions per draw, per pipe, per sampler and shape class: what bench.py's `bounds()` reads)
and profiles/traffic.json (DRAM bytes per launch)."""
import csv, io, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "counts.csv")
man = json.load(open(os.path.join(os.path.dirname(src), "count_manifest.json")))
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
by_id = {}
for r in rows:
    by_id.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"]})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
ids = sorted(by_id)
assert len(ids) == sum(m["launches"] for m in man), (len(ids), man)
counts_path = os.path.join(ROOT, "profiles", "inst_counts.json")
traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
counts = json.load(open(counts_path)) if os.path.exists(counts_path) else {}
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
traffic = {k: v for k, v in traffic.items() if isinstance(v, dict)}
pos = 0
for m in man:
    k = by_id[ids[pos + m["launches"] - 1]]          # the last (warm) launch of the row
    pos += m["launches"]
    draws = m["streams"] * m["draws"]
    def val(name, scale=1.0):
        v, unit = k[name]
        return v * scale
    def byt(name):
        v, unit = k[name]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    inst = val("smsp__inst_executed.sum")
    entry = {"inst_per_draw": inst / draws,
             "fp64_per_draw": val("sm__inst_executed_pipe_fp64.sum") / draws,
             "alu_per_draw": val("sm__inst_executed_pipe_alu.sum") / draws,
             "fma_per_draw": val("sm__inst_executed_pipe_fma.sum") / draws,
             "xu_per_draw": val("sm__inst_executed_pipe_xu.sum") / draws,
             "lsu_per_draw": val("sm__inst_executed_pipe_lsu.sum") / draws,
             "lanes_active": val("smsp__thread_inst_executed.sum") / inst,
             "kernel": k["kernel"][:80], "shape": "%dx%d" % (m["streams"], m["draws"]),
             "per_stream_params": m["per_stream"],
             "note": "warp-level instructions per draw (one draw = one lane's output element): x 32 for thread instructions"}
    key = "%s@%d" % (m["dist"], m["draws"]) + ("/per_stream" if m["per_stream"] else "")
    counts[key] = entry
    traffic["%s_f64@%dx%d" % (m["dist"], m["streams"], m["draws"]) + ("/per_stream" if m["per_stream"] else "")] = {
        "bytes": byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum"),
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (tools/ncu_counts.sh), one launch"}
json.dump(counts, open(counts_path, "w"), indent=1, sort_keys=True)
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
print("wrote", counts_path, traffic_path, len(man), "rows")
