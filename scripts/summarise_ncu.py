"""Turn gpurun_out ncu captures into the committed summaries under profiles/ (run here, no GPU needed).

    python scripts/summarise_ncu.py <launches.csv> <full.ncu-rep> <out.md> [<traffic.json>]

With a fourth argument the DRAM bytes of the full-size launches (scripts/ncu_kernels_fullsize.py, one launch per bench
tag, in TAGS order) are also written as JSON; bench.py reads that file to fill `roofline.traffic`.
"""
import collections, csv, io, subprocess, sys

launch_csv, rep, out = sys.argv[1:4]
traffic_json = sys.argv[4] if len(sys.argv) > 4 else None
lines = [l for l in open(launch_csv) if not l.startswith("==")]
agg, total, n = collections.OrderedDict(), 0.0, 0
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v * 1e3 if u in ("second", "s") else v
    k = row["Kernel Name"].split("(")[0][:100]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; total += v; n += 1
md = ["# ncu summary (round 2)", "", "Source: `%s` (launch list, `--metrics gpu__time_duration.sum --clock-control none`) and `%s` (`--set full`)." % (launch_csv, rep),
      "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
      "## Launch list: %d launches, %.1f ms total" % (n, total), "", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    md.append("| `%s` | %d | %.2f | %.1f%% |" % (k, c, ms, 100 * ms / total))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % active"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "LSU wavefront %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
md += ["", "## `--set full` per kernel (one launch each)", "", "| kernel | " + " | ".join(w[1] for w in want) + " |", "|---|" + "---:|" * len(want)]
seen = set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0][:70]
    key = (name, r[idx["launch__grid_size"]])
    if key in seen and traffic_json is None:
        continue
    seen.add(key)
    vals = []
    for m, _ in want:
        v = r[idx[m]] if m in idx else ""
        try:
            v = "%.4g" % float(v.replace(",", ""))
        except ValueError:
            pass
        vals.append("%s %s" % (v, units[idx[m]] if m in idx and units[idx[m]] not in ("", "%") else ""))
    md.append("| `%s` | " % name + " | ".join(vals) + " |")
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md[-14:]))

if traffic_json:
    import json
    sys.path.insert(0, "scripts")
    from ncu_kernels_fullsize import TAGS
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    body = rows[2:]
    assert len(body) == len(TAGS), (len(body), len(TAGS))
    outj = {"source": rep, "note": "one full-size (config 5) launch per tag, ncu --set full --clock-control none; bytes per launch",
            "kernels": {}}
    for tag, r in zip(TAGS, body):
        rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
        outj["kernels"][tag] = {"kernel": r[idx["Kernel Name"]].split("(")[0], "dram_read_bytes": rd, "dram_write_bytes": wr,
                                "traffic_bytes": rd + wr, "ncu_time_ms": float(r[idx["gpu__time_duration.sum"]])}
    json.dump(outj, open(traffic_json, "w"), indent=1)
