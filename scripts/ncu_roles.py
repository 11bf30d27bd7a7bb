"""Per-role digest of an `ncu --set full --import-source on` capture: warp-stall samples per SASS line, summed per code region.
usage: python scripts/ncu_roles.py <rep> <kernel regex> [launch index]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + rx,
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name"')     # ncu prints every selected launch twice
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
S = lambda r: int(r[ix["# Samples"]] or 0)
tot = sum(S(r) for r in data)
print(rows[0][1][:100], "total samples", tot)
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not" not in k]
# regions = maximal runs of lines with the same order of magnitude of executions are hard to get; print top sites instead
for i, r in sorted(enumerate(data), key=lambda t: -S(t[1]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    top = max(stall_keys, key=lambda k: float(r[ix[k]] or 0))
    print("%5d %-70s %6.2f%% exec=%s %s" % (i, r[ix["Source"]][:70], 100.0 * S(r) / tot, r[ix["Instructions Executed"]], top))
print("-- samples per 50 lines")
for a in range(0, len(data), 50):
    s = sum(S(r) for r in data[a:a + 50])
    if s > tot / 500:
        print(a, "%.1f%%" % (100.0 * s / tot), data[a][ix["Instructions Executed"]])
