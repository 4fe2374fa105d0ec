"""Print the hottest SASS instructions (warp-stall samples) of one kernel from an .ncu-rep source page."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_")]
data = []
for i, r in enumerate(rows[h + 1:]):
    if len(r) <= iex: continue
    try: data.append((int(r[isamp] or 0), r[ia].strip(), int(r[iex] or 0), i, r))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1
print(rows[0][:2], "total samples", tot, "instructions", len(data))
for s, src, ex, i, r in sorted(data, key=lambda d: -d[0])[:top]:
    reasons = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols if (r[c] or "0") != "0"), reverse=True)[:2]
    print(f"{s:6d} {100*s/tot:5.1f}% ex={ex:7d} #{i:5d} {src[:70]:70s} {reasons}")
