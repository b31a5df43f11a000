"""Summarise an ncu report's SASS page: executed instructions by opcode and by contiguous equal-count segment.
usage: python profiles/sass_hotspots.py report.ncu-rep [dump.txt]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if "Source" in r)
rows = rows[rows.index(hdr) + 1:]
ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows:
    if len(r) <= ia or not r[ia].isdigit():
        continue
    data.append((int(r[ia]), int(r[ist]), r[isrc].strip(), {h: int(r[i] or 0) for i, h in stall_cols}))
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
byop = collections.Counter(); st = collections.Counter()
for n, s, src, _ in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src
    byop[op] += n; st[op] += s
print(f"instructions executed {tot}  stall samples {ts}")
for op, n in byop.most_common(18):
    print(f"  {op:10s} {n:12d} {100*n/tot:5.1f}%  stalls {100*st[op]/max(ts,1):5.1f}%")
seg = []; cur = None
for i, (n, s, src, sd) in enumerate(data):
    if cur and abs(n - cur[2]) <= 0.02 * max(n, cur[2], 1):
        cur[1] = i; cur[3] += n; cur[4] += s
        for k, v in sd.items(): cur[5][k] += v
    else:
        if cur: seg.append(cur)
        cur = [i, i, n, n, s, collections.Counter(sd)]
seg.append(cur)
print("segments (>1% of instructions or stalls):")
for a, b, n, t, s, sd in seg:
    if t > 0.01 * tot or s > 0.01 * ts:
        top = ", ".join(f"{k[6:]} {v}" for k, v in sd.most_common(3))
        print(f"  [{a:4d}-{b:4d}] len {b-a+1:4d} count {n:9d} instr {100*t/tot:5.1f}%  stalls {100*s/ts:5.1f}%  ({top})")
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(f"{i:5d} {n:10d} {s:6d} {src}" for i, (n, s, src, _) in enumerate(data)))
