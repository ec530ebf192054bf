"""Dev tool: per-source-line stall samples. usage: python scripts/ncu_samples.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = []; h = None; sect = 0
for r in rows:
    if r and r[0] == "Line No": h = r; sect += 1; continue
    if h and len(r) > 8 and r[0].isdigit():
        s = h.index("# Samples"); i = h.index("Instructions Executed")
        stalls = {}
        for j, n in enumerate(h):
            if n.startswith("stall_") and "Not Issued" not in n:
                try: x = int(r[j] or 0)
                except ValueError: x = 0
                if x: stalls[n[6:]] = x
        try: n = int(r[s] or 0)
        except ValueError: continue
        if n > 0: out.append((n, int(r[i] or 0), sect, int(r[0]), r[1].strip()[:90], stalls))
tot = sum(o[0] for o in out)
print("samples", tot)
out.sort(key=lambda o: -o[0])
for o in out[:top]:
    st = sorted(o[5].items(), key=lambda kv: -kv[1])[:3]
    print("%6d %5.1f%% inst %9d [%d]:%-4d %-90s %s" % (o[0], 100.0 * o[0] / tot, o[1], o[2], o[3], o[4], st))
