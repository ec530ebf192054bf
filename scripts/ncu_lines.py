"""Dev tool: per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = []; h = None; sect = 0; files = []
for r in rows:
    if r and r[0] in ("File Name", "File Path"): files.append(r[1]); continue
    if r and r[0] == "Line No": h = r; sect += 1; continue
    if h and len(r) > 8 and r[0].isdigit():
        i = h.index("Instructions Executed"); s = h.index("# Samples")
        try: n = int(r[i])
        except ValueError: continue
        if n > 0: out.append((n, int(r[s]) if r[s].strip().isdigit() else 0, sect, int(r[0]), r[1].strip()[:120]))
tot = sum(o[0] for o in out); smp = sum(o[1] for o in out)
print("total warp instructions", tot, "samples", smp)
out.sort(reverse=True)
for o in out[:top]:
    print("%10d %5.1f%% smp %5.1f%% [%d]:%-4d %s" % (o[0], 100.0 * o[0] / tot, 100.0 * o[1] / max(smp, 1), o[2], o[3], o[4]))
# per file-section / function-range summary
import collections
agg = collections.Counter()
for n, s, sect, line, src in out:
    key = "sect%d" % sect
    if sect == 6:
        key = ("render: finalize" if 256 <= line <= 323 else "render: emit_post" if 210 <= line <= 243 else "render: texel/blend helpers" if 324 <= line <= 367
               else "render: shade_base_ordered" if 368 <= line <= 397 else "render: rasterise" if 398 <= line <= 477 else "render: other")
    agg[key] += n
for k, v in agg.most_common(): print("%-32s %10d %5.1f%%" % (k, v, 100.0 * v / tot))
print(files)
