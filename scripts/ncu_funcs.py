"""Dev tool: warp instructions and stall samples of one kernel aggregated by ENCLOSING FUNCTION of each source line
(function starts found by a regex over the source files named in the report).
usage: python scripts/ncu_funcs.py report.ncu-rep [frames]     (frames: divide by it to print per-frame numbers)"""
import collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]; frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUNC = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:PG2_DEV\w*|__global__|__device__|inline)\b[^;{]*?\b(\w+)\s*\(")
def starts(path):
    out = []
    try:
        src = open(path).read().splitlines()
    except OSError:
        try: src = open(os.path.join(ROOT, "procgen2_b200", "csrc", os.path.basename(path))).read().splitlines()
        except OSError:
            try: src = open(os.path.join(ROOT, "procgen2_b200", "csrc", "games", os.path.basename(path))).read().splitlines()
            except OSError: return out
    for i, l in enumerate(src, 1):
        m = FUNC.match(l)
        if m and not l.strip().startswith("//"): out.append((i, m.group(1)))
    return out
agg = collections.Counter(); smp = collections.Counter(); h = None; cur = None; st = []
for r in rows:
    if r and r[0] in ("File Name", "File Path"): cur = r[1]; st = starts(cur); continue
    if r and r[0] == "Line No": h = r; continue
    if h and len(r) > 8 and r[0].isdigit():
        i = h.index("Instructions Executed"); s = h.index("# Samples")
        try: n = int(r[i])
        except ValueError: continue
        line = int(r[0]); name = "?"
        for ln, fn in st:
            if ln <= line: name = fn
            else: break
        key = "%s:%s" % (os.path.basename(cur or "?"), name)
        agg[key] += n; smp[key] += int(r[s]) if r[s].strip().isdigit() else 0
tot = sum(agg.values()); stot = max(sum(smp.values()), 1)
print("total warp instructions", tot, ("= %.0f per frame" % (tot / frames)) if frames else "", "samples", stot)
for k, v in agg.most_common(40):
    print("%-52s %11d %5.1f%%  smp %5.1f%%%s" % (k, v, 100.0 * v / tot, 100.0 * smp[k] / stot, ("  %7.0f/frame" % (v / frames)) if frames else ""))
