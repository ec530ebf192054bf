# usage: bash scripts/gpu_evidence.sh TAG   -> gpurun_out/TAG_*  (the evidence set copied into profiles/: test log, bench
# lines, reference arm, smoke, ncu launch list of the default bench command, full captures of the dominant kernels of the
# BASELINE configs reduced to text summaries, dram traffic per launch as JSON)
TAG=$1
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
# launch list of the same command (shorter run, headline workload only): per-launch durations, cold-cache and serialised
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_launches_full.csv python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_launches_full.csv")) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
body = rows[hdr[0]:] if hdr else rows
keep = body[:1] + body[-260:]        # the last launches = the steady-state steps of the timed regions
csv.writer(open("gpurun_out/${TAG}_launches_bossfight16384.csv", "w")).writerows(keep)
PY
rm -f gpurun_out/${TAG}_launches_full.csv
# full captures (steady state: the launch-skip reaches past the 300 burn-in steps)
cap() {  # kernel game envs skip extra-bench-args
  k=$1; g=$2; n=$3; skip=$4; shift 4
  R=gpurun_out/${TAG}_${k}_${g}_${n}
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip $skip --launch-count 1 -f -o $R python bench.py --game $g --envs-per-gpu $n --steps 12 --warmup 5 --no-cpu-baseline "$@" > ${R}.log 2>&1
  { python scripts/ncu_raw.py $R.ncu-rep; python scripts/ncu_funcs.py $R.ncu-rep $n; python scripts/ncu_lines.py $R.ncu-rep 40; python scripts/ncu_samples.py $R.ncu-rep 25; } > ${R}_summary.txt 2>&1
  rm -f $R.ncu-rep ${R}.log
}
cap k_render bossfight 16384 320
cap k_step bossfight 16384 320
cap k_render coinrun 4096 320
cap k_step coinrun 4096 320
cap k_render maze 256 320
cap k_render jumper 32768 320
cap k_reset jumper 32768 330 --max-episode-steps 32
python - <<PY
import glob, json, re
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of k_render from the ${TAG} ncu --set full captures (steady state, caches flushed by ncu); bench.py copies it into roofline.traffic when game / envs match"}
for f in sorted(glob.glob("gpurun_out/${TAG}_k_render_*_summary.txt")):
    m = re.search(r"k_render_(\w+?)_(\d+)_summary", f)
    t = open(f).read()
    def val(name):
        mm = re.search(name + r"\s+([\d.]+)\s+(\w+)", t)
        return float(mm.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[mm.group(2)] if mm else None
    r, w = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    if r is not None and w is not None:
        out["%s@%s" % (m.group(1), m.group(2))] = {"kernel": "k_render<%s>" % m.group(1), "dram_bytes_read": int(r), "dram_bytes_write": int(w), "capture": "${TAG}"}
json.dump(out, open("gpurun_out/${TAG}_traffic.json", "w"), indent=1)
print(json.dumps(out)[:600])
PY
tail -1 gpurun_out/${TAG}_bench_default.json | cut -c1-300
tail -1 gpurun_out/${TAG}_reference.json | cut -c1-200
