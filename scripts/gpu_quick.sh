# usage: bash scripts/gpu_quick.sh TAG "game:envs ..." ["pytest -k expr"]   — fast iteration round: a parity subset + bench lines
TAG=$1; WORK="$2"; K="${3:-golden or test_live_oracle or prefetch or dropin_batched}"
python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
for w in $WORK; do
  g=${w%%:*}; n=${w##*:}
  steps=100; [ "$n" -ge 16384 ] && steps=40
  python bench.py --game $g --envs-per-gpu $n --steps $steps --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_${g}_${n}.json 2>gpurun_out/${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${g}_${n}.json").read().strip().splitlines()[-1])
    print("$g $n", "%.2fM/s"%(d["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], {k: round(v,4) for k,v in d["kernel_ms_per_step"].items()}, "frac %.3f"%d["roofline"]["frac"], "e2e %.2fM"%(d["e2e"]["value"]/1e6))
except Exception as e:
    print("$g $n FAILED", e); print(open("gpurun_out/${TAG}.err").read()[-2000:])
PY
done
