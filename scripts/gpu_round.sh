# usage: bash scripts/gpu_round.sh TAG "game:envs ..." [ncu games]
TAG=$1; shift
WORK="$1"; shift
NCU="$1"
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
for w in $WORK; do
  g=${w%%:*}; n=${w##*:}
  steps=100; [ "$n" -ge 16384 ] && steps=40
  python bench.py --game $g --envs-per-gpu $n --steps $steps --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_${g}_${n}.json 2>gpurun_out/${TAG}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_${g}_${n}.json").read().strip().splitlines()[-1])
print("$g $n", "%.2fM/s"%(d["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], d["kernel_ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2fM"%(d["e2e"]["value"]/1e6))
PY
done
for g in $NCU; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_render --launch-skip 12 --launch-count 1 -f -o gpurun_out/${TAG}_render_$g python bench.py --game $g --envs-per-gpu 4096 --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$g.log 2>&1
done
