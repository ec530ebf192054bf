# usage: bash scripts/gpu_ab.sh TAG "lib:ctas_per_sm ..." "game:envs ..."
TAG=$1; VARS="$2"; WORK="$3"
export PG2_ASSETS=$PWD/procgen2_b200/data/assets.bin
for v in $VARS; do
  lib=${v%%:*}; per=${v##*:}
  for w in $WORK; do
    g=${w%%:*}; n=${w##*:}
    steps=100; [ "$n" -ge 16384 ] && steps=40
    if [ "$lib" = "default" ]; then unset PG2_ENGINE_LIB; else export PG2_ENGINE_LIB=$PWD/exp_libs/$lib; fi
    PG2_RENDER_CTAS_PER_SM=$per python bench.py --game $g --envs-per-gpu $n --steps $steps --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_tmp.json 2>gpurun_out/${TAG}.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().splitlines()[-1])
print("$lib per_sm=$per $g $n", "%.2fM/s"%(d["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], {k:round(v,4) for k,v in d["kernel_ms_per_step"].items()})
PY
  done
done
