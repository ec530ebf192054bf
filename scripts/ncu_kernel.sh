# usage: bash scripts/ncu_kernel.sh TAG kernel_regex "game:envs:skip ..."   (text summaries only; the .ncu-rep is deleted)
TAG=$1; K=$2; WORK="$3"
for w in $WORK; do
  IFS=: read g n skip <<< "$w"
  R=gpurun_out/${TAG}_${K}_${g}_${n}
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip $skip --launch-count 1 -f -o $R python bench.py --game $g --envs-per-gpu $n --steps 12 --warmup 10 --no-cpu-baseline > ${R}.log 2>&1
  python scripts/ncu_raw.py $R.ncu-rep > ${R}_raw.txt 2>&1
  python scripts/ncu_lines.py $R.ncu-rep 70 > ${R}_lines.txt 2>&1
  python scripts/ncu_samples.py $R.ncu-rep 40 > ${R}_samples.txt 2>&1
  python scripts/ncu_funcs.py $R.ncu-rep $n > ${R}_funcs.txt 2>&1
  [ -z "$KEEP_REP" ] && rm -f $R.ncu-rep
done
