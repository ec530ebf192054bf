# usage: bash scripts/gpu_evidence.sh TAG   -> gpurun_out/TAG_*  (bench lines, ncu launch list, full captures as CSV)
TAG=$1
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err
# launch list of the same command (shorter run): per-launch durations, cold-cache and serialised
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_coinrun4096.csv python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
for k in k_render k_step k_reset; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 15 --launch-count 1 -f -o gpurun_out/${TAG}_$k python bench.py --steps 12 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$k.log 2>&1
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_${k}_coinrun4096.csv 2>/dev/null
  python scripts/ncu_raw.py gpurun_out/${TAG}_$k.ncu-rep > gpurun_out/${TAG}_${k}_summary.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/${TAG}_$k.ncu-rep 60 >> gpurun_out/${TAG}_${k}_summary.txt 2>&1
  python scripts/ncu_samples.py gpurun_out/${TAG}_$k.ncu-rep 30 >> gpurun_out/${TAG}_${k}_summary.txt 2>&1
  rm -f gpurun_out/${TAG}_$k.ncu-rep
done
tail -2 gpurun_out/${TAG}_bench_default.json | cut -c1-400
tail -1 gpurun_out/${TAG}_reference.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
