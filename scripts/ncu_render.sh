set -x
for g in coinrun chaser maze bossfight; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_render --launch-skip 12 --launch-count 1 -f -o gpurun_out/r1v_render_$g python bench.py --game $g --envs-per-gpu 4096 --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/r1v_$g.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
