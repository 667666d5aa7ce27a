# ncu captures of the bench command (numbers printed under ncu are never bench values):
#   1. launch list with per-launch durations (share of the step per kernel)
#   2. --set full capture of the dominant kernel, k_extend: bounces 0..2 of one frame
# run on the GPU box:  gpurun -- 'bash tools/gpu/profile.sh r02'   (tag names the output files in gpurun_out/)
TAG=${1:-rXX}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 8 -c 3 -o gpurun_out/${TAG}_extend -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_extend.log 2>&1
ncu -i gpurun_out/${TAG}_extend.ncu-rep --page raw --csv > gpurun_out/${TAG}_extend_raw.csv 2>/dev/null
ls -la gpurun_out/
