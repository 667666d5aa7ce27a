# ncu --set full capture of one kernel of the bench command: bash tools/gpu/profile_kernel.sh k_shade r01c [skip] [count]
K=${1:-k_shade}; TAG=${2:-rXX}; SKIP=${3:-8}; CNT=${4:-2}
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -o gpurun_out/${TAG}_${K} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${K}.log 2>&1
ncu -i gpurun_out/${TAG}_${K}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${K}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_${K}.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/${TAG}_${K}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_${K}*
