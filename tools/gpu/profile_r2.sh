# round-2 ncu captures (numbers printed under ncu are never bench values).  bash tools/gpu/profile_r2.sh r02b
#   1. launch list of the bench command (per-launch durations: kernel shares of the step, builder timeline)
#   2. --set full of one whole frame of the bench (all of k_extend / k_shade / k_connect / k_tail: 28 launches)
#   3. --set full of the builder kernels on the 1M-triangle soup
#   4. --set full of the foliage scene's first-bounce k_extend / k_connect
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain-s 0 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade|k_connect|k_tail' -s 84 -c 28 -o gpurun_out/${TAG}_frame -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain-s 0 > gpurun_out/${TAG}_frame.log 2>&1
ncu -i gpurun_out/${TAG}_frame.ncu-rep --page raw --csv > gpurun_out/${TAG}_frame_raw.csv 2>/dev/null
# source pages (SASS + per-instruction counters) of the launches that matter: k_extend bounce 0 and 1, k_shade 0 and 1, k_connect 0, k_tail (first)
for k in 0 1 2 3 4 12; do ncu -i gpurun_out/${TAG}_frame.ncu-rep --page source --csv --launch-skip $k --launch-count 1 2>/dev/null | cut -d, -f1-12 > gpurun_out/${TAG}_frame_source_$k.csv; done
rm -f gpurun_out/${TAG}_frame.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'k_fit_fine|k_coarse_boxes|k_top_build|k_treelets|k_top_refit_treelets|k_collapse|k_radix_tree|k_write_leaves|k_tri_boxes' -c 10 -o gpurun_out/${TAG}_builder -f python tools/sweep_build.py --sizes 1000000 --no-oracle-above 0 > gpurun_out/${TAG}_builder.log 2>&1
ncu -i gpurun_out/${TAG}_builder.ncu-rep --page raw --csv > gpurun_out/${TAG}_builder_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_builder.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_builder_launches.csv python tools/sweep_build.py --sizes 1000000 --no-oracle-above 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_connect' -s 64 -c 4 -o gpurun_out/${TAG}_foliage -f python tools/frame_time.py foliage 4 > gpurun_out/${TAG}_foliage.log 2>&1
ncu -i gpurun_out/${TAG}_foliage.ncu-rep --page raw --csv > gpurun_out/${TAG}_foliage_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_foliage.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 2>/dev/null | cut -d, -f1-12 > gpurun_out/${TAG}_foliage_source_0.csv
rm -f gpurun_out/${TAG}_foliage.ncu-rep
ls -la gpurun_out/ | grep ${TAG}
