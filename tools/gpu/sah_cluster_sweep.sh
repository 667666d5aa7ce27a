# build time and frame time against the builder's cluster size (HL_OPT_SAH_CLUSTER; 0 = plain LBVH topology).
# run on the GPU box:  gpurun -- 'bash tools/gpu/sah_cluster_sweep.sh > gpurun_out/sah_cluster_sweep.log 2>&1'
for c in 0 2 4 8; do
  for n in 100000 1000000 10000000; do HL_SAH_CLUSTER=$c timeout 300 python tools/build_time.py $n 2>&1 | tail -1; done
  for sc in terrain foliage city; do HL_SAH_CLUSTER=$c timeout 300 python tools/frame_time.py $sc 32 2>&1 | tail -1; done
done
