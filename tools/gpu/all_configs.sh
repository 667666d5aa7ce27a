# frame times of the BASELINE configs with the library defaults: bash tools/gpu/all_configs.sh > gpurun_out/configs_rXX.jsonl
for s in cornell terrain foliage city; do python tools/frame_time.py $s; done
