# `--set full` of the kernels added late in round 2 (one step of bench.py --profile-mode); raw page dumped on the box
set -x
TAG=${1:-new}
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'k_voxelize_fp_long|k_voxelize_fp<|k_sec_mean_narrow|k_seg_reduce|k_group_insert_claim|k_max_i32|k_cl_keys|k_radix_onesweep|k_cl_flatten' -s 0 -c 60 -f -o gpurun_out/${TAG}_full python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*; rm -f gpurun_out/${TAG}_full.ncu-rep
