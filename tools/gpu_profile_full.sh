# launch list + `--set full` of the library's main kernels for one step; raw page dumped on the box
set -x
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/${TAG}_launches.log 2>&1
N=$(grep -c '"' gpurun_out/${TAG}_launches.csv)
HALF=$(( (N - 1) / 2 ))
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_bq_|k_cl_verify|k_cl_flatten|k_cl_label|k_cl_sample|k_voxelize_fp|k_vox_fill|k_vox_rank|k_sec_mean|k_seg_reduce|k_gather_rows|k_iou_count|k_radix_onesweep|k_radix_hist_all|k_scan_fused|k_group_insert|k_group_assign' -s 0 -c 400 -f -o gpurun_out/${TAG}_full python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*; rm -f gpurun_out/${TAG}_full.ncu-rep
