for v in ""; do
  if [ -n "$v" ]; then export PG_B200_LIB=$PWD/d3net_b200/variants/libpg_$v.so; else unset PG_B200_LIB; fi
  python bench.py --no-cpu-baseline --steps 10 > gpurun_out/ab_$v.json 2>/dev/null
  python - <<P
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("variant [$v]", round(d["value"],1), round(d["ms_per_step"],3), {k:v["ms_per_step"] for k,v in list(d["per_kernel"].items())[:6]})
P
done
