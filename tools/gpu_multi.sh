set -x
N=${1:-2}
TAG=${2:-m}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_weak_n$N.json 2> gpurun_out/${TAG}_weak_n$N.err || tail -20 gpurun_out/${TAG}_weak_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_strong_n$N.json 2> gpurun_out/${TAG}_strong_n$N.err || tail -20 gpurun_out/${TAG}_strong_n$N.err
python - <<PY
import json
for f in ("weak", "strong"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s_n$N.json" % f))
        e = d["e2e"]
        print(f, "N", d["n_gpus"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(e["value"], 1), "h2d/rank", e["h2d_GBps_per_rank"], "ceiling", e["h2d_ceiling_gbs"], "frac", e["frac_of_h2d_ceiling"], d["config"]["cpu_binding"])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
