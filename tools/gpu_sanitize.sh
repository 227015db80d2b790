# compute-sanitizer over smoke() (memcheck, racecheck) and over the clustering edge-case tests (memcheck)
set -x
mkdir -p gpurun_out
TAG=${1:-san}
timeout 400 compute-sanitizer --tool memcheck --log-file gpurun_out/${TAG}_memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.out 2>&1; tail -3 gpurun_out/${TAG}_memcheck_smoke.log; tail -2 gpurun_out/${TAG}_memcheck_smoke.out
timeout 400 compute-sanitizer --tool racecheck --log-file gpurun_out/${TAG}_racecheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck_smoke.out 2>&1; tail -3 gpurun_out/${TAG}_racecheck_smoke.log; tail -2 gpurun_out/${TAG}_racecheck_smoke.out
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/${TAG}_memcheck_bfs_tests.log python -m pytest tests/test_gpu_parity.py -q -x -k "truncated or one_way or malformed or cap" > gpurun_out/${TAG}_memcheck_bfs_tests.out 2>&1; tail -3 gpurun_out/${TAG}_memcheck_bfs_tests.log; tail -3 gpurun_out/${TAG}_memcheck_bfs_tests.out
