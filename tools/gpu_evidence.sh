#!/bin/bash
# One GPU-box visit: tests, bench (both arms), launch list, full ncu captures of the dominant kernels.
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by hand afterwards.
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:encode_kernel -s 5 -c 1 -o gpurun_out/${TAG}_encode_full -f python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_ncu_encode.log 2>&1; echo "ncu encode rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/${TAG}_decode_full -f python tools/prof_encode.py > gpurun_out/${TAG}_ncu_decode.log 2>&1; echo "ncu decode rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gdn_tc_fwd2_kernel -s 1 -c 1 -o gpurun_out/${TAG}_gdn_fwd_full -f python tools/prof_gdn.py > gpurun_out/${TAG}_ncu_gdn_fwd.log 2>&1; echo "ncu gdn fwd rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gdn_tc_fwd4_kernel -s 1 -c 1 -o gpurun_out/${TAG}_gdn_fwd192_full -f python tools/prof_gdn.py > gpurun_out/${TAG}_ncu_gdn_fwd192.log 2>&1; echo "ncu gdn fwd C=192 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gdn_tc_bwd3_kernel -s 1 -c 1 -o gpurun_out/${TAG}_gdn_bwd128_full -f python tools/prof_gdn_bwd128.py > gpurun_out/${TAG}_ncu_gdn_bwd128.log 2>&1; echo "ncu gdn bwd C=128 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gdn_tc_bwd_dx2_kernel -s 1 -c 1 -o gpurun_out/${TAG}_gdn_bwd192_dx2_full -f python tools/prof_gdn_bwd192.py > gpurun_out/${TAG}_ncu_gdn_bwd192_dx2.log 2>&1; echo "ncu gdn bwd dx C=192 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gdn_tc_bwd_dgamma2_kernel -s 1 -c 1 -o gpurun_out/${TAG}_gdn_bwd192_dgamma2_full -f python tools/prof_gdn_bwd192.py > gpurun_out/${TAG}_ncu_gdn_bwd192_dgamma2.log 2>&1; echo "ncu gdn bwd dgamma C=192 rc=$?"
ls -la gpurun_out | tail -20
