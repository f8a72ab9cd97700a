#!/bin/bash
# One GPU call that produces the round's committed evidence under gpurun_out/ (copied into profiles/ afterwards):
#   tests, bench line, step breakdown, GEMM power trace, ncu launch list of the bench command, per-op DRAM traffic,
#   one --set full capture of the GEMM and of the fused lookup->conv kernel.
set -x
R=${1:-r02}
python -m pytest tests -m gpu -q > gpurun_out/${R}_gputests.log 2>&1; tail -3 gpurun_out/${R}_gputests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python tools/step_breakdown.py 20 > gpurun_out/${R}_step_breakdown.json 2>/dev/null
python tools/gemm_power.py 5 gpurun_out/${R}_gemm_power.csv build > gpurun_out/${R}_gemm_power.json 2>/dev/null
python tools/gemm_power.py 3 gpurun_out/${R}_lookup_power.csv lookup > gpurun_out/${R}_lookup_power.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/${R}_ncu_bench.log 2>&1
for op in lookup_conv lookup cfg2 cfg3 cfg4; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${R}_ops_${op}.csv python tools/profile_ops.py $op > /dev/null 2>&1
done
EZB_PROFILE_PAIRS=8 ncu --set full --import-source on --clock-control none -k regex:corr_pyramid_gemm -c 1 -o gpurun_out/${R}_gemm_full python tools/profile_ops.py lookup > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:lookup_conv1x1 -c 1 -o gpurun_out/${R}_lookup_conv_full python tools/profile_ops.py lookup_conv > /dev/null 2>&1
ls -la gpurun_out | tail -20
