mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-sustained"
$B > gpurun_out/ev_plain.json 2> gpurun_out/ev_plain.err || exit 1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2x.csv $B > gpurun_out/ev_ncu1.log 2>&1
timeout 300 ncu --nvtx --nvtx-include "expert_ffn/" --set full --clock-control none -c 2 -o gpurun_out/expert_ffn_r2x -f $B > gpurun_out/ev_ncu2.log 2>&1
python tools/ln_prof.py > /dev/null 2>&1
FWD=1 timeout 300 ncu --kernel-name regex:gemm_ln --launch-skip 47 --launch-count 6 --set full --import-source on --clock-control none -o gpurun_out/gemm_ln_r2x -f python tools/one_forward.py > gpurun_out/ev_ncu3.log 2>&1
ls -la gpurun_out | tail -8
