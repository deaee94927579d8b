# Final-state evidence of a round, in one gpurun call: GPU suite, smoke, bench lines (sampling, training, reference arm), the
# ncu launch list of the bench command and the --set full capture of the grouped expert GEMMs.  TAG names the outputs.
TAG=${TAG:-r2z}
mkdir -p gpurun_out
(timeout 1100 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/ev_tests_$TAG.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/ev_smoke_$TAG.log
python bench.py > gpurun_out/ev_bench_$TAG.json 2> gpurun_out/ev_bench_$TAG.err || exit 1
python bench.py --workload train --steps 5 > gpurun_out/ev_bench_train_$TAG.json 2> gpurun_out/ev_bench_train_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev_bench_ref_$TAG.json 2> gpurun_out/ev_bench_ref_$TAG.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-sustained"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ev_ncu1.log 2>&1
timeout 300 ncu --nvtx --nvtx-include "expert_ffn/" --set full --clock-control none -c 2 -o gpurun_out/expert_ffn_$TAG -f $B > gpurun_out/ev_ncu2.log 2>&1
tail -3 gpurun_out/ev_tests_$TAG.log; tail -2 gpurun_out/ev_smoke_$TAG.log; head -c 300 gpurun_out/ev_bench_$TAG.json; echo; head -c 300 gpurun_out/ev_bench_train_$TAG.json; echo; head -c 200 gpurun_out/ev_bench_ref_$TAG.json; echo
ls -la gpurun_out | tail -6
