# round-end artifact run on one GPU: smoke, GPU tests, the bench line, the reference arm, then (after the same command exited 0 plain)
# the ncu launch list of the step and one full capture of the cluster-fused kernels (tools/fused_traffic.py reads it)
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --profile --no-configs > gpurun_out/r2_profile_plain.json 2> gpurun_out/r2_profile_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --profile --no-configs > gpurun_out/r2_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --profile --no-configs --no-graph > gpurun_out/r2_profile_plain_nograph.json 2> gpurun_out/r2_profile_plain_nograph.err && \
ncu --set full --clock-control none --import-source on -k regex:'flow_.wd_fused' -s 3 -c 3 -f -o gpurun_out/r2_fused python bench.py --steps 2 --warmup 3 --profile --no-configs --no-graph > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out | tail -8
