set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1v3_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1v3_pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r1v3_bench.json 2> gpurun_out/r1v3_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1v3_ref.json 2> gpurun_out/r1v3_ref.err
python tools/trace_step.py > gpurun_out/r1v3_trace.log 2>&1; cp gpurun_out/trace_step.csv gpurun_out/r1v3_timeline.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1v3_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r1v3_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'flow_.wd_fused|hypothesis_rows|cond_fwd_direct' -s 8 -c 5 -o gpurun_out/r1v3_full python bench.py --steps 2 --warmup 3 --profile --no-graph > gpurun_out/r1v3_ncu2.log 2>&1
ls -la gpurun_out | tail -12
