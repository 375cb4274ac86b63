for d in 0 1 2 4 8 15; do
python tools/bench_mesh.py > /dev/null 2>&1
MHE_SKIN_DEBUG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:skin -s 3 -c 1 --csv --log-file gpurun_out/skin_dbg_$d.csv python tools/bench_mesh.py > /dev/null 2>&1
echo "dbg=$d $(grep -o '\"[0-9,]*\"$' gpurun_out/skin_dbg_$d.csv | tail -1)"
done
