"""Compact CSV of an ncu report's raw page (the columns the roofline discussion uses): python tools/ncu_extract.py in.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEYS = ('Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration', 'dram__bytes', 'pipe_tensor_cycles_active', 'warps_active', 'lts__t_bytes.sum',
        'lts__throughput', 'sm__throughput', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'launch__registers',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'gpu__dram_throughput', 'l1tex__throughput', 'lts__t_sector_hit_rate')
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEYS)]
with open(sys.argv[2], 'w', newline='') as fh:
    w = csv.writer(fh)
    w.writerow([hdr[i] for i in keep])
    w.writerow([units[i] for i in keep])
    for d in data:
        w.writerow([d[i] for i in keep])
print(f'{len(data)} launches, {len(keep)} columns -> {sys.argv[2]}')
