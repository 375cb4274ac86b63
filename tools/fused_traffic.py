"""profiles/r2_fused_bwd_traffic.json from an `ncu --set full` report of the cluster-fused kernels (what bench.py's roofline.traffic
reads): DRAM bytes per launch of flow_bwd_fused_kernel, stamped with the hash of the kernel source it was captured from.
usage: python tools/fused_traffic.py gpurun_out/r2_fused.ncu-rep   (also writes profiles/r2_fused_kernels_ncu_full.csv)"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
rep = sys.argv[1]
subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_extract.py'), rep, os.path.join(ROOT, 'profiles', 'r2_fused_kernels_ncu_full.csv')], check=True)
rows = list(csv.reader(open(os.path.join(ROOT, 'profiles', 'r2_fused_kernels_ncu_full.csv'))))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    i = col[name]
    return float(r[i].replace(',', '')) * SCALE.get(units[i], 1.0)


launches = []
for r in data:
    launches.append({'kernel': r[col['Kernel Name']][:24], 'us': val(r, 'gpu__time_duration.sum'), 'dram_read': val(r, 'dram__bytes_read.sum'),
                     'dram_write': val(r, 'dram__bytes_write.sum'),
                     'tensor_pipe_active_pct': val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                     'warps_active_pct': val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')})
bwd = [l for l in launches if 'bwd' in l['kernel']]
src = os.path.join(ROOT, 'mhentropy_b200', 'csrc', 'flow_fused.cu')
out = {'kernel': 'flow_bwd_fused_kernel',
       'what': f'dram__bytes_read.sum + dram__bytes_write.sum per launch (6 layers of data gradients at B=64 x S=10), mean of {len(bwd)} launches, '
               f'ncu --set full --clock-control none ({rep}; raw page in profiles/r2_fused_kernels_ncu_full.csv)',
       'dram_bytes_per_launch': sum(l['dram_read'] + l['dram_write'] for l in bwd) / max(len(bwd), 1),
       'flow_fused_cu_sha256_16': hashlib.sha256(open(src, 'rb').read()).hexdigest()[:16], 'launches': launches}
json.dump(out, open(os.path.join(ROOT, 'profiles', 'r2_fused_bwd_traffic.json'), 'w'), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != 'launches'}, indent=1))
