"""Fused vs per-GEMM flow pass at several row counts (forward without saving): where should MHE_FUSED_MAX_ROWS sit?"""
import os, subprocess, sys
for B, S in ((64, 16), (128, 10), (128, 16), (256, 10), (256, 16)):
    for mr in ('100000', '0'):
        env = dict(os.environ, MHE_FUSED_MAX_ROWS=mr)
        out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), 'bench_flow_pass.py'), str(B), str(S)], env=env, capture_output=True, text=True)
        print(out.stdout.strip() or out.stderr[-400:])
