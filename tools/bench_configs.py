"""Secondary workloads of BASELINE.json through the drop-in API on one GPU (not bench lines; for the record in profiles/):
config 3: inference sampling B=256 x S=100 (flow sample + MANO mesh + reprojection), config 4: NLL scoring of 16384 poses; plus the
SURVEY 8f-1 evaluation kernels at the reference's evaluation size (200 hypotheses per image, CrossModalHand.py:357-361): top-k selection and
the MHEntLoss metrics, with their achieved bandwidth (they read the (N, B, .) outputs once: HBM / launch bound)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import MHEntHead
from mhentropy_b200.mano_assets import synthetic_mano

dev = torch.device('cuda')
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


out = {}
B, S = 256, 100
feat = torch.randn(B, 512, device=dev)
z0 = torch.randn(B * S, 45, device=dev) * 0.8
z_det = torch.cat([0.5 * torch.randn(B, 3), 0.02 * torch.randn(B, 10), torch.randn(B, 1) * 0.1 - 1.2, 0.1 * torch.randn(B, 2)], 1).to(dev)
res = {}


def sample():
    res['o'] = head.sample(feat, N=S, temp=0.8, mods={'uv', 'xyz', 'verts'}, z0=z0, z_det=z_det)


ms = timeit(sample)
o = res['o']
assert all(torch.isfinite(v).all() for v in o.values() if torch.is_tensor(v) and v.is_floating_point())
out['config3_inference_sampling'] = {'B': B, 'S': S, 'ms': ms, 'hypotheses_per_s': B * S / (ms * 1e-3),
                                     'outputs': {k: list(v.shape) for k, v in o.items() if torch.is_tensor(v)}}
R = 16384
x = 0.5 * torch.randn(R, 45, device=dev)
featr = torch.randn(R, 512, device=dev)


def nll():
    with torch.no_grad():
        res['lp'] = head.q_z_giv_i.log_prob(x, logvar=featr)


ms = timeit(nll)
assert torch.isfinite(res['lp']).all()
out['config4_nll_scoring'] = {'rows': R, 'ms': ms, 'poses_per_s': R / (ms * 1e-3)}
# ---- MANO mesh alone at the config-3 row count: pose kernel + pose-blend GEMM (tensor cores) + shape blend / LBS (HBM-bound part)
from mhentropy_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
Rm = B * S
consts = head.mano_dec.mano_layer._consts(dev)
zz = torch.cat([0.5 * torch.randn(Rm, 3), 1.0 * torch.randn(Rm, 45), 0.02 * torch.randn(Rm, 10), torch.zeros(Rm, 3)], 1).to(dev).contiguous()
verts, jtr = torch.empty(Rm, 778, 3, device=dev), torch.empty(Rm, 21, 3, device=dev)
mwsb = lib().mhe_mano_workspace_bytes(Rm, 0)
mws = torch.empty(mwsb, dtype=torch.uint8, device=dev)


def mesh():
    check(lib().mhe_mano_fwd(consts, zz.data_ptr(), 61, zz.data_ptr() + 48 * 4, 61, Rm, 1, ptr(verts), ptr(jtr), None, ptr(mws), mwsb,
                             stream_ptr(dev)), 'mano_fwd')


ms = timeit(mesh, n=10)
lbs_bytes = Rm * (9336 + 9344)            # vertices written + pose offsets read per row
out['mano_mesh_fwd'] = {'rows': Rm, 'ms': ms, 'rows_per_s': Rm / (ms * 1e-3), 'lbs_algorithmic_bytes': lbs_bytes,
                        'GB_per_s_whole_mesh_path': lbs_bytes / (ms * 1e-3) / 1e9,
                        'pose_blend_gflop': Rm * 630180 / 1e9, 'note': 'whole mesh path (pose kernel + split + pose-blend GEMM + skinning) timed '
                        'together; the LBS bytes over that time is a lower bound of the skinning kernel\'s own bandwidth'}
# ---- section 8f-1: hypothesis selection + multi-hypothesis metrics, N = 200 hypotheses x B = 256 images
from mhentropy_b200 import hypothesis_metrics, topk_hypotheses  # noqa: E402
N, Bm = 200, 256
pose3d = torch.randn(Bm, 63, device=dev)
xyz = pose3d[None] + 0.3 * torch.randn(N, Bm, 63, device=dev)
crop_uv = torch.rand(Bm, 42, device=dev) * 2 - 1
uv = (crop_uv[None] + 1) / 2 * 256 + 8. * torch.randn(N, Bm, 42, device=dev)
scale = 0.05 + 0.1 * torch.rand(Bm, device=dev)
vis = (torch.rand(Bm, 21, device=dev) < 0.7).float()
log_q = torch.randn(N, Bm, device=dev)
ms = timeit(lambda: res.__setitem__('m', hypothesis_metrics(xyz, uv, pose3d, scale, crop_uv, vis)), n=20)
nbytes = 2 * N * Bm * (63 + 42) * 4          # two passes over xyz / uv (error terms, then the centred second moment)
out['metrics_8f1'] = {'N': N, 'B': Bm, 'ms': ms, 'algorithmic_bytes': nbytes, 'GB_per_s': nbytes / (ms * 1e-3) / 1e9,
                      'hypotheses_per_s': N * Bm / (ms * 1e-3)}
ms = timeit(lambda: res.__setitem__('k', topk_hypotheses(log_q, 20)), n=20)
out['topk_8f1'] = {'N': N, 'B': Bm, 'k': 20, 'ms': ms, 'hypotheses_per_s': N * Bm / (ms * 1e-3)}
print(json.dumps(out))
