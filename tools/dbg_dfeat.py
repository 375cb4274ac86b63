import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mhentropy_b200 import RealNVP
from oracle import flow_oracle as fo
fx = dict(np.load('/root/repo/tests/golden/flow_prod.npz'))
sd = fo.init_state_dict(seed=int(fx['seed']))
PROD = dict(dim=45, tsfm_on=512, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)
def rel(a,b):
    a=a.detach().cpu().double().numpy(); b=np.asarray(b,np.float64); return np.abs(a-b).max()/np.abs(b).max()
# fp64 oracle
sd64={k:v.double().clone().requires_grad_(k!='mask') for k,v in sd.items()}
f64=torch.tensor(fx['feat']).double().requires_grad_(True); z64=torch.tensor(fx['z0']).double().requires_grad_(True)
x64=fo.sample(sd64,z64,f64); (x64*torch.tensor(fx['wx']).double()).sum().backward()
print('golden(fp32 ref) vs fp64: dfeat', rel(torch.tensor(fx['sample_dfeat']), f64.grad.numpy()))
for prec in ('fp32','bf16x3'):
    flow = RealNVP(**PROD); flow.load_state_dict(sd); flow.precision=prec; flow=flow.cuda()
    feat=torch.tensor(fx['feat']).cuda().requires_grad_(True); z0=torch.tensor(fx['z0']).cuda().requires_grad_(True)
    x=flow.forward_p(z0,cond=feat); (x*torch.tensor(fx['wx']).cuda()).sum().backward()
    print(prec,'x abs err vs fp64', (x.detach().cpu().double()-x64.detach()).abs().max().item(), 'dfeat vs golden', rel(feat.grad, fx['sample_dfeat']), 'vs fp64', rel(feat.grad, f64.grad.numpy()), 'dz0 vs fp64', rel(z0.grad, z64.grad.numpy()))
    g=dict(flow.named_parameters())
    worst=max((abs(float(p.grad.double().norm())-float(fx['gsnorm/'+k]))/float(fx['gsnorm/'+k]),k) for k,p in g.items())
    print('   worst grad-norm rel diff', worst)
