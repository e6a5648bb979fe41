import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import stove_oracle as so
from util import NoiseReplay, VARIANTS, make_model
from stove_b200 import synth
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
kw, seed = VARIANTS['plain']
oc, sd, model = make_model(kw, seed, att_gain=0.5)
n, T = 256, 8
x = synth.billiards(n, T, 3, res=32, seed=5)['x']
gen = torch.Generator().manual_seed(9)
noise = [torch.randn(n, 3, 12, 1, generator=gen, dtype=torch.float64) for _ in range(2)] + \
        [torch.randn(n, 3, 18, generator=gen, dtype=torch.float64) for _ in range(T - 2)]
P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
parts = {}
elbo_o, prop_o, _ = so.stove_forward(oc, P, x.double(), noise, parts=parts)
model._standard_normal = NoiseReplay(noise, 'cuda')
elbo, prop, _ = model(x.cuda(), 0)
dz = (prop['z'].double().cpu() - prop_o['z']).abs()
print('z err per t', dz.amax((0, 2, 3)))
per_seq = dz.amax((1, 2, 3))
top = per_seq.topk(5)
print('worst seqs', top.indices.tolist(), top.values.tolist())
b = int(top.indices[0])
print('z err worst seq per (t,o,feature)'); print(dz[b].amax(0))
print('z_dyn err', (prop['z_dyn'].double().cpu() - prop_o['z_dyn']).abs().amax((0, 2, 3)))
print('z_sup err', (prop['z_sup'].double().cpu() - prop_o['z_sup']).abs().amax((0, 2, 3)))
# encoder only
xb = so.bw_transform(x.double())
zp_o = so.encoder(oc, P, xb.flatten(0, 1))
with torch.no_grad():
    zp = model.sup.encoder(model.sup.encoder.rnn.weight_ih_l0.new_tensor(xb.flatten(0, 1).float().numpy()).cuda() if False else xb.flatten(0, 1).float().cuda())
print('encoder raw err', (zp.double().cpu() - zp_o).abs().max().item(), 'max', zp_o.abs().max().item())
# a pure fp32 CPU oracle run for comparison (what fp32 arithmetic itself costs)
P32 = {k: v.float() for k, v in sd.items()}
with torch.no_grad():
    e32, p32, _ = so.stove_forward(oc, P32, x.float(), [t.float() for t in noise])
d32 = (p32['z'].double() - prop_o['z']).abs()
print('fp32-oracle z err per t', d32.amax((0, 2, 3)), 'elbo', float(e32), float(elbo_o), float(elbo))
print('worst seqs fp32 oracle', d32.amax((1, 2, 3)).topk(5))
