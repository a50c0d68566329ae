import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
from helpers import *
from gcbf_b200 import ops
from gcbf_b200.data import agent_row_index
from gcbf_b200.nn.gnn import cached_rowptr
dev = torch.device('cuda:0')
meta = dict(env='DubinsCar', n=24, obs=4, graphs=3, area=2.0, seed=44, init_seed=2)
sb = case_inputs(meta)
env, algo = seeded_algo('DubinsCar', 24, dev, 2, {'num_obs': 4, 'area_size': 2.0})
data = product_batch(env, sb, dev)
layer = algo.cbf.feat_transformer.module_0
spec = layer.net_spec(algo.cbf.feat_2_CBF)
rowptr = cached_rowptr(data.edge_index, data.x.shape[0])
out, ctx = ops.net_forward(spec, data.x, data.edge_attr, data.edge_index, rowptr, agent_row_index(data), None, True)
c_phi, c_gate, c_gamma, c_head, msg, att, Nn, E = ctx
print('inv_sigma gamma:', [float(s) for s in c_gamma.inv_sigma], 'phi:', [float(s) for s in c_phi.inv_sigma])
g = torch.Generator().manual_seed(1)
d_feat = torch.randn(c_gamma.acts[-1].shape, generator=g).to(dev)
d_gin, grads = ops.mlp_backward(c_gamma, spec.gamma, d_feat, True)
# fp64 torch reference on the same saved input
x0 = c_gamma.acts[0].double().requires_grad_(True)
Ws = [L.W.detach().double().requires_grad_(True) for L in spec.gamma]
x = x0
for i, L in enumerate(spec.gamma):
    u, v = c_gamma.uv[i]
    sigma = torch.dot(u.double(), Ws[i] @ v.double())
    x = torch.nn.functional.linear(x, Ws[i] / sigma, L.b.detach().double())
    print(f'layer {i}: fwd maxdiff {(x.relu() if i < 2 else x).sub(c_gamma.acts[i + 1].double()).abs().max().item():.3e} sigma {sigma.item():.8f}')
    if i < 2:
        x = torch.relu(x)
(x * d_feat.double()).sum().backward()
print('d_gin rel err', ((d_gin.double() - x0.grad).norm() / x0.grad.norm()).item())
for i in range(3):
    print(f'dW[{i}] rel err', ((grads[i][0].double() - Ws[i].grad).norm() / Ws[i].grad.norm()).item())
# stepwise: dz_2 reference
dz2_ref = (d_feat.double() @ (Ws[2].detach() * c_gamma.inv_sigma[2].double())) * (c_gamma.acts[2].double() > 0)
dz2 = ops.linear_bwd_data(d_feat, spec.gamma[2].W, c_gamma.inv_sigma[2], c_gamma.acts[2])
print('dz2 rel err (with alpha)', ((dz2.double() - dz2_ref).norm() / dz2_ref.norm()).item())
dz2b = ops.linear_bwd_data(d_feat, spec.gamma[2].W, None, c_gamma.acts[2])
print('dz2 rel err (alpha None vs ref/alpha)', ((dz2b.double() * c_gamma.inv_sigma[2].double() - dz2_ref).norm() / dz2_ref.norm()).item())
