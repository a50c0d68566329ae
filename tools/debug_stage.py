import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import gcbf_oracle as O
import torch.nn.functional as F
from helpers import *
from gcbf_b200 import ops
from gcbf_b200.data import agent_row_index
from gcbf_b200.nn.gnn import cached_rowptr
dev = torch.device('cuda:0')
meta = dict(env='DubinsCar', n=24, obs=4, graphs=3, area=2.0, seed=44, init_seed=2)
sb = case_inputs(meta)
env, algo = seeded_algo('DubinsCar', 24, dev, 2, {'num_obs': 4, 'area_size': 2.0})
data = product_batch(env, sb, dev)
ob = oracle_batch(sb)
# ---- CPU fp64 reference with retained intermediates (same code path as oracle.cbf_forward) ----
torch.set_default_dtype(torch.float64)
sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd_clone(algo.cbf).items()}
for k in O.trainable_keys(sd): sd[k].requires_grad_(True)
x = ob['x'].double(); ei = ob['edge_index']; ea = O.edge_attr('DubinsCar', sb.states.double(), ei)
keep = {}
def mlp(prefix, x, n, sn, out_act=None, tag=''):
    for i in range(n):
        key = f'{prefix}.net.{2*i}'
        W = O._sn_weight(sd, key) if sn else sd[key + '.weight']
        x = F.linear(x, W, sd[key + '.bias'])
        if i < n - 1: x = torch.relu(x)
        x.retain_grad(); keep[f'{tag}{i}'] = x
    if out_act == 'tanh': x = torch.tanh(x)
    return x
P = 'feat_transformer.module_0'
info = torch.cat([x[ei[1]], x[ei[0]], ea], 1)
m = mlp(P + '.phi', info, 3, True, tag='phi')
gate = mlp(P + '.aggr_module.gate_nn', m, 3, False, tag='gate')
att = O.segment_softmax(gate, ei[1], x.shape[0])
aggr = torch.zeros(x.shape[0], 256).index_add(0, ei[1], att * m); aggr.retain_grad(); keep['aggr'] = aggr
feat_all = mlp(P + '.gamma', torch.cat([aggr, x], 1), 3, True, tag='gamma')
feat = feat_all[ob['agent_mask']]; feat.retain_grad(); keep['feat'] = feat
h = mlp('feat_2_CBF', feat, 4, False, 'tanh', tag='head')
g = torch.Generator().manual_seed(1)
wh = torch.randn(h.shape, generator=g, dtype=torch.float32).double()
(h * wh).sum().backward()
torch.set_default_dtype(torch.float32)
# ---- GPU pipeline, stage by stage ----
layer = algo.cbf.feat_transformer.module_0
spec = layer.net_spec(algo.cbf.feat_2_CBF)
rowptr = cached_rowptr(data.edge_index, data.x.shape[0])
ridx = agent_row_index(data)
out, ctx = ops.net_forward(spec, data.x, data.edge_attr, data.edge_index, rowptr, ridx, None, True)
c_phi, c_gate, c_gamma, c_head, msg, attg, Nn, E = ctx
am = ob['agent_mask']
def rel(a, b): return ((a.cpu().double() - b).norm() / (b.norm() + 1e-300)).item()
print('fwd h', rel(out, h.detach()), ' gamma acts', [rel(c_gamma.acts[i + 1], keep[f'gamma{i}'].detach()[am]) for i in range(3)])
print('mask mismatch gamma act1:', ((c_gamma.acts[2].cpu() > 0) != (keep['gamma1'].detach()[am] > 0)).sum().item())
d_hin, g_head = ops.mlp_backward(c_head, spec.head, wh.float().to(dev), True)
print('d_feat', rel(d_hin, keep['feat'].grad))
# manual gamma backward
dz = d_hin
for l in (2, 1, 0):
    L = spec.gamma[l]
    ref_dz = keep[f'gamma{l}'].grad[am]     # grad wrt post-activation output of layer l
    if l < 2:   # my dz is grad wrt PRE-activation (masked); reference grad wrt post-relu output -> mask it
        ref_dz = ref_dz * (keep[f'gamma{l}'].detach()[am] > 0)
    print(f'gamma dz[{l}] rel err', rel(dz, ref_dz))
    dW, db = ops.linear_bwd_weight(dz, c_gamma.acts[l], c_gamma.inv_sigma[l])
    print(f'   db[{l}] rel err', rel(db, sd[f"{P}.gamma.net.{2*l}.bias"].grad), ' raw dW vs (ref before fixup n/a)')
    dz = ops.linear_bwd_data(dz, L.W, c_gamma.inv_sigma[l], c_gamma.acts[l] if l > 0 else None)
