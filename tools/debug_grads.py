"""Debug helper (GPU box): per-tensor gradient error of the CUDA path vs the fp64 oracle, next to the error of the
fp32 CPU oracle vs fp64 (the intrinsic conditioning)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import gcbf_oracle as O
from helpers import *

env_name = sys.argv[1] if len(sys.argv) > 1 else 'DubinsCar'
mode = sys.argv[2] if len(sys.argv) > 2 else 'step'
n, obs, B, area = (24, 4, 3, 2.0) if env_name != 'SimpleDrone' else (12, 12, 2, 0.8)
meta = dict(env=env_name, n=n, obs=obs, graphs=B, area=area, seed=44, init_seed=2)
dev = torch.device('cuda:0')
sb = case_inputs(meta)
env, algo = seeded_algo(env_name, n, dev, 2, {'num_obs': sb.num_obs, 'area_size': area})
data = product_batch(env, sb, dev)
ob = oracle_batch(sb)
res = {}
for dt in (torch.float32, torch.float64):
    torch.set_default_dtype(dt)
    cbf = {k: v.to(dt) if v.is_floating_point() else v for k, v in sd_clone(algo.cbf).items()}
    act = {k: v.to(dt) if v.is_floating_point() else v for k, v in sd_clone(algo.actor).items()}
    K = ob['K'].to(dt) if ob['K'] is not None else None
    if mode == 'step':
        r = O.update_step(env_name, cbf, act, {}, {}, sb.states.to(dt), sb.goals.to(dt), ob['edge_index'], ob['u_ref'].to(dt), B, n, sb.num_obs, K=K, apply_optim=False)
        res[dt] = r['raw_grads']
        res[(dt, 'h')] = r['h']
    else:   # single forward of each net with a fixed cotangent
        for sdd in (cbf, act):
            for k in O.trainable_keys(sdd):
                sdd[k].requires_grad_(True)
        ea = O.edge_attr(env_name, sb.states.to(dt), ob['edge_index'])
        h = O.cbf_forward(cbf, ob['x'].to(dt), ea, ob['edge_index'], ob['agent_mask'])
        u = O.actor_forward(act, ob['x'].to(dt), ea, ob['edge_index'], ob['agent_mask'], ob['u_ref'].to(dt))
        g = torch.Generator().manual_seed(1)
        wh = torch.randn(h.shape, generator=g, dtype=torch.float32).to(dt)
        wu = torch.randn(u.shape, generator=g, dtype=torch.float32).to(dt)
        ((h * wh).sum() + (u * wu).sum()).backward()
        res[dt] = dict(cbf={k: cbf[k].grad for k in O.trainable_keys(cbf)}, actor={k: act[k].grad for k in O.trainable_keys(act)})
        res[(dt, 'h')] = h.detach()
torch.set_default_dtype(torch.float32)
if mode == 'step':
    out = algo.train_step(data, apply_optim=False)
    hg = out['h']
else:
    hg, ug = algo.cbf(data), algo.actor(data)
    g = torch.Generator().manual_seed(1)
    wh = torch.randn(hg.shape, generator=g).to(dev); wu = torch.randn(ug.shape, generator=g).to(dev)
    ((hg * wh).sum() + (ug * wu).sum()).backward()
print('h: gpu-vs-fp64', (hg.detach().cpu().double() - res[(torch.float64, 'h')]).abs().max().item(),
      'cpu32-vs-fp64', (res[(torch.float32, 'h')].double() - res[(torch.float64, 'h')]).abs().max().item())
for net, mod in (('cbf', algo.cbf), ('actor', algo.actor)):
    ref = res[torch.float64][net]; c32 = res[torch.float32][net]
    tot = torch.sqrt(sum((g ** 2).sum() for g in ref.values())).item()
    print(f'--- {net}: |g|={tot:.4e}   (errors relative to the NET gradient norm)')
    eg = ec = 0
    for name, p in mod.named_parameters():
        r = ref[name]
        e_gpu = (p.grad.cpu().double() - r).norm().item() / tot
        e_cpu = (c32[name].double() - r).norm().item() / tot
        eg += e_gpu ** 2; ec += e_cpu ** 2
        print(f'  {name:62s} |g|/tot={r.norm().item()/tot:8.2e} gpu={e_gpu:8.2e} cpu32={e_cpu:8.2e}')
    print(f'  TOTAL gpu={eg ** 0.5:.3e} cpu32={ec ** 0.5:.3e}')
