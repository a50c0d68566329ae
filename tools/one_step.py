"""One train step of a BASELINE config on one stream (for ncu captures): prints / dumps the tensor-core launches of that step in
launch order (product, M, N, K) so that an ncu capture of the same command can be labelled launch by launch.
    GCBF_TWO_STREAMS=0 python tools/one_step.py C3 [shapes.json]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200'), ROOT]
os.environ.setdefault('GCBF_TWO_STREAMS', '0')
import bench
from gcbf_b200 import ops
cfg = sys.argv[1] if len(sys.argv) > 1 else 'C3'
dev = torch.device('cuda', 0)
sb, env, algo = bench.build_case(cfg, dev, 0)
data = env.graph_from_states(sb.states.to(dev))
torch.cuda.synchronize()
ops.GEMM_TIMER.enable()
algo.train_step(data)
torch.cuda.synchronize()
recs = ops.GEMM_TIMER._collect_native()
ops.GEMM_TIMER.disable()
names = ('forward', 'data-grad', 'weight-grad')
tc = [dict(product=names[k], M=M, N=N, K=K, ms=round(ms, 4)) for ms, fl, k, M, N, K in recs if k in (0, 1, 2)]
print(f'{cfg}: {len(tc)} tensor-core launches in one step, E = {int(data.edge_index.shape[1])}')
if len(sys.argv) > 2:
    json.dump(dict(config=cfg, edges=int(data.edge_index.shape[1]), launches=tc), open(sys.argv[2], 'w'), indent=0)
