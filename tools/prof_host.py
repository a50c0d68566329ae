"""Host-side profile of GCBF.train_step (cProfile, per-step device sync so queue back-pressure does not hide host time)."""
import cProfile, os, pstats, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gcbf-pytorch_b200'))
sys.path.insert(0, ROOT)
import bench
dev = torch.device('cuda', 0)
sb, env, algo = bench.build_case(sys.argv[1] if len(sys.argv) > 1 else 'C2', dev, 0)
data = env.graph_from_states(sb.states.to(dev))
for _ in range(3):
    algo.train_step(data)
torch.cuda.synchronize()
pr = cProfile.Profile()
t_host = 0.0
for _ in range(6):
    t0 = time.perf_counter()
    pr.enable()
    algo.train_step(data)
    pr.disable()
    t_host += time.perf_counter() - t0
    torch.cuda.synchronize()
print(f'host time per step (profiled): {t_host / 6 * 1e3:.2f} ms')
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(45)
st.sort_stats('cumulative').print_stats(35)
