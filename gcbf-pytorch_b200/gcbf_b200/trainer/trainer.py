"""Outer training loop with the call surface of the reference's `gcbf.trainer.Trainer` (gcbf/trainer/trainer.py:15-141):
`Trainer(env, env_test, algo, log_dir).train(steps, eval_interval, eval_epi)` and `.eval(step, eval_epi)`.

The loop itself is host glue; what it drives -- the actor forward inside `algo.step`, `env.step`, `algo.update`, and
`algo.apply` during evaluation -- is the kernel path.  Behaviour kept from the reference: the exploration probability decays
linearly from 1 to 0 over the run, u_ref is attached to a graph before the algorithm sees it, checkpoints go to
`<log_dir>/models/step_<k>`, the evaluation episodes use the test-time controller and report the mean episode reward, the
fraction of agents that never collided and the fraction that reached their goal.  TensorBoard is optional (scalars are dropped
when it is not installed); progress lines go to stdout."""
import os
import time
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from ..data import Data


class _DropScalars:
    def add_scalar(self, *args, **kwargs):
        return None


def _make_writer(path: str):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir=path)
    except Exception:
        return _DropScalars()


class Trainer:

    def __init__(self, env, env_test, algo, log_dir: str):
        self.env, self.env_test, self.algo = env, env_test, algo
        self.log_dir = log_dir
        self.model_dir = os.path.join(log_dir, 'models')
        os.makedirs(self.model_dir, exist_ok=True)
        self.writer = _make_writer(os.path.join(log_dir, 'summary'))

    # ---- rollout --------------------------------------------------------------------------------------------
    @staticmethod
    def _with_u_ref(env, graph):
        graph.update(Data(u_ref=env.u_ref(graph)))
        return graph

    def _rollout_step(self, graph, explore_prob: float):
        """One environment transition of the training env; returns the graph the next transition starts from."""
        env, algo = self.env, self.algo
        self._with_u_ref(env, graph)
        action = algo.step(graph, prob=explore_prob)
        nxt, reward, done, _ = env.step(action)
        self._with_u_ref(env, nxt)
        algo.post_step(graph, action, reward, done, nxt)
        return env.reset() if done else nxt

    # ---- training -------------------------------------------------------------------------------------------
    def train(self, steps: int, eval_interval: int, eval_epi: int):
        t0 = time.time()
        graph = self.env.reset()
        last_update: Optional[Dict[str, float]] = None
        for k in range(steps):
            step = k + 1
            graph = self._rollout_step(graph, explore_prob=1.0 - k / steps)
            if self.algo.is_update(step):
                last_update = self.algo.update(step, self.writer)
            if eval_interval > 0 and step % eval_interval == 0:
                self._checkpoint_and_report(step, eval_epi, last_update, time.time() - t0)
        print(f'> Done in {time.time() - t0:.0f} seconds')

    def _checkpoint_and_report(self, step: int, eval_epi: int, last_update, elapsed: float):
        if eval_epi > 0:
            reward, info = self.eval(step, eval_epi)
            extras = ''.join(f', {name}: {value}' for name, value in info.items())
            print(f'step: {step}, time: {elapsed:.0f}s, reward: {reward:.2f}{extras}')
        if last_update is not None:
            print(f'step: {step}' + ''.join(f', {name}: {value:.3f}' for name, value in last_update.items()))
        self.algo.save(os.path.join(self.model_dir, f'step_{step}'))
        self.algo._env = self.env                      # eval() pointed the algorithm at the test env

    # ---- evaluation -----------------------------------------------------------------------------------------
    def _episode(self, env) -> Tuple[float, float, torch.Tensor]:
        """One episode under `algo.apply`: (sum over steps of the mean agent reward, fraction of agents that never collided,
        per-agent reach flags of the last step)."""
        never_hit = torch.ones(env.num_agents, dtype=torch.bool)
        reach = torch.zeros(env.num_agents, dtype=torch.bool)
        total = 0.0
        graph = env.reset()
        done = False
        while not done:
            action = self.algo.apply(self._with_u_ref(env, graph))
            graph, reward, done, info = env.step(action)
            total += float(np.mean(reward))
            hit = info.get('collision')
            if hit is not None and len(hit):
                never_hit[torch.as_tensor(hit).cpu().long()] = False
            if 'reach' in info:
                reach = torch.as_tensor(info['reach']).cpu().bool()
        return total, float(never_hit.float().mean()), reach

    def eval(self, step: int, eval_epi: int) -> Tuple[float, dict]:
        env = self.env_test
        self.algo._env = env
        rewards, safe, reach = [], [], torch.zeros(env.num_agents, dtype=torch.bool)
        for _ in range(eval_epi):
            r, s, reach = self._episode(env)
            rewards.append(r)
            safe.append(s)
        mean_reward, mean_safe = float(np.mean(rewards)), float(np.mean(safe))
        self.writer.add_scalar('test/reward', mean_reward, step)
        self.writer.add_scalar('test/safe_rate', mean_safe, step)
        return mean_reward, {'safe': round(mean_safe, 2), 'reach': round(float(reach.float().mean()), 2)}
