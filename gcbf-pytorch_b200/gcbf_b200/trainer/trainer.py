"""Trainer with the interface of reference gcbf/trainer/trainer.py:15-141 (host loop; the arithmetic it drives --
actor forward, env.step, algo.update -- is the kernel path).  TensorBoard is optional."""
import os
from time import time
from typing import Tuple

import numpy as np
import torch

from ..data import Data


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass


class Trainer:

    def __init__(self, env, env_test, algo, log_dir: str):
        self.env, self.env_test, self.algo, self.log_dir = env, env_test, algo, log_dir
        self.model_dir = os.path.join(log_dir, 'models')
        os.makedirs(self.model_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=os.path.join(log_dir, 'summary'))
        except Exception:   # tensorboard missing: keep training
            self.writer = _NullWriter()

    def train(self, steps: int, eval_interval: int, eval_epi: int):
        start = time()
        data = self.env.reset()
        verbose = None
        for step in range(1, steps + 1):
            data.update(Data(u_ref=self.env.u_ref(data)))
            action = self.algo.step(data, prob=1 - (step - 1) / steps)
            next_data, reward, done, info = self.env.step(action)
            next_data.update(Data(u_ref=self.env.u_ref(next_data)))
            self.algo.post_step(data, action, reward, done, next_data)
            data = self.env.reset() if done else next_data
            if self.algo.is_update(step):
                verbose = self.algo.update(step, self.writer)
            if eval_interval > 0 and step % eval_interval == 0:
                if eval_epi > 0:
                    reward, eval_info = self.eval(step, eval_epi)
                    print(f'step: {step}, time: {time() - start:.0f}s, reward: {reward:.2f}, ' +
                          ', '.join(f'{k}: {v}' for k, v in eval_info.items()))
                if verbose is not None:
                    print(f'step: {step}, ' + ', '.join(f'{k}: {v:.3f}' for k, v in verbose.items()))
                self.algo.save(os.path.join(self.model_dir, f'step_{step}'))
                self.algo._env = self.env
        print(f'> Done in {time() - start:.0f} seconds')

    def eval(self, step: int, eval_epi: int) -> Tuple[float, dict]:
        """Rollouts with the test-time controller `algo.apply` (reference trainer.py:95-141)."""
        rewards, safes = [], []
        self.algo._env = self.env_test
        for _ in range(eval_epi):
            data = self.env_test.reset()
            ep_reward, ep_safe, t = 0., [], 0
            while True:
                data.update(Data(u_ref=self.env_test.u_ref(data)))
                action = self.algo.apply(data)
                data, reward, done, info = self.env_test.step(action)
                ep_reward += float(np.mean(reward))
                ep_safe.append(info['safe'])
                t += 1
                if done:
                    break
            rewards.append(ep_reward)
            safes.append(float(np.mean(ep_safe)))
        self.writer.add_scalar('test/reward', float(np.mean(rewards)), step)
        self.writer.add_scalar('test/safe_rate', float(np.mean(safes)), step)
        return float(np.mean(rewards)), {'safe': float(np.mean(safes))}
