"""Vectorised rollouts on the device: `num_envs` independent copies of one environment -- each with its own
initial state, obstacles and goal set -- advance as ONE batch per step.

The reference's data collection is one environment stepped from Python (gcbf/trainer/trainer.py:60-70): per env step one actor
forward on a single 16-agent graph (gcbf/algo/gcbf.py:128-139), one `env.step` (simple_car.py:146-176: u_ref, clamp, dynamics,
radius graph, collision masks) and one `unsafe_mask(...).any()` host sync -- ~5 ms of host time per 16-agent step.
Here one `step()` is: u_ref for all envs (`gcbf_u_ref_multi`, per-env goals), ONE batched radius graph + edge features, ONE actor
forward over the block-diagonal batch, the unsafe masks in one launch, ONE dynamics launch (`gcbf_step_fwd_multi`, every env is a
single graph -> reach-freeze branch), and an append of all `num_envs` graphs to the device-resident replay ring with the
safe / unsafe flags staying on the device.  The only host sync per step is the batched radius graph's edge count.

Episode semantics follow the reference's `step`: an env is done when t reaches max_episode_steps or when all its agents are
within dist2goal of their goals (checked on the device; the flags are read back every `reset_check_every` steps, so a finished
env may run a few extra steps before it is re-sampled -- its agents are frozen at their goals by then).
"""
import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from .. import _C, ops
from ..data import Data


class VectorRollout:
    def __init__(self, env, algo, num_envs: int, reset_check_every: int = 16, states: Optional[torch.Tensor] = None,
                 goals: Optional[torch.Tensor] = None):
        """states [num_envs * N, s] / goals [num_envs * n, goal_dim]: explicit initial conditions (e.g. synthetic BASELINE states:
        the reference's rejection sampler cannot place >~ 290 agents); default: env.reset() per env."""
        self.env, self.algo, self.B = env, algo, int(num_envs)
        self.n, self.N = env.num_agents, env.nodes_per_graph
        self.dev = env.device
        self.reset_check_every = reset_check_every
        self.states: Optional[torch.Tensor] = None       # [B * N, s]
        self.goals: Optional[torch.Tensor] = None        # [B * n, goal_dim]
        self.t = np.zeros(self.B, dtype=np.int64)
        self.steps = 0
        self._done_host = None
        self._auto_reset = states is None
        if states is None:
            self.reset()
        else:
            self.states = states.to(self.dev, torch.float32).contiguous().clone()
            self.goals = goals.to(self.dev, torch.float32).contiguous().clone()

    # ---- host side: initial conditions (the reference's rejection sampler, one env at a time) -------------------------
    def _sample_one(self):
        data = self.env.reset()
        return data.states.detach().clone(), self.env._goal.detach().clone()

    def reset(self, which=None):
        idx = range(self.B) if which is None else which
        if self.states is None:
            st, gl = self._sample_one()
            self.states = st.new_zeros(self.B * self.N, st.shape[1])
            self.goals = gl.new_zeros(self.B * self.n, gl.shape[1])
        for i in idx:
            st, gl = self._sample_one()
            self.states[i * self.N:(i + 1) * self.N] = st
            self.goals[i * self.n:(i + 1) * self.n] = gl
            self.t[i] = 0

    # ---- one vectorised step ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, prob: float = 0.0, store: bool = True) -> Dict[str, torch.Tensor]:
        """All envs advance one step under the actor (each env's action zeroed with probability `prob`, the reference's
        exploration schedule gcbf.py:131-132).  Returns device tensors: reach [B, n], collision [B, n] (after the step),
        is_safe [B] (before the step: what the stored graph is labelled with, gcbf.py:133-137)."""
        env, B, n, N = self.env, self.B, self.n, self.N
        cfg = env._cfg(B)
        st, ld = ops._mat(self.states)
        goal, ldg = ops._mat(self.goals)
        a_dim = env.action_dim
        u_ref = torch.empty(B * n, a_dim, device=self.dev, dtype=torch.float32)
        _C.call('gcbf_u_ref_multi', ctypes.byref(cfg), _C.ptr(st), ld, _C.ptr(goal), ldg, _C.ptr(env._gain()), _C.ptr(u_ref))
        data = env.add_communication_links(env.make_graph(self.states))          # ONE radius graph + edge features for all envs
        data.update(Data(u_ref=u_ref))
        action = self.algo.actor(data)                                           # ONE actor forward (block-diagonal batch)
        if prob > 0:
            keep = torch.from_numpy((np.random.rand(B) >= prob).astype(np.float32)).to(self.dev, non_blocking=True)
            action = action * keep.repeat_interleave(n).unsqueeze(1)
        masks = env._masks(data)
        is_safe = ~masks[1].view(B, n).any(dim=1)
        if store:
            buf = self.algo.buffer
            if not hasattr(buf, 'append_batch'):
                raise RuntimeError('vectorised rollouts store into the device replay ring: call algo.use_device_replay() first')
            buf.append_batch(self.states.view(B, N, -1), u_ref.view(B, n, a_dim), is_safe, self.goals.view(B, n, -1))
        nxt = torch.empty_like(st)
        pass_mask = torch.empty(B * n, a_dim, device=self.dev, dtype=torch.uint8)
        # every env is a SINGLE graph in the reference's loop: reach-freeze branch of dynamics() (dubins_car.py:126-130)
        _C.call('gcbf_step_fwd_multi', ctypes.byref(cfg), _C.ptr(st), ld, _C.ptr(action.contiguous()), _C.ptr(goal), ldg,
                _C.ptr(env._gain()), 1, _C.ptr(nxt), _C.ptr(pass_mask))
        self.states = nxt
        pd = env.POS_DIM
        agents = nxt.view(B, N, -1)[:, :n, :pd]
        reach = (agents - self.goals.view(B, n, -1)[:, :, :pd]).norm(dim=2) < env._params['dist2goal']
        coll_data = env.make_graph(nxt)
        collision = env._masks(coll_data)[2].view(B, n)
        self.t += 1
        self.steps += 1
        # episode ends: time limit known on the host; "all agents reached" read back with a delay (no sync in the common step)
        if not self._auto_reset:
            return dict(reach=reach, collision=collision, is_safe=is_safe, action=action, edge_count=int(data.edge_index.shape[1]))
        if self._done_host is not None and self.steps % self.reset_check_every == 0:
            flags, ev = self._done_host
            ev.synchronize()
            done = np.nonzero(flags.numpy())[0].tolist()
            self._done_host = None
            if done:
                self.reset(done)
        if self._done_host is None:
            flags = torch.empty(B, dtype=torch.bool).pin_memory()
            flags.copy_(reach.all(dim=1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._done_host = (flags, ev)
        timeout = np.nonzero(self.t >= env.max_episode_steps)[0].tolist()
        if timeout:
            self.reset(timeout)
        return dict(reach=reach, collision=collision, is_safe=is_safe, action=action, edge_count=int(data.edge_index.shape[1]))
