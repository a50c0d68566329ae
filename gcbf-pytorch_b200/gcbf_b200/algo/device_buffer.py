"""Device-resident replay ring (the replay half): the reference keeps every visited graph as a Python
`Data` object in a list (gcbf/algo/buffer.py:11-55) and `GCBF.update` collates the sampled ones with `Batch.from_data_list`
(gcbf/algo/gcbf.py:149-159) -- ~10 `torch.cat`s over B objects per inner iteration, all on the host.  Here a graph is what it
is a function of: its states [N, s] and the nominal control u_ref [n, a] it was stored with, kept in two device tensors;
edges, edge features and node types are rebuilt for the whole sampled batch by the batched radius-graph kernels (K1, K2),
which reproduce the stored ones bit for bit because they are the same deterministic function of the states.

Index semantics (append / merge / drop-oldest / segment sampling, host RNG consumption) are those of `Buffer`, which
mirrors the reference; `tests/test_host_cpu.py` checks both against each other on the CPU.
"""
from typing import List, Optional

import torch

from .buffer import Buffer


class DeviceReplay(Buffer):
    """Same interface as `Buffer`; payloads live in device rings instead of a Python list."""

    def __init__(self, device=None, capacity: int = 4096):
        super().__init__()
        self.device = torch.device(device) if device is not None else None
        self._cap = capacity
        self._states: Optional[torch.Tensor] = None      # [cap, N, s]
        self._u_ref: Optional[torch.Tensor] = None       # [cap, n, a]
        self._goal: Optional[torch.Tensor] = None        # [cap, n, goal_dim]: only filled by append_batch(goals=...) (vector rollouts)
        self._head = 0                                   # physical slot of logical index 0
        self._n = 0
        self._pending = []                               # (first logical index, count, pinned uint8 flags, CUDA event) of append_batch

    # ---- Buffer interface -------------------------------------------------------------------------------
    size = property(lambda self: self._n)

    @property
    def data(self):
        raise AttributeError('DeviceReplay holds no per-graph objects: use sample_batch(env, ...) or states_of(indices)')

    def _ensure(self, states: torch.Tensor, u_ref: torch.Tensor, need: int):
        if self._states is None:
            self.device = self.device or states.device
            cap = max(self._cap, need)
            self._states = torch.empty((cap,) + tuple(states.shape), device=self.device, dtype=torch.float32)
            self._u_ref = torch.empty((cap,) + tuple(u_ref.shape), device=self.device, dtype=torch.float32)
            self._cap = cap
        elif need > self._cap:
            cap = min(max(2 * self._cap, need), max(self.MAX_SIZE, need))
            order = self._slots(range(self._n))
            st = torch.empty((cap,) + tuple(self._states.shape[1:]), device=self.device, dtype=torch.float32)
            ur = torch.empty((cap,) + tuple(self._u_ref.shape[1:]), device=self.device, dtype=torch.float32)
            if self._n:
                st[:self._n] = self._states[order]
                ur[:self._n] = self._u_ref[order]
            if self._goal is not None:
                gl = torch.zeros((cap,) + tuple(self._goal.shape[1:]), device=self.device, dtype=torch.float32)
                if self._n:
                    gl[:self._n] = self._goal[order]
                self._goal = gl
            self._states, self._u_ref, self._cap, self._head = st, ur, cap, 0

    def _slots(self, logical) -> torch.Tensor:
        idx = torch.as_tensor(list(logical), dtype=torch.int64)
        return ((idx + self._head) % self._cap).to(self.device)

    def append(self, graph, is_safe: bool):
        """graph: anything with `.states` [N, s] and `.u_ref` [n, a] (a `Data` from env.step / env.reset)."""
        states, u_ref = graph.states.detach(), graph.u_ref.detach()
        if self._n == self.MAX_SIZE:
            self._drop_oldest(1)
        self._ensure(states, u_ref, self._n + 1)
        slot = (self._head + self._n) % self._cap
        self._states[slot].copy_(states)
        self._u_ref[slot].copy_(u_ref)
        (self.safe_data if is_safe else self.unsafe_data).append(self._n)
        self._n += 1

    def append_batch(self, states: torch.Tensor, u_ref: torch.Tensor, is_safe: torch.Tensor, goals: Optional[torch.Tensor] = None):
        """B graphs at once, from a vectorised rollout: states [B, N, s], u_ref [B, n, a], is_safe [B] bool ON THE DEVICE (and the
        goal set each graph was collected under, [B, n, goal_dim]).  No host sync: the safe / unsafe index lists (host side,
        reference buffer.py:18-30) are completed lazily from an asynchronous copy of the flags the next time they are needed."""
        B = int(states.shape[0])
        if self._n + B > self.MAX_SIZE:
            self._resolve_pending()
            self._drop_oldest(self._n + B - self.MAX_SIZE)
        self._ensure(states[0], u_ref[0], self._n + B)
        if goals is not None and (self._goal is None or self._goal.shape[0] != self._cap):
            old = self._goal
            self._goal = torch.zeros((self._cap,) + tuple(goals.shape[1:]), device=self.device, dtype=torch.float32)
            if old is not None:
                self._goal[:old.shape[0]] = old
        slots = (torch.arange(self._n, self._n + B, device=self.device) + self._head) % self._cap
        self._states[slots] = states.detach()
        self._u_ref[slots] = u_ref.detach()
        if goals is not None:
            self._goal[slots] = goals.detach()
        flags = torch.empty(B, dtype=torch.uint8).pin_memory()
        flags.copy_(is_safe.to(torch.uint8), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((self._n, B, flags, ev))
        self._n += B

    def _resolve_pending(self):
        for first, count, flags, ev in self._pending:
            ev.synchronize()
            for i, f in enumerate(flags.tolist()):
                (self.safe_data if f else self.unsafe_data).append(first + i)
        self._pending = []

    def sample_windows(self, n: int, m: int = 1, balanced_sampling: bool = False):
        self._resolve_pending()
        return super().sample_windows(n, m, balanced_sampling)

    def goals_of(self, indices) -> Optional[torch.Tensor]:
        return None if self._goal is None else self._goal[self._slots(indices)]

    def _drop_oldest(self, k: int):
        self._resolve_pending()
        self._head = (self._head + k) % self._cap
        self._n -= k
        self.safe_data = [i - k for i in self.safe_data if i >= k]
        self.unsafe_data = [i - k for i in self.unsafe_data if i >= k]

    def merge(self, other: 'DeviceReplay'):
        if other.size == 0:
            return
        self._resolve_pending()
        other._resolve_pending()
        total = self._n + other.size
        drop = max(0, total - self.MAX_SIZE)
        base = self._n
        src = other._slots(range(other.size))
        self._ensure(other._states[0], other._u_ref[0], min(total, self.MAX_SIZE) + drop)
        dst = self._slots(range(base, base + other.size))
        self._states[dst] = other._states[src]
        self._u_ref[dst] = other._u_ref[src]
        if other._goal is not None:
            if self._goal is None or self._goal.shape[0] != self._cap:
                old = self._goal
                self._goal = torch.zeros((self._cap,) + tuple(other._goal.shape[1:]), device=self.device, dtype=torch.float32)
                if old is not None:
                    self._goal[:old.shape[0]] = old
            self._goal[dst] = other._goal[src]
        self.safe_data += [i + base for i in other.safe_data]
        self.unsafe_data += [i + base for i in other.unsafe_data]
        self._n = total
        if drop:
            self._drop_oldest(drop)

    def clear(self):
        self._head, self._n = 0, 0
        self.safe_data, self.unsafe_data = [], []
        self._pending = []

    def sample(self, n: int, m: int = 1, balanced_sampling: bool = False) -> List[int]:
        """Logical indices of the sampled graphs (the list `Buffer.sample` would return objects for)."""
        out: List[int] = []
        for lo, hi in self.sample_windows(n, m, balanced_sampling):
            out.extend(range(lo, hi))
        return out

    # ---- device side -----------------------------------------------------------------------------------
    def states_of(self, indices) -> torch.Tensor:
        return self._states[self._slots(indices)]

    def u_ref_of(self, indices) -> torch.Tensor:
        return self._u_ref[self._slots(indices)]


def collate(env, parts) -> 'object':
    """Collated batch of the graphs `parts` = [(replay, indices), ...] name: one gather per ring, then the batched graph
    build (radius graph + edge features, K1 / K2) -- what `Batch.from_data_list` does object by object on the host."""
    states = torch.cat([r.states_of(idx) for r, idx in parts if len(idx)], dim=0)
    u_ref = torch.cat([r.u_ref_of(idx) for r, idx in parts if len(idx)], dim=0)
    B, N, s = states.shape
    data = env.add_communication_links(env.make_graph(states.reshape(B * N, s)))
    from ..data import Data
    data.update(Data(u_ref=u_ref.reshape(B * u_ref.shape[1], u_ref.shape[2])))
    goals = [r.goals_of(idx) for r, idx in parts if len(idx)]
    if goals and all(g is not None for g in goals):
        # graphs collected by a vectorised rollout carry the goal set of their own environment: the train step's
        # forward_graph (u_ref recomputed inside, simple_car.py:180) then uses it instead of one goal set shared by the batch
        g = torch.cat(goals, dim=0)
        data.update(Data(goal=g.reshape(B * g.shape[1], g.shape[2]).contiguous()))
    return data
