"""Replay storage for graphs: host-side list with safe / unsafe index sets and segment sampling
(semantics of reference gcbf/algo/buffer.py:11-95; the sampled list only defines the batch *shape*, the
arithmetic on it is the hot path)."""
import random
from typing import List

import numpy as np


class Buffer:
    MAX_SIZE = 100000

    def __init__(self):
        self._data = []
        self.safe_data: List[int] = []
        self.unsafe_data: List[int] = []

    data = property(lambda self: self._data)
    size = property(lambda self: len(self._data))

    def append(self, graph, is_safe: bool):
        self._data.append(graph)
        (self.safe_data if is_safe else self.unsafe_data).append(len(self._data) - 1)
        if len(self._data) > self.MAX_SIZE:
            self._drop_oldest(1)

    def _drop_oldest(self, k: int):
        del self._data[:k]
        self.safe_data = [i - k for i in self.safe_data if i >= k]
        self.unsafe_data = [i - k for i in self.unsafe_data if i >= k]

    def merge(self, other: 'Buffer'):
        base = len(self._data)
        self._data += other.data
        self.safe_data += [i + base for i in other.safe_data]
        self.unsafe_data += [i + base for i in other.unsafe_data]
        if len(self._data) > self.MAX_SIZE:
            self._drop_oldest(len(self._data) - self.MAX_SIZE)

    def clear(self):
        self._data = []
        self.safe_data, self.unsafe_data = [], []

    def sample_windows(self, n: int, m: int = 1, balanced_sampling: bool = False) -> List[tuple]:
        """The reference's segment sampling (buffer.py:57-95) as index windows: n centres (np.random.randint, or
        random.choices over the unsafe then the safe positions when balanced), each expanded to [c - m//2, c + m//2],
        clipped to the buffer and to the end of the previous window (no graph twice).  Consumes the host RNG streams
        exactly like the reference, so a seeded run samples the same graphs."""
        assert self.size >= max(n, m)
        if balanced_sampling:
            picks = []
            if self.unsafe_data:
                picks += random.choices(self.unsafe_data, k=n // 2)
            if self.safe_data:
                picks += random.choices(self.safe_data, k=n // 2)
            centres = sorted(picks)
        else:
            centres = np.sort(np.random.randint(0, self.size, n))
        out, hi = [], 0
        for c in centres:
            lo = max(int(c) - m // 2, hi)
            hi = min(int(c) + m // 2 + 1, self.size)
            out.append((lo, hi))
        return out

    def sample(self, n: int, m: int = 1, balanced_sampling: bool = False) -> list:
        """n centre indices, each expanded to a window of length <= m of consecutive graphs, de-duplicated."""
        out = []
        for lo, hi in self.sample_windows(n, m, balanced_sampling):
            out.extend(self._data[lo:hi])
        return out
