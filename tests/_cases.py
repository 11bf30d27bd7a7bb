"""Seeded synthetic graphs shared by the GPU parity tests."""
import numpy as np
import torch

from oracle import graph_oracle as go


def make_graph(seed, n, e0, rev="halves", self_loops=True, isolated=2):
    """rev: None | "halves" (add_reversed_edges layout) | "shuffled" (arbitrary flag pattern)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = max(n - isolated, 1)
    u, v = go.erdos_renyi(rng, m, e0, self_loops=self_loops)
    if rev is None:
        return u, v, None
    s, d, r = go.add_reversed_edges(u, v)
    if rev == "shuffled":
        p = rng.permutation(len(s))
        s, d, r = s[p], d[p], r[p]
    return s, d, r


def hub_graph(seed, n, e0, hub_deg):
    """ER graph plus one node with a very long in-segment (long-segment path of the reduce kernel)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u, v = go.erdos_renyi(rng, n, e0, self_loops=False)
    hu = rng.integers(1, n, size=hub_deg)
    u = np.concatenate([u, hu])
    v = np.concatenate([v, np.zeros(hub_deg, np.int64)])
    p = rng.permutation(len(u))
    return go.add_reversed_edges(u[p], v[p])


def t(x, device="cuda"):
    return None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(device)
