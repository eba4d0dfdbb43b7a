"""Network descriptors mirroring mava/networks.py (MLPTorso :39-58, DiscreteActionHead :88-124,
FeedForwardActor :172-183, FeedForwardValueNet :186-207).

The classes hold shapes and build parameters; the arithmetic runs in the CUDA kernels
(csrc/mlp_f32.cu, csrc/mlp_tc.cu).  Parameters are one flat float32 vector per network in flax
order; ``to_flax_tree`` / ``from_flax_tree`` convert to the reference's pytree naming so weights can
be exchanged with a JAX run.
"""
from __future__ import annotations

import importlib
from typing import Any, Dict, List, Sequence

import numpy as np

from . import native


def instantiate(node: Dict[str, Any], **kwargs: Any) -> Any:
    """hydra.utils.instantiate for ``_target_`` nodes; ``mava.*`` targets resolve to this package."""
    node = dict(node)
    target = node.pop("_target_")
    if target.startswith("mava."):
        target = "mava_b200." + target[len("mava."):]
    mod, _, name = target.rpartition(".")
    cls = getattr(importlib.import_module(mod), name)
    node.update(kwargs)
    return cls(**node)


def orthogonal(rng: np.random.Generator, shape: Sequence[int], scale: float) -> np.ndarray:
    """flax.linen.initializers.orthogonal(scale) (QR of a normal matrix, sign-fixed).

    The draw uses numpy's generator, not jax.random.normal, so values differ from a JAX run with
    the same seed; the distribution is the same."""
    n_rows, n_cols = int(np.prod(shape[:-1])), int(shape[-1])
    big, small = max(n_rows, n_cols), min(n_rows, n_cols)
    a = rng.standard_normal((big, small))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))[None, :]
    if n_rows < n_cols:
        q = q.T
    return (scale * q.reshape(shape)).astype(np.float32)


class MLPTorso:
    def __init__(self, layer_sizes: Sequence[int], activation: str = "relu",
                 use_layer_norm: bool = False):
        self.layer_sizes = [int(s) for s in layer_sizes]
        if activation != "relu" or use_layer_norm:
            raise NotImplementedError(
                "the mava_b200 kernels implement MLPTorso with relu and no layer norm "
                "(configs/network/mlp.yaml defaults)")
        self.activation, self.use_layer_norm = activation, use_layer_norm


class DiscreteActionHead:
    def __init__(self, action_dim: int):
        self.action_dim = int(action_dim)


class _FeedForward:
    """Shared flat-parameter handling of the two feed-forward networks."""

    torso: MLPTorso
    out_dim: int
    head_scale: float
    head_path: List[str]

    def desc(self, num_agents: int, view_dim: int, add_agent_id: bool, input_mode: int):
        if len(self.torso.layer_sizes) != 2:
            raise NotImplementedError("MLPTorso with exactly two hidden layers is supported")
        h1, h2 = self.torso.layer_sizes
        return native.mlp_desc(input_mode, add_agent_id, num_agents, view_dim, h1, h2, self.out_dim)

    def shapes(self, in_dim: int):
        h1, h2 = self.torso.layer_sizes
        return [(in_dim, h1), (h1,), (h1, h2), (h2,), (h2, self.out_dim), (self.out_dim,)]

    def init(self, key: np.ndarray, in_dim: int) -> np.ndarray:
        """Parameters as the reference initialises them: orthogonal(sqrt 2) torso kernels,
        orthogonal(head_scale) head kernel, zero biases (networks.py:54,114,205)."""
        rng = np.random.default_rng([int(key[0]), int(key[1])])
        out = []
        for i, s in enumerate(self.shapes(in_dim)):
            if len(s) == 1:
                out.append(np.zeros(s, np.float32))
            else:
                out.append(orthogonal(rng, s, self.head_scale if i == 4 else float(np.sqrt(2))))
        return np.concatenate([p.ravel() for p in out])

    def to_flax_tree(self, flat: np.ndarray, in_dim: int) -> Dict[str, Any]:
        parts, off = [], 0
        for s in self.shapes(in_dim):
            n = int(np.prod(s))
            parts.append(np.asarray(flat[off:off + n]).reshape(s))
            off += n
        torso = {"Dense_0": {"kernel": parts[0], "bias": parts[1]},
                 "Dense_1": {"kernel": parts[2], "bias": parts[3]}}
        head = {"kernel": parts[4], "bias": parts[5]}
        tree: Dict[str, Any] = {"torso": torso}
        node = tree
        for k in self.head_path[:-1]:
            node = node.setdefault(k, {})
        node[self.head_path[-1]] = head
        return {"params": tree}

    def from_flax_tree(self, tree: Dict[str, Any]) -> np.ndarray:
        p = tree["params"]
        head = p
        for k in self.head_path:
            head = head[k]
        t = p["torso"]
        parts = [t["Dense_0"]["kernel"], t["Dense_0"]["bias"], t["Dense_1"]["kernel"],
                 t["Dense_1"]["bias"], head["kernel"], head["bias"]]
        return np.concatenate([np.asarray(x, np.float32).ravel() for x in parts])


class FeedForwardActor(_FeedForward):
    head_scale = 0.01
    head_path = ["action_head", "Dense_0"]

    def __init__(self, torso: MLPTorso, action_head: DiscreteActionHead):
        self.torso, self.action_head = torso, action_head
        self.out_dim = action_head.action_dim


class FeedForwardValueNet(_FeedForward):
    head_scale = 1.0
    head_path = ["Dense_0"]
    out_dim = 1

    def __init__(self, torso: MLPTorso, centralised_critic: bool = False):
        self.torso, self.centralised_critic = torso, centralised_critic
