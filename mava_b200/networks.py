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


# ---------------------------------------------------------------------------------------------
# recurrent networks (mava/networks.py:238-331)
# ---------------------------------------------------------------------------------------------
def lecun_normal(rng: np.random.Generator, shape: Sequence[int]) -> np.ndarray:
    """flax default_kernel_init = variance_scaling(1.0, "fan_in", "truncated_normal")."""
    fan_in = int(shape[0])
    std = np.sqrt(1.0 / fan_in) / 0.87962566103423978  # stddev of a unit normal truncated at +-2
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


class ScannedRNN:
    """networks.py:238-266: reset-masked flax GRUCell scanned over the leading (time) axis; the
    scan runs in csrc/rnn_f32.cu."""

    def __init__(self, hidden_state_dim: int = 128):
        self.hidden_state_dim = int(hidden_state_dim)

    @staticmethod
    def initialize_carry(batch_size: Sequence[int], hidden_size: int, device=None):
        import torch

        return torch.zeros(*batch_size, hidden_size, device=device)


class _Recurrent:
    """Flat-parameter handling of RecurrentActor / RecurrentValueNet (layout: include/mava_b200.h,
    ``mava_rnn_desc``)."""

    pre_torso: MLPTorso
    post_torso: MLPTorso
    hidden_state_dim: int
    out_dim: int
    head_scale: float
    head_path: List[str]

    def _dims(self):
        if len(self.pre_torso.layer_sizes) != 1 or len(self.post_torso.layer_sizes) != 1:
            raise NotImplementedError(
                "recurrent networks with one pre-torso and one post-torso layer are supported "
                "(configs/network/rnn.yaml)")
        H, Q = self.pre_torso.layer_sizes[0], self.post_torso.layer_sizes[0]
        if H != self.hidden_state_dim:
            raise ValueError("hidden_state_dim must equal the pre_torso width: flax's GRUCell "
                             "takes features = ins.shape[-1] (networks.py:258)")
        return H, Q

    def desc(self, num_agents: int, view_dim: int, add_agent_id: bool, input_mode: int,
             dense_in_dim: int = 0, rows_per_env=None, precision: int = 0):
        H, Q = self._dims()
        return native.rnn_desc(input_mode, add_agent_id, num_agents, view_dim, H, Q, self.out_dim,
                               dense_in_dim, rows_per_env, precision)

    def shapes(self, in_dim: int):
        H, Q = self._dims()
        return [("pre_w", (in_dim, H)), ("pre_b", (H,)), ("wi", (H, 3 * H)), ("bi", (3 * H,)),
                ("wh", (H, 3 * H)), ("hn_b", (H,)), ("post_w", (H, Q)), ("post_b", (Q,)),
                ("head_w", (Q, self.out_dim)), ("head_b", (self.out_dim,))]

    def init(self, key: np.ndarray, in_dim: int) -> np.ndarray:
        """orthogonal(sqrt 2) torsos (networks.py:54), flax GRUCell defaults (lecun-normal input
        kernels, orthogonal recurrent kernels, zero biases), orthogonal(head_scale) head."""
        H, _ = self._dims()
        rng = np.random.default_rng([int(key[0]), int(key[1])])
        out = []
        for name, s in self.shapes(in_dim):
            if len(s) == 1:
                out.append(np.zeros(s, np.float32))
            elif name == "wi":
                out.append(np.concatenate([lecun_normal(rng, (H, H)) for _ in range(3)], 1))
            elif name == "wh":
                out.append(np.concatenate([orthogonal(rng, (H, H), 1.0) for _ in range(3)], 1))
            else:
                out.append(orthogonal(rng, s, self.head_scale if name == "head_w"
                                      else float(np.sqrt(2))))
        return np.concatenate([p.ravel() for p in out])

    def _split(self, flat: np.ndarray, in_dim: int) -> Dict[str, np.ndarray]:
        parts, off = {}, 0
        for name, s in self.shapes(in_dim):
            n = int(np.prod(s))
            parts[name] = np.asarray(flat[off:off + n]).reshape(s)
            off += n
        return parts

    def to_flax_tree(self, flat: np.ndarray, in_dim: int) -> Dict[str, Any]:
        H, _ = self._dims()
        p = self._split(flat, in_dim)
        cell = {}
        for g, name in enumerate(("r", "z", "n")):
            cell["i" + name] = {"kernel": p["wi"][:, g * H:(g + 1) * H],
                                "bias": p["bi"][g * H:(g + 1) * H]}
            cell["h" + name] = {"kernel": p["wh"][:, g * H:(g + 1) * H]}
        cell["hn"]["bias"] = p["hn_b"]
        tree: Dict[str, Any] = {
            "pre_torso": {"Dense_0": {"kernel": p["pre_w"], "bias": p["pre_b"]}},
            "ScannedRNN_0": {"GRUCell_0": cell},
            "post_torso": {"Dense_0": {"kernel": p["post_w"], "bias": p["post_b"]}}}
        node = tree
        for k in self.head_path[:-1]:
            node = node.setdefault(k, {})
        node[self.head_path[-1]] = {"kernel": p["head_w"], "bias": p["head_b"]}
        return {"params": tree}

    def from_flax_tree(self, tree: Dict[str, Any]) -> np.ndarray:
        p = tree["params"]
        cell = p["ScannedRNN_0"]["GRUCell_0"]
        head = p
        for k in self.head_path:
            head = head[k]
        parts = [p["pre_torso"]["Dense_0"]["kernel"], p["pre_torso"]["Dense_0"]["bias"],
                 np.concatenate([cell["i" + g]["kernel"] for g in "rzn"], 1),
                 np.concatenate([cell["i" + g]["bias"] for g in "rzn"]),
                 np.concatenate([cell["h" + g]["kernel"] for g in "rzn"], 1), cell["hn"]["bias"],
                 p["post_torso"]["Dense_0"]["kernel"], p["post_torso"]["Dense_0"]["bias"],
                 head["kernel"], head["bias"]]
        return np.concatenate([np.asarray(x, np.float32).ravel() for x in parts])


class RecurrentActor(_Recurrent):
    """networks.py:269-294."""

    head_scale = 0.01
    head_path = ["action_head", "Dense_0"]

    def __init__(self, pre_torso: MLPTorso, post_torso: MLPTorso, action_head: DiscreteActionHead,
                 hidden_state_dim: int = 128):
        self.pre_torso, self.post_torso, self.action_head = pre_torso, post_torso, action_head
        self.hidden_state_dim = int(hidden_state_dim)
        self.out_dim = action_head.action_dim


class RecurrentValueNet(_Recurrent):
    """networks.py:297-331."""

    head_scale = 1.0
    head_path = ["Dense_0"]
    out_dim = 1

    def __init__(self, pre_torso: MLPTorso, post_torso: MLPTorso, centralised_critic: bool = False,
                 hidden_state_dim: int = 128):
        self.pre_torso, self.post_torso = pre_torso, post_torso
        self.centralised_critic = centralised_critic
        self.hidden_state_dim = int(hidden_state_dim)
