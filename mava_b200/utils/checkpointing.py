"""Parameter export / import in the reference's pytree naming.

The reference checkpoints with Orbax (mava/utils/checkpointing.py:116-207: a ``learner_state`` pytree
whose ``params`` entry is ``Params(actor_params, critic_params)`` of flax dicts).  Orbax, flax and jax
are not installable in this image, so this module writes and reads the same *tree* --
``{"learner_state": {"params": {"actor_params": {...}, "critic_params": {...}}}}`` with flax's
``{"params": {"torso": {"Dense_k": {"kernel", "bias"}}, ...}}`` naming (mava_b200.networks
``to_flax_tree`` / ``from_flax_tree``) -- in two self-describing containers:

* ``.npz``: one array per leaf, keys are '/'-joined tree paths;
* ``.msgpack``: flax.serialization's wire format [RECALL: msgpack map of maps, ndarray leaves as
  ExtType(1, packb((shape, dtype.name, bytes)))], which ``flax.serialization.msgpack_restore``
  reads on a machine that has flax.

A Mava user can therefore move weights either way with a few lines.  Optimiser
moments use optax's names (``mu`` / ``nu`` / ``count``, ScaleByAdamState).
"""
from __future__ import annotations

from typing import Any, Dict

import msgpack
import numpy as np


def flatten_tree(tree: Dict[str, Any], prefix: str = "") -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for k, v in tree.items():
        path = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten_tree(v, path))
        else:
            out[path] = np.asarray(v)
    return out


def unflatten_tree(flat: Dict[str, np.ndarray]) -> Dict[str, Any]:
    tree: Dict[str, Any] = {}
    for path, v in flat.items():
        node = tree
        keys = path.split("/")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = np.asarray(v)
    return tree


def _pack_leaf(obj: Any) -> Any:
    if isinstance(obj, np.ndarray):
        payload = msgpack.packb((list(obj.shape), obj.dtype.name, obj.tobytes()), use_bin_type=True)
        return msgpack.ExtType(1, payload)
    if isinstance(obj, np.generic):
        return _pack_leaf(np.asarray(obj))
    raise TypeError(f"cannot serialise {type(obj)}")


def _unpack_ext(code: int, data: bytes) -> Any:
    if code == 1:
        shape, dtype, buf = msgpack.unpackb(data, raw=False)
        return np.frombuffer(buf, dtype=np.dtype(dtype)).reshape(shape).copy()
    return msgpack.ExtType(code, data)


def to_msgpack(tree: Dict[str, Any]) -> bytes:
    return msgpack.packb(tree, default=_pack_leaf, use_bin_type=True, strict_types=True)


def from_msgpack(blob: bytes) -> Dict[str, Any]:
    return msgpack.unpackb(blob, ext_hook=_unpack_ext, raw=False)


def learner_tree(actor_network, critic_network, actor_flat: np.ndarray, critic_flat: np.ndarray,
                 actor_in_dim: int, critic_in_dim: int, opt: Dict[str, Any] | None = None
                 ) -> Dict[str, Any]:
    """The reference's checkpoint tree for one learner (replica / device axes dropped: the reference
    saves the unreplicated state, ff_mappo.py:520-528)."""
    tree: Dict[str, Any] = {"learner_state": {"params": {
        "actor_params": actor_network.to_flax_tree(np.asarray(actor_flat), actor_in_dim),
        "critic_params": critic_network.to_flax_tree(np.asarray(critic_flat), critic_in_dim)}}}
    if opt is not None:
        tree["learner_state"]["opt_states"] = opt
    return tree


def save(path: str, tree: Dict[str, Any]) -> None:
    if path.endswith(".npz"):
        np.savez(path, **flatten_tree(tree))
    elif path.endswith(".msgpack"):
        with open(path, "wb") as f:
            f.write(to_msgpack(tree))
    else:
        raise ValueError("checkpoint path must end in .npz or .msgpack")


def load(path: str) -> Dict[str, Any]:
    if path.endswith(".npz"):
        with np.load(path) as z:
            return unflatten_tree({k: z[k] for k in z.files})
    if path.endswith(".msgpack"):
        with open(path, "rb") as f:
            return from_msgpack(f.read())
    raise ValueError("checkpoint path must end in .npz or .msgpack")


def restore_params(tree: Dict[str, Any], actor_network, critic_network):
    """(actor_flat, critic_flat) float32 vectors in this library's layout from a checkpoint tree
    (the counterpart of Checkpointer.restore_params, mava/utils/checkpointing.py:150-207)."""
    p = tree["learner_state"]["params"]
    return (actor_network.from_flax_tree(p["actor_params"]),
            critic_network.from_flax_tree(p["critic_params"]))


CHECKPOINTER_VERSION = 1.0


class Checkpointer:
    """The reference's ``Checkpointer`` surface (mava/utils/checkpointing.py:36-207: same constructor
    arguments, ``save`` and ``restore_params``) on the tree containers above.  Directory layout
    ``<cwd>/<rel_dir>/<model_name>/<checkpoint_uid>/<timestep>/checkpoint.msgpack`` like the
    reference's Orbax manager; the files hold the same *tree* but are NOT Orbax directories
    (orbax / jax are not installable here)."""

    def __init__(self, model_name: str, metadata: Any = None, rel_dir: str = "checkpoints",
                 checkpoint_uid: str | None = None, save_interval_steps: int = 1,
                 max_to_keep: int | None = 1, keep_period: int | None = None):
        import os
        from datetime import datetime

        uid = checkpoint_uid if checkpoint_uid else datetime.now().strftime("%Y%m%d%H%M%S")
        self.directory = os.path.join(os.getcwd(), rel_dir, model_name, uid)
        self.save_interval_steps = max(1, int(save_interval_steps))
        self.max_to_keep, self.keep_period = max_to_keep, keep_period
        self.metadata = {"checkpointer_version": CHECKPOINTER_VERSION}
        if metadata is not None:
            self.metadata["config"] = _json_ready(metadata)
        self._saved: Dict[int, float] = {}  # timestep -> episode_return (best_fn of the reference)
        self._n_calls = 0

    def _path(self, timestep: int) -> str:
        import os

        return os.path.join(self.directory, str(int(timestep)), "checkpoint.msgpack")

    def save(self, timestep: int, unreplicated_learner_state: Dict[str, Any],
             episode_return: float = 0.0) -> bool:
        """``unreplicated_learner_state`` is a checkpoint tree (``learner_tree``)."""
        import json
        import os
        import shutil

        self._n_calls += 1
        if (self._n_calls - 1) % self.save_interval_steps != 0:
            return False
        path = self._path(timestep)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        save(path, unreplicated_learner_state)
        with open(os.path.join(self.directory, "metadata.json"), "w") as f:
            json.dump(self.metadata, f)
        self._saved[int(timestep)] = float(episode_return)
        if self.max_to_keep is not None:
            keep_always = {t for t in self._saved if self.keep_period and t % self.keep_period == 0}
            ranked = sorted((t for t in self._saved if t not in keep_always),
                            key=lambda t: (self._saved[t], t), reverse=True)
            for t in ranked[self.max_to_keep:]:
                shutil.rmtree(os.path.dirname(self._path(t)), ignore_errors=True)
                del self._saved[t]
        return True

    def restore_tree(self, timestep: int | None = None) -> Dict[str, Any]:
        import os

        if timestep is None:
            steps = [int(d) for d in os.listdir(self.directory) if d.isdigit()]
            if not steps:
                raise FileNotFoundError(f"no checkpoint under {self.directory}")
            timestep = max(steps)
        return load(self._path(timestep))

    def restore_params(self, actor_network, critic_network, timestep: int | None = None):
        """(actor_flat, critic_flat) of the latest (or the given) checkpoint."""
        return restore_params(self.restore_tree(timestep), actor_network, critic_network)


def _json_ready(obj: Any) -> Any:
    if isinstance(obj, dict) or hasattr(obj, "items"):
        return {str(k): _json_ready(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_json_ready(v) for v in obj]
    if isinstance(obj, (bool, str, int, float, type(None))):
        return obj
    return str(obj)
