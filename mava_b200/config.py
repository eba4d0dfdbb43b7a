"""A small Hydra-compatible config composer for the Mava config surface.

The reference is driven by Hydra 1.3 (``@hydra.main(config_path="../../configs", config_name=
"default_ff_mappo.yaml")``, mava/systems/ppo/ff_mappo.py:556-565) over the YAML tree in
``mava/configs``.  Hydra/OmegaConf are not part of this image, so this module implements the
subset of Hydra that tree uses:

* ``defaults`` lists with ``group: option`` entries, sibling includes (``- base_logger``) and
  ``_self_`` (appended last when absent, as Hydra >= 1.1 does);
* nested groups (``env/rware.yaml`` selecting ``scenario: tiny-2ag``), packaged under the group
  path (``env.scenario``);
* command line overrides: ``env=lbf``, ``env/scenario=tiny-4ag``, ``arch.num_envs=2048``,
  ``+new.key=1``, ``~key``; values are YAML (``~``/``null`` -> None).

Configs come either from the built-in tree below (the hot-path subset of mava/configs, same keys
and default values) or, with ``config_dir=...``, verbatim from a directory of YAML files such as the
reference's own ``mava/configs``.
"""
from __future__ import annotations

import copy
import sys
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional

import yaml


class Config(dict):
    """dict with attribute access, the slice of OmegaConf's DictConfig the systems rely on."""

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = _wrap(v)

    def __delattr__(self, k: str) -> None:
        del self[k]

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_container(self) -> Dict[str, Any]:
        return {k: (v.to_container() if isinstance(v, Config) else copy.deepcopy(v))
                for k, v in self.items()}


def _wrap(v: Any) -> Any:
    if isinstance(v, Config):
        return v
    if isinstance(v, dict):
        return Config({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    return v


def _merge(dst: Dict, src: Dict) -> Dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


# ---------------------------------------------------------------------------------------------
# built-in tree: hot-path subset of mava/configs (keys and defaults as in the reference files)
# ---------------------------------------------------------------------------------------------
def _ppo_system(recurrent: bool) -> Dict:
    d = dict(total_timesteps=None, num_updates=1000, seed=42, add_agent_id=True, actor_lr=2.5e-4,
             critic_lr=2.5e-4, update_batch_size=2, rollout_length=128, ppo_epochs=4,
             num_minibatches=2, gamma=0.99, gae_lambda=0.95, clip_eps=0.2, ent_coef=0.01,
             vf_coef=0.5, max_grad_norm=0.5, decay_learning_rates=False)
    if recurrent:
        d["recurrent_chunk_size"] = None
    return d


def _torso(sizes: List[int]) -> Dict:
    return {"_target_": "mava.networks.MLPTorso", "layer_sizes": list(sizes),
            "use_layer_norm": False, "activation": "relu"}


def _rware_scenario(task: str, agents: int, queue: int, rows: int = 1) -> Dict:
    return dict(name="RobotWarehouse-v0", task_name=task,
                task_config=dict(column_height=8, shelf_rows=rows, shelf_columns=3,
                                 num_agents=agents, sensor_range=1, request_queue_size=queue),
                env_kwargs={})


def _lbf_scenario(task: str, grid: int, fov: int, agents: int, food: int, coop: bool) -> Dict:
    return dict(name="LevelBasedForaging-v0", task_name=task,
                task_config=dict(grid_size=grid, fov=fov, num_agents=agents, num_food=food,
                                 max_agent_level=2, force_coop=coop),
                env_kwargs={})


def _builtin_tree() -> Dict[str, Dict]:
    t: Dict[str, Dict] = {}
    for sysname, net in (("ff_ippo", "mlp"), ("ff_mappo", "mlp"), ("rec_ippo", "rnn"),
                         ("rec_mappo", "rnn")):
        t[f"default_{sysname}"] = {"defaults": [{"logger": sysname}, {"arch": "anakin"},
                                                {"system": f"ppo/{sysname}"}, {"network": net},
                                                {"env": "rware"}, "_self_"]}
        t[f"logger/{sysname}"] = {"defaults": ["base_logger"], "system_name": sysname}
        t[f"system/ppo/{sysname}"] = _ppo_system(sysname.startswith("rec"))
    t["logger/base_logger"] = dict(
        base_exp_path="results", use_console=True, use_tb=False, use_json=False, use_neptune=False,
        kwargs=dict(neptune_project="Instadeep/Mava", neptune_tag=["rware"],
                    detailed_neptune_logging=False, json_path=None, upload_json_data=False),
        checkpointing=dict(save_model=False,
                           save_args=dict(save_interval_steps=1, max_to_keep=1, keep_period=None,
                                          checkpoint_uid=None),
                           load_model=False, load_args=dict(checkpoint_uid="")))
    t["arch/anakin"] = dict(num_envs=16, evaluation_greedy=False, num_evaluation=200,
                            num_eval_episodes=32, num_absolute_metric_eval_episodes=320,
                            absolute_metric=True)
    head = {"_target_": "mava.networks.DiscreteActionHead"}
    t["network/mlp"] = dict(actor_network=dict(pre_torso=_torso([128, 128])), action_head=head,
                            critic_network=dict(pre_torso=_torso([128, 128])))
    t["network/rnn"] = dict(hidden_state_dim=128,
                            actor_network=dict(pre_torso=_torso([128]), post_torso=_torso([128])),
                            action_head=dict(head),
                            critic_network=dict(pre_torso=_torso([128]), post_torso=_torso([128])),
                            q_network=dict(pre_torso=_torso([128]), post_torso=_torso([128])))
    t["env/rware"] = {"defaults": ["_self_", {"scenario": "tiny-2ag"}],
                      "env_name": "RobotWarehouse", "eval_metric": "episode_return",
                      "implicit_agent_id": False, "log_win_rate": False,
                      "kwargs": {"time_limit": 500}}
    t["env/lbf"] = {"defaults": ["_self_", {"scenario": "2s-8x8-2p-2f-coop"}],
                    "env_name": "LevelBasedForaging", "use_individual_rewards": False,
                    "eval_metric": "episode_return", "implicit_agent_id": False,
                    "log_win_rate": False, "kwargs": {"time_limit": 100}}
    t["env/smax"] = {"defaults": ["_self_"], "env_name": "Smax",
                     "scenario": {"name": "HeuristicEnemySMAX", "task_name": "2s3z"},
                     "eval_metric": "win_rate", "implicit_agent_id": False, "log_win_rate": True,
                     "kwargs": {"see_enemy_actions": True, "walls_cause_death": True,
                                "attack_mode": "closest"}}
    # SMAX-shaped synthetic step source (benchmark only, see wrappers/synthetic.py); 3s5z: 8 allies
    # vs 8 enemies -> 5 + 8 actions; obs / state sizes are parameters (jaxmarl is not available)
    t["env/smax_synthetic"] = {"defaults": ["_self_", {"scenario": "synthetic-3s5z"}],
                               "env_name": "SmaxSynthetic", "eval_metric": "episode_return",
                               "implicit_agent_id": False, "log_win_rate": False,
                               "kwargs": {"time_limit": 100}}
    t["env/scenario/synthetic-3s5z"] = dict(
        name="SmaxSynthetic", task_name="synthetic-3s5z",
        task_config=dict(num_agents=8, obs_dim=205, state_dim=168, num_actions=13), env_kwargs={})
    t["env/scenario/tiny-2ag"] = _rware_scenario("tiny-2ag", 2, 2)
    t["env/scenario/tiny-4ag"] = _rware_scenario("tiny-4ag", 4, 4)
    t["env/scenario/tiny-4ag-easy"] = _rware_scenario("tiny-4ag-easy", 4, 8)
    t["env/scenario/small-4ag"] = _rware_scenario("small-4ag", 4, 4, rows=2)
    t["env/scenario/2s-8x8-2p-2f-coop"] = _lbf_scenario("2s-8x8-2p-2f-coop", 8, 2, 2, 2, True)
    t["env/scenario/8x8-2p-2f-coop"] = _lbf_scenario("8x8-2p-2f-coop", 8, 8, 2, 2, True)
    t["env/scenario/2s-10x10-3p-3f"] = _lbf_scenario("2s-10x10-3p-3f", 10, 2, 3, 3, False)
    t["env/scenario/10x10-3p-3f"] = _lbf_scenario("10x10-3p-3f", 10, 10, 3, 3, False)
    t["env/scenario/15x15-3p-5f"] = _lbf_scenario("15x15-3p-5f", 15, 15, 3, 5, False)
    t["env/scenario/15x15-4p-3f"] = _lbf_scenario("15x15-4p-3f", 15, 15, 4, 3, False)
    t["env/scenario/15x15-4p-5f"] = _lbf_scenario("15x15-4p-5f", 15, 15, 4, 5, False)
    return t


class _Store:
    def __init__(self, config_dir: Optional[str]):
        self.dir = Path(config_dir) if config_dir else None
        self.tree = None if self.dir else _builtin_tree()

    def load(self, path: str) -> Dict:
        path = path[:-5] if path.endswith(".yaml") else path
        if self.dir is not None:
            f = self.dir / (path + ".yaml")
            if not f.exists():
                raise FileNotFoundError(f"config '{path}' not found under {self.dir}")
            return yaml.safe_load(f.read_text()) or {}
        if path not in self.tree:
            raise FileNotFoundError(
                f"config '{path}' is not in the built-in tree; options: "
                f"{sorted(k for k in self.tree if k.rsplit('/', 1)[0] == path.rsplit('/', 1)[0])}")
        return copy.deepcopy(self.tree[path])


def _compose_node(store: _Store, path: str, group: str, choices: Dict[str, str]) -> Dict:
    """Compose the config file ``path`` whose config group is ``group`` ('' for the root)."""
    node = store.load(path)
    defaults = node.pop("defaults", None) or []
    if "_self_" not in defaults:
        defaults = list(defaults) + ["_self_"]
    out: Dict = {}
    for d in defaults:
        if d == "_self_":
            _merge(out, node)
        elif isinstance(d, str):  # sibling file in the same group, same package
            sib = f"{group}/{d}" if group else d
            _merge(out, _compose_node(store, sib, group, choices))
        elif isinstance(d, dict):
            (sub, option), = d.items()
            sub_group = f"{group}/{sub}" if group else sub
            option = choices.get(sub_group, option)
            if option is None:
                continue
            child = _compose_node(store, f"{sub_group}/{option}", sub_group, choices)
            _merge(out.setdefault(sub, {}), child)
        else:
            raise ValueError(f"unsupported defaults entry {d!r} in {path}")
    return out


def _set_dotted(cfg: Dict, dotted: str, value: Any, must_exist: bool) -> None:
    keys = dotted.split(".")
    cur = cfg
    for k in keys[:-1]:
        if k not in cur or not isinstance(cur[k], dict):
            if must_exist:
                raise KeyError(f"override '{dotted}': key '{k}' not in config (use +{dotted}=...)")
            cur[k] = {}
        cur = cur[k]
    if must_exist and keys[-1] not in cur:
        raise KeyError(f"override '{dotted}': no such key (use +{dotted}=... to add it)")
    cur[keys[-1]] = value


def compose(config_name: str, overrides: Iterable[str] = (), config_dir: Optional[str] = None
            ) -> Config:
    """Hydra-style composition of ``config_name`` (e.g. ``default_ff_mappo.yaml``)."""
    store = _Store(config_dir)
    choices: Dict[str, str] = {}
    value_overrides: List[tuple] = []
    for ov in overrides:
        if ov.startswith("~"):
            value_overrides.append(("del", ov[1:].split("=")[0], None))
            continue
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        key, raw = ov.split("=", 1)
        add = key.startswith("+")
        key = key.lstrip("+")
        value = yaml.safe_load(raw) if raw != "" else ""
        if "." not in key and (("/" in key) or _is_group(store, key)):
            choices[key] = value  # config group selection
        else:
            value_overrides.append(("add" if add else "set", key, value))
    cfg = _compose_node(store, config_name, "", choices)
    for kind, key, value in value_overrides:
        if kind == "del":
            cur = cfg
            ks = key.split(".")
            for k in ks[:-1]:
                cur = cur[k]
            cur.pop(ks[-1], None)
        else:
            _set_dotted(cfg, key, value, must_exist=(kind == "set"))
    return _wrap(cfg)


def _is_group(store: _Store, key: str) -> bool:
    if store.dir is not None:
        return (store.dir / key).is_dir()
    return any(p.startswith(key + "/") for p in store.tree)


def parse_cli(argv: Optional[List[str]] = None):
    """Split ``--config-dir DIR`` from Hydra-style overrides."""
    argv = list(sys.argv[1:] if argv is None else argv)
    config_dir = None
    out = []
    i = 0
    while i < len(argv):
        if argv[i] in ("--config-dir", "-cd"):
            config_dir = argv[i + 1]
            i += 2
        elif argv[i].startswith("--config-dir="):
            config_dir = argv[i].split("=", 1)[1]
            i += 1
        else:
            out.append(argv[i])
            i += 1
    return out, config_dir
