"""Console / JSON-lines logging with the reference's interface (mava/utils/logger.py:44-105).

Only what the hot path's run loop needs: the Neptune / TensorBoard / marl-eval back ends of the
reference are out of scope."""
from __future__ import annotations

import json
import os
import time
from enum import Enum
from typing import Dict

import numpy as np
import torch


class LogEvent(Enum):
    ACT = "actor"
    TRAIN = "trainer"
    EVAL = "evaluator"
    ABSOLUTE = "absolute"
    MISC = "misc"


def _to_np(v) -> np.ndarray:
    if isinstance(v, torch.Tensor):
        return v.detach().float().cpu().numpy()
    return np.asarray(v, dtype=np.float64)


def describe(x: np.ndarray) -> Dict[str, float]:
    if x.size <= 1:
        return {"mean": float(x.mean()) if x.size else 0.0}
    return {"mean": float(x.mean()), "std": float(x.std()), "min": float(x.min()),
            "max": float(x.max())}


class MavaLogger:
    def __init__(self, config, rank: int = 0):
        self.cfg = config.logger
        self.rank = rank
        self.log_win_rate = bool(config.env.get("log_win_rate", False))
        self._json = None
        if rank == 0 and self.cfg.use_json:
            path = self.cfg.kwargs.json_path or os.path.join(
                self.cfg.base_exp_path, "json", f"{self.cfg.system_name}_{int(time.time())}")
            os.makedirs(path, exist_ok=True)
            self._json = open(os.path.join(path, "metrics.jsonl"), "a")

    def log(self, metrics: Dict, t: int, t_eval: int, event: LogEvent) -> None:
        if self.rank != 0:
            return
        metrics = {k: _to_np(v) for k, v in metrics.items()}
        if "won_episode" in metrics and self.log_win_rate:
            won = metrics.pop("won_episode")
            metrics["win_rate"] = np.asarray(100.0 * won.sum() / max(1, won.size))
        if event == LogEvent.TRAIN:  # train metrics are averaged (logger.py:72-74)
            flat = {k: float(v.mean()) for k, v in metrics.items()}
        else:
            flat = {}
            for k, v in metrics.items():
                for stat, val in describe(v).items():
                    flat[f"{k}/{stat}" if v.size > 1 else k] = val
        self._emit(flat, t, t_eval, event)

    def _emit(self, flat: Dict[str, float], t: int, t_eval: int, event: LogEvent) -> None:
        if self.cfg.use_console:
            body = " | ".join(f"{k.replace('_', ' ').capitalize()}: {v:.3f}" for k, v in
                              flat.items() if k.endswith("mean") or "/" not in k)
            print(f"{event.value.upper()} - {body}", flush=True)
        if self._json is not None:
            self._json.write(json.dumps({"event": event.value, "t": t, "eval": t_eval, **flat}) + "\n")
            self._json.flush()

    def log_summary(self, summary: Dict[str, Dict[str, float]], scalars: Dict[str, float], t: int,
                    t_eval: int, event: LogEvent) -> None:
        """Log metrics that arrive already described ({name: {mean, std, min, max}}, the device-side
        reduction of the finished episodes) next to plain scalars; same keys as ``log``."""
        if self.rank != 0:
            return
        flat = {}
        for k, v in summary.items():
            if isinstance(v, dict):
                for stat, val in v.items():
                    flat[f"{k}/{stat}"] = float(val)
        flat.update({k: float(v) for k, v in scalars.items()})
        self._emit(flat, t, t_eval, event)

    def stop(self) -> None:
        if self._json is not None:
            self._json.close()


def get_final_step_metrics(metrics: Dict[str, torch.Tensor]):
    """mava/wrappers/episode_metrics.py:114-132: keep the metrics of finished episodes only."""
    metrics = dict(metrics)
    is_final = metrics.pop("is_terminal_step")
    has_final = bool(is_final.any().item())
    if not has_final:
        return {k: torch.zeros_like(v) for k, v in metrics.items()}, False
    return {k: v[is_final] for k, v in metrics.items()}, True
