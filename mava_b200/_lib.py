"""ctypes binding of ``libmava_b200.so`` (include/mava_b200.h).  No fallback: if the library is
missing or a symbol is absent, importing callers fail loudly."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmava_b200.so"

c_void = C.c_void_p
c_int = C.c_int
c_i64 = C.c_int64
c_f32 = C.c_float


class RwareConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "column_height", "shelf_rows", "shelf_columns", "num_agents", "sensor_range",
        "request_queue_size", "time_limit")]


class LbfConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "grid_size", "fov", "num_agents", "num_food", "max_agent_level", "force_coop",
        "time_limit", "use_individual_rewards")]


class EnvDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "num_agents", "view_dim", "num_actions", "state_stride", "time_limit", "grid_h",
        "grid_w", "aux0", "aux1", "algo_bytes_per_step")]


class MlpDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "input_mode", "add_agent_id", "num_agents", "view_dim", "in_dim", "h1", "h2", "out_dim")]


class RnnDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "input_mode", "add_agent_id", "num_agents", "view_dim", "in_dim", "rows_per_env", "hidden",
        "post", "out_dim", "precision")]


class SynthConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_agents", "obs_dim", "state_dim", "num_actions")] + [
        ("done_prob", C.c_float), ("reward_std", C.c_float)]


class PeerGroup(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("buf", C.c_void_p * 8)]


class PpoHyper(C.Structure):
    _fields_ = [("clip_eps", c_f32), ("ent_coef", c_f32), ("vf_coef", c_f32)]


P = C.POINTER
# name -> (restype, argtypes); every symbol declared in include/mava_b200.h
SIGNATURES = {
    "mava_abi_version": (c_int, []),
    "mava_error_string": (C.c_char_p, [c_int]),
    "mava_device_info": (c_int, [P(c_int)]),
    "mava_prng_split_chain": (c_int, [c_void, c_void, c_int, c_void]),
    "mava_prng_split": (c_int, [c_void, c_void, c_int, c_void]),
    "mava_prng_random_bits": (c_int, [c_void, c_void, c_i64, c_void]),
    "mava_sort_workspace_bytes": (c_i64, [c_i64]),
    "mava_sort_by_key": (c_int, [c_void, c_void, c_void, c_i64, c_void, c_void, c_void]),
    "mava_env_create": (c_int, [c_int, c_void, C.c_size_t, P(c_void)]),
    "mava_env_destroy": (c_int, [c_void]),
    "mava_env_dims_of": (c_int, [c_void, P(EnvDims)]),
    "mava_env_reset": (c_int, [c_void, c_void, c_void, c_void, c_void, c_int, c_void]),
    "mava_env_step": (c_int, [c_void] * 9 + [c_int, c_int, c_void]),
    "mava_env_peek": (c_int, [c_void, c_void, c_int, c_void, c_int, c_void]),
    "mava_mlp_param_count": (c_i64, [P(MlpDesc)]),
    "mava_ff_act": (c_int, [P(MlpDesc), c_void, P(MlpDesc), c_void, c_void, c_void, c_void,
                            c_int, c_int, c_int, c_void, c_void, c_void, c_void, c_void]),
    "mava_ff_value": (c_int, [P(MlpDesc), c_void, c_void, c_int, c_void, c_void]),
    "mava_gae": (c_int, [c_void] * 5 + [c_f32, c_f32, c_int, c_int, c_int, c_int, c_void, c_void,
                                        c_void]),
    "mava_episode_stats": (c_int, [c_void, c_void, c_void, c_i64, c_int, c_void, c_void]),
    "mava_ppo_minibatch_rows": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_void]),
    "mava_ppo_workspace_bytes": (c_i64, [P(MlpDesc), P(MlpDesc), c_int]),
    "mava_ppo_loss_grad": (c_int, [P(MlpDesc), c_void, P(MlpDesc), c_void, P(PpoHyper)] +
                           [c_void] * 8 + [c_int, c_int, c_void, c_void, c_void]),
    "mava_clip_adam_pair": (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_f32,
                                    c_f32, c_f32, c_f32, c_int, c_int, c_void]),
    "mava_clip_adam_pair_pack": (c_int, [c_void, c_void, c_void, c_void, c_void, P(MlpDesc), c_void,
                                         P(MlpDesc), c_void, c_f32, c_f32, c_f32, c_f32, c_int,
                                         c_int, c_void]),
    "mava_mlp_pack_bytes": (c_i64, [P(MlpDesc)]),
    "mava_mlp_pack_bf16": (c_int, [P(MlpDesc), c_void, c_void, c_void]),
    "mava_ff_act_bf16": (c_int, [P(MlpDesc), c_void, c_void, P(MlpDesc), c_void, c_void, c_void,
                                 c_void, c_void, c_int, c_int, c_int, c_void, c_void, c_void,
                                 c_void, c_void]),
    "mava_ppo_workspace_bytes_bf16": (c_i64, [P(MlpDesc), P(MlpDesc), c_int]),
    "mava_ppo_loss_grad_bf16": (c_int, [P(MlpDesc), c_void, c_void, P(MlpDesc), c_void, c_void,
                                        P(PpoHyper)] + [c_void] * 8 +
                                [c_int, c_int, c_void, c_void, c_void]),
    "mava_ppo_adv_stats": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void, c_void]),
    "mava_ppo_loss_grad_bf16_stats": (c_int, [P(MlpDesc), c_void, c_void, P(MlpDesc), c_void, c_void,
                                              P(PpoHyper)] + [c_void] * 8 +
                                      [c_int, c_int, c_void, c_void, c_void, c_void]),
    "mava_tc_selftest": (c_int, [c_int, c_void, c_void, c_void, c_int, c_int, c_void]),
    "mava_ff_rollout_bf16": (c_int, [c_void, P(MlpDesc)] + [c_void] * 6 + [c_int] * 3 +
                             [c_void] * 7),
    "mava_ff_rollout_bf16_ex": (c_int, [c_void, P(MlpDesc)] + [c_void] * 6 + [c_int] * 6 +
                                [c_void] * 7),
    "mava_episode_first_terminal": (c_int, [c_void, c_void, c_void, c_int, c_int, c_void, c_void,
                                            c_void]),
    "mava_synth_reset": (c_int, [P(SynthConfig)] + [c_void] * 5 + [c_int, c_void]),
    "mava_synth_step": (c_int, [P(SynthConfig)] + [c_void] * 10 + [c_int, c_void]),
    "mava_gemm": (c_int, [c_int, c_void, c_int, c_i64, c_void, c_int, c_i64, c_void, c_i64, c_int,
                          c_int, c_int, c_void, c_int, c_void, c_i64, c_int, c_int, c_void]),
    "mava_rnn_param_count": (c_i64, [P(RnnDesc)]),
    "mava_rec_act_workspace_bytes": (c_i64, [P(RnnDesc), P(RnnDesc), c_int]),
    "mava_rec_act": (c_int, [P(RnnDesc), c_void, P(RnnDesc), c_void] + [c_void] * 10 +
                     [c_int, c_int, c_int] + [c_void] * 6),
    "mava_rec_ppo_workspace_bytes": (c_i64, [P(RnnDesc), P(RnnDesc), c_int, c_int]),
    "mava_rec_ppo_loss_grad": (c_int, [P(RnnDesc), c_void, P(RnnDesc), c_void, P(PpoHyper)] +
                               [c_void] * 13 + [c_int] * 5 + [c_void] * 3),
    "mava_peer_buffer_bytes": (c_i64, [c_i64]),
    "mava_peer_alloc": (c_int, [c_i64, P(c_void), c_void]),
    "mava_peer_open": (c_int, [c_void, P(c_void)]),
    "mava_peer_close": (c_int, [c_void]),
    "mava_peer_free": (c_int, [c_void]),
    "mava_peer_status": (c_int, [c_void, c_i64, P(C.c_uint32), P(C.c_uint32), c_void]),
    "mava_reduce_clip_adam_pair": (c_int, [c_void, c_void, c_void, c_void, P(PeerGroup), c_void,
                                           c_i64, c_i64, P(MlpDesc), c_void, P(MlpDesc), c_void,
                                           c_f32, c_f32, c_f32, c_f32, c_int, c_int, c_void, c_void]),
    "mava_reduce_clip_adam_pair_acc": (c_int, [c_void, c_void, c_void, c_void, P(PeerGroup), c_void,
                                               c_i64, c_i64, P(MlpDesc), c_void, P(MlpDesc), c_void,
                                               c_f32, c_f32, c_f32, c_f32, c_int, c_int, c_void,
                                               c_void, P(PpoHyper), c_i64, c_void]),
    "mava_ppo_loss_grad_bf16_acc": (c_int, [P(MlpDesc), c_void, c_void, P(MlpDesc), c_void, c_void,
                                            P(PpoHyper)] + [c_void] * 8 +
                                    [c_int, c_int, c_void, c_void, c_void, c_void]),
    "mava_clip_adam": (c_int, [c_void, c_void, c_void, c_void, c_void, c_i64, c_f32, c_f32, c_f32,
                               c_int, c_int, c_void]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m mava_b200.build` "
            "(there is no CPU or PyTorch fallback for the mava_b200 kernels)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.mava_abi_version() != 1:
        raise RuntimeError("libmava_b200.so ABI version mismatch")
    _lib = lib
    return lib


class MavaNativeError(RuntimeError):
    pass


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().mava_error_string(code).decode()
        raise MavaNativeError(f"{what}: {msg} (code {code})")
