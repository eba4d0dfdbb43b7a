from .native_env import EnvState, NativeMarlEnv, ObservationSpec  # noqa: F401
