from .native_env import EnvState, NativeMarlEnv, ObservationSpec  # noqa: F401
from .synthetic import SyntheticSmaxEnv  # noqa: F401
