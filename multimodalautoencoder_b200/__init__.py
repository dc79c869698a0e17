"""multimodalautoencoder_b200 -- B200-native engine for the MultimodalAutoencoder hot path."""
from .engine import Engine, EngineConfig, EngineError, debug_gemm  # noqa: F401
from .multimodal_autoencoder import MultimodalAutoencoder, get_rmse  # noqa: F401,E402
from .data_funcs import DataLoader  # noqa: F401,E402
