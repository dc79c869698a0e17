"""multimodalautoencoder_b200 -- B200-native engine for the MultimodalAutoencoder hot path."""
from .engine import Engine, EngineConfig, EngineError, debug_gemm  # noqa: F401
