"""vlite_b200 -- B200-native retrieval hot path behind vlite's VLite / EmbeddingModel / Ctx API."""
__version__ = "0.1.0"
