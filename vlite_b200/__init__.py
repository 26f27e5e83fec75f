"""vlite_b200 -- B200-native retrieval hot path behind vlite's VLite / EmbeddingModel / Ctx API.

Exports the same two names the reference package does (vlite/__init__.py): ``VLite`` and
``EmbeddingModel``; plus ``Ctx`` / ``CtxFile`` (vlite/ctx.py) and the low-level ``Corpus`` handle.
"""
__version__ = "0.1.0"

from .ctx import Ctx, CtxFile, CtxSectionType  # noqa: E402,F401
from .main import ShardedVLite, VLite  # noqa: E402,F401
from .model import EmbeddingModel  # noqa: E402,F401
from ._capi import Corpus  # noqa: E402,F401

__all__ = ["VLite", "ShardedVLite", "EmbeddingModel", "Ctx", "CtxFile", "CtxSectionType", "Corpus"]
