"""``EmbeddingModel`` -- host-side mirror of the reference's ``vlite/model.py`` for the hot path.

Same names and argument meaning as the reference class; the arithmetic runs in the CUDA library:

  * ``search(query_embedding, embeddings, top_k)`` (vlite/model.py:76-90) -- the reference's
    stateless call: corpus passed by value, XOR-sum of byte values, ``top_k = min(top_k, N)``,
    k smallest.  Here the (N, W) array is repacked to raw ubinary ON THE DEVICE
    (``vl_corpus_append_f32``) and scanned by the sm_100a kernel; the result order is the
    canonical ascending (score, index) -- the reference's ``torch.topk`` leaves tie order
    unspecified (vlite/model.py:85).
  * ``embed(texts, precision="binary")`` (vlite/model.py:25-48) -- the Hugging Face forward pass
    stays on the reference path (torch + transformers, out of scope for this build); the
    quantisation tail ``normalize -> CLS -> (x > 0) -> packbits -> int8 - 128`` (vlite/model.py:35-44)
    runs on the device (``vl_quantize_embeddings`` for a CUDA tensor, ``vl_quantize_binary`` for a
    host array); ``embed_device`` also hands out the int8 / float32 forms the rescorer takes.  Returned values are the reference's stored encoding
    ``int8(u8) - 128`` as int16 in [-256, -1] (what numpy 1.x produces for that line).
  * ``hamming_distance`` (vlite/model.py:71-74) -- on 0/1 arrays, computed as popcount(XOR) of the
    packed bits on the device.

There is no CPU fallback: without the CUDA library and a B200 these methods raise.
"""
from __future__ import annotations

import numpy as np

from . import _capi

XORSUM, HAMMING = _capi.XORSUM, _capi.HAMMING


def to_reference_encoding(u8: np.ndarray) -> np.ndarray:
    """raw ubinary bytes -> the reference's stored ``int8(u8) - 128`` values (vlite/model.py:44)."""
    return (np.asarray(u8, dtype=np.uint8) ^ np.uint8(0x80)).astype(np.int16) - np.int16(256)


def from_reference_encoding(vals) -> np.ndarray:
    """The reference's stored values (or raw bytes in [0, 255]) -> raw ubinary bytes."""
    v = np.rint(np.asarray(vals)).astype(np.int64)
    out = np.where(v < 0, (v + 256) ^ 0x80, v)
    if ((out < 0) | (out > 255)).any():
        raise ValueError("binary vector values must lie in [-256, 255]")
    return out.astype(np.uint8)


class EmbeddingModel:
    def __init__(self, model_name="mixedbread-ai/mxbai-embed-large-v1", device="cuda", log_enabled=True,
                 embedder=None, gpu_index: int = 0):
        """``embedder``: optional callable ``texts -> float array (B, D)`` that stands in for the
        Hugging Face model (tests, or an embedding service).  The HF model itself is loaded lazily
        on the first ``embed`` call, so constructing the class needs neither network nor weights."""
        self.log_enabled = log_enabled
        self.model_name = model_name
        self.device = device
        self.gpu_index = gpu_index
        self.dimension = 1024          # vlite/model.py:19
        self.context_length = 512      # vlite/model.py:20
        self.embedding_dtype = "float32"
        self.binary_dims = 512         # vlite/model.py:39 (MRL truncation before sign quantisation)
        self._embedder = embedder
        self.tokenizer = None
        self.model = None

    # ---- embedding (forward pass stays on the reference path) ----
    def _load_hf(self):
        if self.model is not None:
            return
        import torch  # noqa: F401
        from transformers import AutoModel, AutoTokenizer
        self.tokenizer = AutoTokenizer.from_pretrained(self.model_name)
        self.model = AutoModel.from_pretrained(self.model_name).to(self.device)

    def _forward(self, texts):
        """float embeddings (B, D): the CLS token of the last hidden state (vlite/model.py:30-36).
        An ``embedder`` callable returns whatever it returns (numpy or a CUDA tensor); the HF path
        returns the raw CLS rows as a CUDA tensor -- normalising one token's row equals taking that
        token from the normalised tensor (model.py:35-36), and the fused kernel does it."""
        if self._embedder is not None:
            x = self._embedder(texts)
            return x if hasattr(x, "data_ptr") else np.asarray(x, dtype=np.float32)
        import torch
        self._load_hf()
        inputs = self.tokenizer(texts, padding=True, truncation=True, return_tensors="pt").to(self.device)
        with torch.no_grad():
            outputs = self.model(**inputs).last_hidden_state
        return outputs[:, 0].float().contiguous()

    def embed(self, texts, precision="binary"):
        if isinstance(texts, str):
            texts = [texts]
        if precision != "binary":
            raise ValueError(f"Unsupported precision: {precision}")
        return to_reference_encoding(self.embed_ubinary(texts))

    def embed_float(self, texts):
        """The float embeddings of ``texts`` before quantisation, (B, D): numpy or a CUDA tensor.
        ``VLite.retrieve`` hands them to ``vl_search_embeddings`` (quantise + scan in one call)."""
        if isinstance(texts, str):
            texts = [texts]
        x = self._forward(texts)
        if isinstance(x, np.ndarray):
            return np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
        return x.float().contiguous()

    def embed_device(self, texts, int8=False, float32=False, i8_scale=127.0):
        """The quantisation tail fused on the device (``vl_quantize_embeddings``): returns a dict of
        CUDA tensors -- ``"ubinary"`` (B, binary_dims/8) uint8 for the scan and, on request,
        ``"int8"`` (B, D) = rint(clamp(x_norm * i8_scale, -127, 127)) and ``"float32"`` (B, D) =
        the L2-normalised embedding, the forms ``vl_rescore`` takes.  Nothing touches the host."""
        import torch
        if isinstance(texts, str):
            texts = [texts]
        x = self._forward(texts)
        dev = torch.device("cuda", self.gpu_index)
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        x = x.to(dev, torch.float32).contiguous()
        n, dims = x.shape
        out = {"ubinary": torch.empty((n, self.binary_dims // 8), dtype=torch.uint8, device=dev)}
        if int8:
            out["int8"] = torch.empty((n, dims), dtype=torch.int8, device=dev)
        if float32:
            out["float32"] = torch.empty((n, dims), dtype=torch.float32, device=dev)
        _capi.quantize_embeddings(x, self.binary_dims, out["ubinary"], out.get("int8"), out.get("float32"),
                                  i8_scale=i8_scale, device=self.gpu_index,
                                  stream=torch.cuda.current_stream(dev).cuda_stream)
        return out

    def embed_ubinary(self, texts) -> np.ndarray:
        """Same as ``embed`` but returns raw ubinary bytes (B, binary_dims/8), the device format."""
        if isinstance(texts, str):
            texts = [texts]
        x = self._forward(texts)
        if not isinstance(x, np.ndarray):                      # CUDA tensor: fused tail, no host round trip
            import torch
            x = x.to(torch.float32).contiguous()
            if x.shape[1] % 32 or self.binary_dims % 32 or self.binary_dims > x.shape[1]:
                raise ValueError("embedding width must be a multiple of 32 bits")
            out = torch.empty((x.shape[0], self.binary_dims // 8), dtype=torch.uint8, device=x.device)
            _capi.quantize_embeddings(x, self.binary_dims, out, device=x.device.index or 0,
                                      stream=torch.cuda.current_stream(x.device).cuda_stream)
            return out.cpu().numpy()
        x = np.ascontiguousarray(x[:, : self.binary_dims], dtype=np.float32)
        if x.shape[1] % 32:
            raise ValueError("embedding width must be a multiple of 32 bits")
        return _capi.quantize_binary(x, device=self.gpu_index)

    # ---- search (vlite/model.py:76-90) ----
    def search(self, query_embedding, embeddings, top_k, metric: int = XORSUM):
        """query_embedding (W,), embeddings (N, W) in the reference's stored encoding (any numeric
        dtype; float32 is what VLite.rank_and_filter passes, vlite/main.py:146) or raw bytes.
        Returns (indices int64[k], scores int64[k]), k = min(top_k, N)."""
        e = np.asarray(embeddings)
        q = from_reference_encoding(np.asarray(query_embedding).reshape(-1))
        n, w = e.shape
        if w != q.shape[0]:
            raise ValueError("query and corpus widths differ")
        k = min(int(top_k), n)                                  # vlite/model.py:84
        if k <= 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        if k > _capi.MAX_K:
            raise ValueError(f"top_k={k} exceeds the device top-k limit VL_MAX_K={_capi.MAX_K}")
        with _capi.Corpus(w, device=self.gpu_index, capacity_rows=n) as c:
            if e.dtype == np.uint8:
                c.append(e)
            else:
                c.append_f32(np.ascontiguousarray(e, dtype=np.float32), _capi.SRC_F32_OFFSET)
            scores, rows = c.search(q[None, :], k, metric)
        return rows[0].astype(np.int64), scores[0].astype(np.int64)

    def hamming_distance(self, embedding1, embedding2):
        """np.sum(e1 != e2, axis=1) for 0/1 arrays (vlite/model.py:71-74): e1 (1, D) or (D,),
        e2 (N, D); D a multiple of 512 or 1024 bits the library supports."""
        a = np.asarray(embedding1).reshape(-1)
        b = np.atleast_2d(np.asarray(embedding2))
        if not (np.isin(a, (0, 1)).all() and np.isin(b, (0, 1)).all()):
            raise ValueError("hamming_distance expects unpacked 0/1 bit arrays")
        if a.shape[0] != b.shape[1] or a.shape[0] % 32:
            raise ValueError("bit arrays must have equal widths, a multiple of 32")
        # pack on the device: (x > 0) -> MSB-first bytes, same kernel as embed()'s quantisation
        qa = _capi.quantize_binary(a.astype(np.float32)[None, :], device=self.gpu_index)[0]
        pb = _capi.quantize_binary(np.ascontiguousarray(b, dtype=np.float32), device=self.gpu_index)
        with _capi.Corpus(pb.shape[1], device=self.gpu_index, capacity_rows=len(pb)) as c:
            c.append(pb)
            return c.scores(qa, HAMMING).astype(np.int64)

    def token_count(self, text: str) -> int:
        self._load_hf()
        return len(self.tokenizer(text, add_special_tokens=True)["input_ids"])
