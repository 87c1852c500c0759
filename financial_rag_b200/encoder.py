"""B200QueryEncoder -- the query-side BERT forward pass on the GPU.

The reference embeds every query (variant) on the CPU before it searches: ``SentenceTransformer(model).encode(q)``
(rag_backend.py:677, parent_child/retriever.py:87), or its own fallback ``LocalEmbedder.encode``
(local_embedder.py:155-191) -- tokenise, BERT forward, pooling (:171-179), L2 normalisation (:182).  Both encoders of
the ensemble are 12-layer BERT-384 models (local_models/BAAI-bge-small-en-v1.5: CLS pooling; local_models/thenlper-gte-small:
mean pooling, per ``1_Pooling/config.json``).  Here the forward pass runs in hand-written sm_100a kernels
(csrc/encoder.cu: tcgen05 GEMMs fed by TMA, fp32 attention / LayerNorm / pooling) and hands the normalised query
block to the scan where it already is: ``encode_ids_device`` returns a CUDA tensor laid out as ``search_device`` reads
its queries.  Tokenisation stays on the host (a WordPiece tokenizer object is injected or built from ``vocab.txt``).

``encode`` has SentenceTransformer's call shape, so an instance can stand in for ``member["embedder"]`` of the
ensemble (rag_backend.py:611-643) and for ``ParentContextRetriever``'s embedders unchanged.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Any, Callable, Dict, List, Mapping, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check
from .index import _stream_ptr

_POOLING = {"cls": 0, "mean": 1}


def pooling_of(model_dir: str) -> str:
    """``1_Pooling/config.json`` of a sentence-transformers model directory -> 'cls' | 'mean'."""
    with open(os.path.join(model_dir, "1_Pooling", "config.json")) as f:
        cfg = json.load(f)
    if cfg.get("pooling_mode_cls_token"):
        return "cls"
    if cfg.get("pooling_mode_mean_tokens"):
        return "mean"
    raise ValueError(f"{model_dir}: only CLS and mean pooling are implemented (got {cfg})")


class B200QueryEncoder:
    def __init__(self, config: Mapping[str, Any], state_dict: Mapping[str, Any], pooling: str = "cls",
                 normalize: bool = True, tokenizer: Optional[Callable] = None, device: int = 0):
        """``config``: the model's ``config.json`` (BertConfig fields); ``state_dict``: Hugging Face ``BertModel``
        parameter names -> fp32 arrays / tensors; ``tokenizer(list_of_texts) -> {"input_ids": [[...], ...]}``."""
        if pooling not in _POOLING:
            raise ValueError(f"pooling must be 'cls' or 'mean', got {pooling!r}")
        if config.get("hidden_act", "gelu") != "gelu" or config.get("position_embedding_type", "absolute") != "absolute":
            raise ValueError("the encoder kernels implement BERT with erf-GELU and absolute position embeddings")
        self._lib = _lib.load()
        self.device = int(device)
        self.pooling, self.normalize, self.tokenizer = pooling, bool(normalize), tokenizer
        self.hidden_size = int(config["hidden_size"])
        self.max_tokens = int(config.get("max_position_embeddings", 512))
        self.pad_token_id = int(config.get("pad_token_id", 0))
        h = ctypes.c_void_p()
        check(self._lib.fr_encoder_create(self.device, int(config["vocab_size"]), self.hidden_size,
                                          int(config["num_hidden_layers"]), int(config["num_attention_heads"]),
                                          int(config["intermediate_size"]), self.max_tokens,
                                          int(config.get("type_vocab_size", 2)), float(config.get("layer_norm_eps", 1e-12)),
                                          ctypes.byref(h)))
        self._h = h
        used = ctypes.c_int()
        for name, value in state_dict.items():
            if hasattr(value, "detach"):
                value = value.detach().cpu().numpy()
            a = np.ascontiguousarray(value, dtype=np.float32)
            check(self._lib.fr_encoder_set_tensor(h, name.encode(), a.ctypes.data, a.size, ctypes.byref(used)))
        check(self._lib.fr_encoder_finalize(h))  # names whatever is missing

    @classmethod
    def from_pretrained_dir(cls, model_dir: str, device: int = 0, tokenizer: Optional[Callable] = None) -> "B200QueryEncoder":
        """A sentence-transformers model directory as the reference ships them under ``local_models/``:
        ``config.json``, ``1_Pooling/config.json``, ``model.safetensors`` (or ``pytorch_model.bin``), ``vocab.txt``."""
        with open(os.path.join(model_dir, "config.json")) as f:
            config = json.load(f)
        st = os.path.join(model_dir, "model.safetensors")
        if os.path.exists(st) and os.path.getsize(st) > 4096:  # (a few hundred bytes = a git-lfs pointer, not weights)
            from safetensors.numpy import load_file

            state = load_file(st)
        else:
            import torch

            pt = os.path.join(model_dir, "pytorch_model.bin")
            if not os.path.exists(pt) or os.path.getsize(pt) <= 4096:
                raise FileNotFoundError(f"{model_dir}: no weights (model.safetensors / pytorch_model.bin are missing or "
                                        "git-lfs pointers)")
            state = torch.load(pt, map_location="cpu")
        if tokenizer is None:
            from transformers import BertTokenizerFast

            tok = BertTokenizerFast(vocab_file=os.path.join(model_dir, "vocab.txt"), do_lower_case=True)
            max_len = int(config.get("max_position_embeddings", 512))
            tokenizer = lambda texts: tok(list(texts), truncation=True, max_length=max_len)  # noqa: E731
        normalize = os.path.exists(os.path.join(model_dir, "2_Normalize")) or True  # both reference models normalise
        return cls(config, state, pooling_of(model_dir), normalize, tokenizer, device)

    # -- lifecycle ----------------------------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.fr_encoder_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("B200QueryEncoder is closed")
        return self._h

    # -- token ids in -------------------------------------------------------------------------------------------
    @staticmethod
    def pad_ids(id_lists: Sequence[Sequence[int]], pad_id: int = 0):
        """Ragged token id lists -> (ids [B, T] int32 right-padded, lens [B] int32)."""
        lens = np.array([len(x) for x in id_lists], dtype=np.int32)
        t = int(lens.max()) if len(lens) else 1
        ids = np.full((len(id_lists), max(t, 1)), pad_id, dtype=np.int32)
        for i, x in enumerate(id_lists):
            ids[i, :len(x)] = x
        return ids, lens

    def encode_ids(self, ids, lens=None, return_hidden: bool = False):
        """Host buffers: ids [B, T] (right-padded), lens [B] valid tokens (default: all T).  Returns the pooled
        [B, hidden] fp32 block (and the last hidden state [B, T, hidden] when asked)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if ids.ndim != 2:
            raise ValueError("ids must be [B, T]")
        b, t = ids.shape
        lens = np.full((b,), t, dtype=np.int32) if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
        out = np.empty((b, self.hidden_size), dtype=np.float32)
        hid = np.empty((b, t, self.hidden_size), dtype=np.float32) if return_hidden else None
        check(self._lib.fr_encoder_forward(self._handle(), ids.ctypes.data, lens.ctypes.data, b, t,
                                           _POOLING[self.pooling], 1 if self.normalize else 0, out.ctypes.data,
                                           hid.ctypes.data if hid is not None else None))
        return (out, hid) if return_hidden else out

    def encode_ids_device(self, ids, lens, out=None, stream=None):
        """CUDA tensors in (ids int32 [B, T], lens int32 [B]) -> CUDA tensor [B, hidden] fp32, enqueued on ``stream``
        (default: torch's current stream of this device): feed it straight to ``search_device``."""
        import torch

        if not (ids.is_cuda and lens.is_cuda and ids.dtype == torch.int32 and lens.dtype == torch.int32
                and ids.is_contiguous() and lens.is_contiguous() and ids.ndim == 2 and ids.device.index == self.device):
            raise ValueError(f"ids / lens must be contiguous int32 CUDA tensors on device {self.device}")
        b, t = ids.shape
        if out is None:
            out = torch.empty((b, self.hidden_size), dtype=torch.float32, device=ids.device)
        check(self._lib.fr_encoder_forward_device(self._handle(), ids.data_ptr(), lens.data_ptr(), b, t,
                                                  _POOLING[self.pooling], 1 if self.normalize else 0, out.data_ptr(), None,
                                                  _stream_ptr(stream, self.device)))
        return out

    # -- SentenceTransformer's call shape ----------------------------------------------------------------------------
    def tokenize(self, texts: Sequence[str]):
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer attached: pass tokenizer= or use encode_ids")
        enc = self.tokenizer(list(texts))
        id_lists = enc["input_ids"] if isinstance(enc, Mapping) or hasattr(enc, "keys") else enc
        return self.pad_ids([list(x)[: self.max_tokens] for x in id_lists], self.pad_token_id)

    def encode(self, sentences, convert_to_numpy: bool = True, batch_size: int = 64, **_unused):
        """``SentenceTransformer.encode``: a str gives a (hidden,) vector, a list gives (n, hidden)."""
        single = isinstance(sentences, str)
        texts = [sentences] if single else list(sentences)
        parts = []
        for lo in range(0, len(texts), batch_size):
            ids, lens = self.tokenize(texts[lo:lo + batch_size])
            parts.append(self.encode_ids(ids, lens))
        out = np.concatenate(parts) if parts else np.zeros((0, self.hidden_size), np.float32)
        if single:
            out = out[0]
        if convert_to_numpy:
            return out
        import torch

        return torch.from_numpy(out)

    def get_sentence_embedding_dimension(self) -> int:
        return self.hidden_size


__all__ = ["B200QueryEncoder", "pooling_of"]
