"""GPU parity tests of the query-side encoder: the hand-written BERT forward pass
(csrc/encoder.cu) against transformers.BertModel built from the reference's own model configs
(tests/golden/encoder_models.json <- local_models/*/config.json).  The reference tree holds git-lfs pointers instead
of weights, so the model is randomly initialised (and scaled so that attention and LayerNorm are not trivial);
the forward pass is what is under test.  Tolerance: 1e-3 relative to the largest reference value, fp32
(the kernels reach ~1e-5: tensor-core GEMMs on two bf16 terms per operand)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BGE, GTE = "BAAI/bge-small-en-v1.5", "thenlper/gte-small"


@pytest.fixture(scope="module")
def frb():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback to test)")
    import financial_rag_b200 as f

    return f


@pytest.fixture(scope="module")
def models():
    with open(os.path.join(ROOT, "tests", "golden", "encoder_models.json")) as f:
        return json.load(f)


def reference_model(cfg, seed, layers=None):
    from transformers import BertConfig, BertModel

    cfg = dict(cfg)
    if layers is not None:
        cfg["num_hidden_layers"] = layers
    torch.manual_seed(seed)
    model = BertModel(BertConfig(**cfg), add_pooling_layer=False).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("LayerNorm.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif name.endswith(".bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif "embeddings" in name:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.06 * torch.randn(p.shape, generator=g))  # 3 x the init std: attention is far from uniform
    return model


def reference_forward(model, ids, lens, pooling):
    ids_t = torch.from_numpy(ids.astype(np.int64))
    mask = (torch.arange(ids.shape[1])[None, :] < torch.from_numpy(lens.astype(np.int64))[:, None]).to(torch.int64)
    with torch.no_grad():
        hid = model(input_ids=ids_t, attention_mask=mask).last_hidden_state
        if pooling == "cls":
            pooled = hid[:, 0]
        else:  # local_embedder.py:171-179
            m = mask.unsqueeze(-1).to(hid.dtype)
            pooled = (hid * m).sum(1) / m.sum(1)
        pooled = torch.nn.functional.normalize(pooled, p=2, dim=1)  # local_embedder.py:182
    return hid.numpy(), pooled.numpy(), mask.numpy().astype(bool)


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", [BGE, GTE])
@pytest.mark.parametrize("B,T", [(1, 5), (3, 16), (9, 200), (2, 512)])
def test_encoder_matches_transformers_bert(frb, models, name, B, T):
    spec = models[name]
    cfg = spec["config"]
    model = reference_model(cfg, seed=B * 1000 + T)
    enc = frb.B200QueryEncoder(cfg, model.state_dict(), pooling=spec["pooling"])
    rng = np.random.default_rng(B + T)
    ids = rng.integers(0, cfg["vocab_size"], size=(B, T)).astype(np.int32)
    lens = rng.integers(max(1, T // 3), T + 1, size=B).astype(np.int32)
    lens[0] = T
    for b in range(B):
        ids[b, lens[b]:] = cfg["pad_token_id"]
    want_h, want_p, mask = reference_forward(model, ids, lens, spec["pooling"])
    got_p, got_h = enc.encode_ids(ids, lens, return_hidden=True)
    assert got_p.shape == (B, 384) and got_h.shape == (B, T, 384)
    assert rel_err(got_h[mask], want_h[mask]) < 1e-3, "last hidden state (valid tokens)"
    assert rel_err(got_p, want_p) < 1e-3, "pooled, normalised embedding"
    np.testing.assert_allclose(np.linalg.norm(got_p, axis=1), 1.0, atol=1e-5)
    assert float((got_p * want_p).sum(1).min()) > 1.0 - 1e-6
    # what the kernels actually reach (two bf16 terms per operand): an order of magnitude inside the gate
    assert rel_err(got_p, want_p) < 2e-4
    enc.close()


def test_encoder_feeds_the_scan_without_leaving_the_device(frb, models):
    """token ids -> normalised query block -> exact top-k, all on the device; equals the host round trip, and
    ``encode`` keeps SentenceTransformer's call shape (str -> (d,), list -> (n, d)) so it can stand in for the
    ensemble's embedders (rag_backend.py:611-643, retriever.py:87)."""
    spec = models[BGE]
    cfg = dict(spec["config"])
    model = reference_model(cfg, seed=5, layers=3)
    cfg["num_hidden_layers"] = 3
    vocab = {w: i + 1000 for i, w in enumerate("what was the revenue growth in fiscal 2023 for segment a b c".split())}

    def toy_tokenizer(texts):  # stands in for WordPiece: [CLS] words [SEP]
        return {"input_ids": [[101] + [vocab.get(w, 100) for w in t.lower().split()] + [102] for t in texts]}

    enc = frb.B200QueryEncoder(cfg, model.state_dict(), pooling="cls", tokenizer=toy_tokenizer)
    texts = ["what was the revenue growth", "fiscal 2023 segment a", "c"]
    emb = enc.encode(texts)
    assert emb.shape == (3, 384) and enc.encode(texts[0]).shape == (384,)
    np.testing.assert_allclose(enc.encode(texts[1]), emb[1], atol=1e-6)   # padding does not change a sentence's vector
    assert isinstance(enc.encode(texts, convert_to_numpy=False), torch.Tensor)
    ids, lens = enc.tokenize(texts)
    _, want_p, _ = reference_forward(model, ids, lens, "cls")
    assert rel_err(emb, want_p) < 1e-3
    # a collection of documents embedded by the same encoder; queries never leave the device
    rng = np.random.default_rng(0)
    doc_ids = rng.integers(1000, 1013, size=(500, 12)).astype(np.int32)
    doc_ids[:, 0] = 101
    docs = enc.encode_ids(doc_ids)
    ix = frb.ShardIndex(dim=384, dtype="f32")
    ix.upsert(docs, np.arange(500, dtype=np.int64))
    dev = torch.device("cuda", 0)
    q_dev = enc.encode_ids_device(torch.from_numpy(doc_ids[:7]).to(dev), torch.full((7,), 12, dtype=torch.int32, device=dev))
    d_dev, k_dev = ix.search_device(q_dev, 5)
    torch.cuda.synchronize()
    assert k_dev[:, 0].cpu().tolist() == list(range(7))          # each document finds itself first
    d_host, k_host = ix.search(docs[:7], 5)
    assert (k_dev.cpu().numpy() == k_host).all()
    np.testing.assert_allclose(d_dev.cpu().numpy(), d_host, atol=2e-6)
    ix.close()
    enc.close()


def test_encoder_refuses_incomplete_weights(frb, models):
    cfg = dict(models[GTE]["config"])
    cfg["num_hidden_layers"] = 2
    model = reference_model(cfg, seed=1, layers=2)
    state = {k: v for k, v in model.state_dict().items() if "layer.1.output.dense" not in k}
    with pytest.raises(Exception, match="encoder.layer.1.output.dense"):
        frb.B200QueryEncoder(cfg, state, pooling="mean")
    bad = dict(model.state_dict())
    bad["encoder.layer.0.intermediate.dense.weight"] = torch.zeros(7, 7)
    with pytest.raises(Exception, match="elements"):
        frb.B200QueryEncoder(cfg, bad, pooling="mean")
