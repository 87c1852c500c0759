#!/usr/bin/env python3
"""Generate tests/golden/encoder_models.json from the read-only reference tree (authoring container only):
the BertConfig fields and the pooling mode of the two encoders of the reference's ensemble
(local_models/<model>/config.json, local_models/<model>/1_Pooling/config.json, modules.json).
The weights themselves are git-lfs pointers in the reference tree, so the GPU parity test builds a randomly
initialised transformers.BertModel from these configs and compares forward passes.

    python tests/golden/make_encoder_golden.py
"""
import json
import os

REF = "/root/reference/local_models"
HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = ("vocab_size", "hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "hidden_act",
        "max_position_embeddings", "type_vocab_size", "layer_norm_eps", "pad_token_id", "position_embedding_type")


def main() -> None:
    out = {}
    for d, hub_name in (("BAAI-bge-small-en-v1.5", "BAAI/bge-small-en-v1.5"), ("thenlper-gte-small", "thenlper/gte-small")):
        cfg = json.load(open(os.path.join(REF, d, "config.json")))
        pool = json.load(open(os.path.join(REF, d, "1_Pooling", "config.json")))
        modules = [m["type"].rsplit(".", 1)[1] for m in json.load(open(os.path.join(REF, d, "modules.json")))]
        out[hub_name] = {
            "source": f"local_models/{d}/config.json, 1_Pooling/config.json, modules.json",
            "config": {k: cfg[k] for k in KEEP if k in cfg},
            "pooling": "cls" if pool.get("pooling_mode_cls_token") else "mean",
            "modules": modules,  # Transformer -> Pooling -> Normalize
        }
    with open(os.path.join(HERE, "encoder_models.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
