#!/usr/bin/env python3
"""Generate the committed golden fixtures from the read-only reference tree.

Run in the authoring container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (small, committed):
  tests/golden/chroma_fixture.json   -- the 18 fp32 vectors Chroma's WAL holds
      (reference: .chroma_children/chroma.sqlite3, table embeddings_queue), hex-encoded
      little-endian fp32 blobs so they survive bit-exactly, with ids, collection names and
      the stored metadata (parent_id / snippet / context).
  tests/golden/rrf_traces.json       -- every ``retrieval_score`` the reference wrote into
      test_logs/query_trace_*.json (rag_backend.py:1258-1289).  Each is a sum of
      1/(60+rank) terms (rag_backend.py:720-731), which pins k_rrf=60 and rank base 1.

Nothing here is executed at test time; the tests read the two JSON files.
"""
import glob
import hashlib
import json
import os
import sqlite3

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
EXPECTED_SHA256 = "7f7ad93e8959bcbee9363d75aed434e0809abc938919f6ed6d7100ff84cb5e32"


def main() -> None:
    con = sqlite3.connect(
        f"file:{REF}/.chroma_children/chroma.sqlite3?mode=ro&immutable=1", uri=True
    )
    cur = con.cursor()
    coll_by_uuid = {}
    coll_cfg = {}
    for cid, name, dim, cfg in cur.execute(
        "select id, name, dimension, config_json_str from collections"
    ):
        coll_by_uuid[cid] = name
        coll_cfg[name] = {"dimension": dim, "config": json.loads(cfg)}
    meta = {
        cid: {"key": k, "value": v}
        for cid, k, v in cur.execute(
            "select collection_id, key, str_value from collection_metadata"
        )
    }
    rows = []
    h = hashlib.sha256()
    for seq_id, op, topic, rid, vec, enc, md in cur.execute(
        "select seq_id, operation, topic, id, vector, encoding, metadata "
        "from embeddings_queue order by seq_id"
    ):
        assert enc == "FLOAT32" and len(vec) == 384 * 4
        h.update(vec)
        rows.append(
            {
                "seq_id": seq_id,
                "operation": op,
                "collection": coll_by_uuid[topic.rsplit("/", 1)[1]],
                "id": rid,
                "vector_f32le_hex": vec.hex(),
                "metadata": json.loads(md) if md else None,
            }
        )
    digest = h.hexdigest()
    assert digest == EXPECTED_SHA256, digest
    out = {
        "source": ".chroma_children/chroma.sqlite3 (embeddings_queue), reference tree",
        "sha256_of_blobs_in_seq_order": digest,
        "collections": coll_cfg,
        "collection_metadata": {coll_by_uuid[c]: m for c, m in meta.items()},
        "rows": rows,
    }
    with open(os.path.join(HERE, "chroma_fixture.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)

    traces = []
    for p in sorted(glob.glob(f"{REF}/test_logs/query_trace_*.json")):
        d = json.load(open(p))
        rc = d.get("retrieved_children") or []
        if not rc:
            continue
        traces.append(
            {
                "file": os.path.basename(p),
                "n_queries": len(d.get("generated_queries") or []),
                "children": [
                    {"child_id": str(c["child_id"]), "retrieval_score": c["retrieval_score"]}
                    for c in rc
                ],
            }
        )
    with open(os.path.join(HERE, "rrf_traces.json"), "w") as f:
        json.dump({"source": "test_logs/query_trace_*.json", "traces": traces}, f, indent=0)
    print(f"wrote {len(rows)} vectors, {len(traces)} traces")


if __name__ == "__main__":
    main()
