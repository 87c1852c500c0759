"""One-off migration: replay a Chroma persist directory's write-ahead log into B200 collections.

A deployment that already holds ``.chroma_children/chroma.sqlite3`` (the reference's PersistentClient
directory, chroma_child_store.py:29-34) does not have to re-embed its documents: chromadb keeps every
add / update / upsert / delete in the ``embeddings_queue`` table (fp32 little-endian blobs + JSON metadata,
one topic per collection) until the log is purged.  This module replays that log, in ``seq_id`` order, into
the same-named B200 collections.  It reads sqlite only -- chromadb itself is not needed.

    python -m financial_rag_b200.migrate_chroma /path/to/.chroma_children [/path/to/new_persist_dir]

Limits: rows whose WAL entries chromadb has already purged (after its HNSW segment was synced and the
log cleaned) live only in the HNSW ``data_level0.bin`` files and are NOT recovered here; the function
reports how many live rows chromadb's own ``embeddings`` table lists so a short count is visible.
"""
from __future__ import annotations

import json
import os
import sqlite3
import sys
from typing import Dict, Optional

import numpy as np

from .collection import B200Client

# chromadb.types.Operation
OP_ADD, OP_UPDATE, OP_UPSERT, OP_DELETE = 0, 1, 2, 3


def replay_wal(chroma_dir: str, client: Optional[B200Client] = None, *, batch_rows: int = 4096) -> Dict[str, Dict[str, int]]:
    """Replay ``<chroma_dir>/chroma.sqlite3`` into ``client`` (default: a B200Client on the same directory).
    Returns {collection: {"replayed": ops applied, "count": rows now live, "chroma_rows": rows chromadb lists}}."""
    db_path = os.path.join(chroma_dir, "chroma.sqlite3")
    con = sqlite3.connect(f"file:{db_path}?mode=ro", uri=True)
    client = client or B200Client(path=chroma_dir)
    report: Dict[str, Dict[str, int]] = {}
    try:
        colls = {cid: (name, dim) for cid, name, dim in con.execute("select id, name, dimension from collections")}
        space_of = {}
        for cid, key, val in con.execute("select collection_id, key, str_value from collection_metadata"):
            if key == "hnsw:space":
                space_of[cid] = val
        for cid, (name, dim) in colls.items():
            cfg = con.execute("select config_json_str from collections where id = ?", (cid,)).fetchone()[0]
            space = space_of.get(cid)
            if space is None and cfg:
                try:
                    space = (json.loads(cfg).get("vector_index") or {}).get("hnsw", {}).get("space")
                except (ValueError, AttributeError):
                    space = None
            col = client.get_or_create_collection(name, metadata={"hnsw:space": space or "l2"})
            n_ops = 0
            pend_ids, pend_vecs, pend_meta = [], [], []

            def flush():
                if pend_ids:
                    col.upsert(ids=list(pend_ids), embeddings=np.stack(pend_vecs), metadatas=list(pend_meta))
                    pend_ids.clear(), pend_vecs.clear(), pend_meta.clear()

            q = ("select operation, id, vector, encoding, metadata from embeddings_queue "
                 "where topic like ? order by seq_id")
            for op, rid, vec, enc, md in con.execute(q, (f"%/{cid}",)):
                n_ops += 1
                if op == OP_DELETE:
                    flush()
                    col.delete(ids=[rid])
                    continue
                if vec is None:  # metadata-only update
                    flush()
                    key = col.key_of(rid)
                    if key is not None and md:
                        old = col.metadata_of_key(key) or {}
                        old.update(json.loads(md))
                        col.set_metadata(rid, old)
                    continue
                if enc != "FLOAT32":
                    raise ValueError(f"{name}: unsupported vector encoding {enc!r}")
                if op == OP_ADD and (rid in pend_ids or col.key_of(rid) is not None):
                    continue  # chromadb's add leaves an existing id untouched
                if rid in pend_ids:
                    flush()  # keep last-writer-wins order inside one batch explicit
                pend_ids.append(rid)
                pend_vecs.append(np.frombuffer(vec, dtype="<f4").astype(np.float32))
                pend_meta.append(json.loads(md) if md else None)
                if len(pend_ids) >= batch_rows:
                    flush()
            flush()
            chroma_rows = con.execute(
                "select count(*) from embeddings e join segments s on e.segment_id = s.id where s.collection = ?",
                (cid,)).fetchone()[0]
            report[name] = {"replayed": n_ops, "count": col.count(), "chroma_rows": int(chroma_rows)}
    finally:
        con.close()
    return report


if __name__ == "__main__":
    src = sys.argv[1]
    dst = B200Client(path=sys.argv[2]) if len(sys.argv) > 2 else None
    for name, r in replay_wal(src, dst).items():
        flag = "" if r["count"] >= r["chroma_rows"] else "   <-- WAL already purged for some rows: re-ingest those"
        print(f"{name}: {r['replayed']} log entries replayed, {r['count']} rows live (chromadb lists {r['chroma_rows']}){flag}")
