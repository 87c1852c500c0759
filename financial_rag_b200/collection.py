"""Chroma-``Collection``-shaped object backed by a B200 ShardIndex.

The reference talks to chromadb through exactly this surface:
  * ``PersistentClient(path).get_or_create_collection(name, metadata={"hnsw:space": "cosine"})``
        parent_child/chroma_child_store.py:32-34, parent_child/multivector_store.py:51-52
  * ``col.upsert(ids, embeddings, metadatas)`` / ``col.delete(ids)`` / ``col.add(...)``
        chroma_child_store.py:54-59, multivector_store.py:136-139
  * ``col.query(query_embeddings=[...], n_results=k, include=[...])`` -> dict of lists of lists
        chroma_child_store.py:63-67, multivector_store.py:151-154
  * ``col.count()``   chroma_child_store.py:78
so ``multivector_store.py`` and ``chroma_child_store.py`` run unchanged on top of it.

Vectors live in HBM inside the ShardIndex; ids / metadata (the payload the reference stores as
``{"parent_id", "snippet", "context"}``) stay on the host keyed by the int64 row key.  The GPU index
must outlive the per-request store objects the reference constructs (rag_backend.py:611-643), so
collections live in a process-global registry keyed by (persist_dir, name).

Persistence replaces chromadb's ``.chroma_children/`` (sqlite WAL + HNSW bin files)
with one directory per collection, ``<persist_dir>/<name>.b200/``:
    meta.json        static description: format version, name, space, dtype, dim (written once, atomically)
    rows.bin         [rows][dim] bf16/fp32 -- the rows bit for bit as they sit in HBM, in global insertion order
                     (independent of how many GPUs the collection is sharded over)
    keys.bin         [rows] int64 row keys in insertion order (INT64_MIN = deleted row)
    payload.sqlite3  payload(key INTEGER PRIMARY KEY, id, metadata, document); groups(ordinal, name);
                     state(name, value): rows, generation, next_synthetic -- the COMMIT POINT of every flush
    patch.journal    rows overwritten / deleted in place by the flush in progress (absent otherwise)
    LOCK             flock()ed while a process has the collection open: a second process is refused, not corrupted
Restart = mmap rows.bin/keys.bin + one H2D copy (fr_index_import_raw): a reloaded collection returns
bit-identical results.  Like chromadb's PersistentClient, every mutation is flushed before the call
returns (``B200_CHILD_AUTOPERSIST=0`` turns that off; ``persist()`` flushes on demand); flushes are
incremental -- appended rows are appended, overwritten / deleted rows are patched in place.

A flush survives a crash at any point (``persist``): (1) appended rows are written past the committed row count and
fsynced -- invisible until the commit; (2) in-place patches go to patch.journal, fsynced; (3) ONE sqlite transaction
commits the payload changes together with the new row count and generation; (4) the patches are applied to
rows.bin / keys.bin and fsynced; (5) the journal is removed.  On load a journal whose generation equals the
committed one is replayed (the crash was after the commit), any other journal is discarded (the crash was before
it: rows.bin still holds the old bits, the payload is the old payload), and the row files are checked against the
committed row count.
"""
from __future__ import annotations

import fcntl
import json
import os
import sqlite3
import struct
import threading
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .group import ShardGroup, parse_devices
from .index import FR_MAX_K, ShardIndex, canonical_space

_INT64_MAX = (1 << 63) - 1
_INT64_MIN = -(1 << 63)
FORMAT_VERSION = 2
_JOURNAL_MAGIC = b"FRB2PJ01"


def _fsync_file(path: str) -> None:
    fd = os.open(path, os.O_RDONLY)
    try:
        os.fsync(fd)
    finally:
        os.close(fd)


def _write_atomic(path: str, data: bytes) -> None:
    """tmp + fsync + rename + fsync of the directory: the file is either the old one or the new one."""
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(data)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)
    _fsync_file(os.path.dirname(path) or ".")


def _autopersist() -> bool:
    return os.getenv("B200_CHILD_AUTOPERSIST", "1").strip().lower() not in ("0", "false", "no", "off")


def _as_matrix(embeddings, dim: Optional[int]) -> np.ndarray:
    """Accept list[list[float]], numpy, or torch tensors (retriever.py:87 passes tensors)."""
    if hasattr(embeddings, "detach"):
        embeddings = embeddings.detach().cpu().numpy()
    elif isinstance(embeddings, (list, tuple)) and len(embeddings) and hasattr(embeddings[0], "detach"):
        embeddings = [e.detach().cpu().numpy() for e in embeddings]
    m = np.asarray(embeddings, dtype=np.float32)
    if m.ndim == 1:
        m = m[None, :]
    if m.ndim == 3 and m.shape[1] == 1:  # [[tensor(1,d)]] from the local wrapper (local_embedder.py:187-191)
        m = m[:, 0, :]
    if m.ndim != 2:
        raise ValueError(f"embeddings must be 2-D, got shape {m.shape}")
    if dim is not None and m.shape[1] != dim:
        raise ValueError(f"embedding dimension {m.shape[1]} does not match collection dimension {dim}")
    return np.ascontiguousarray(m)


class B200Collection:
    def __init__(self, name: str, metadata: Optional[Dict[str, Any]] = None, *, dtype: Optional[str] = None,
                 device: Optional[int] = None, directory: Optional[str] = None,
                 devices: Optional[Sequence[int]] = None):
        """``devices`` (or ``B200_CHILD_DEVICES=0,1,...|all``): row-shard the collection over these GPUs (SURVEY.md 8e:
        cyclic placement, local top-k per GPU, NCCL all-gather, merge kernel) -- same calls, same answers as on one GPU.
        Without it the collection is one shard on ``device`` / ``B200_CHILD_DEVICE``."""
        self.name = name
        self.directory = directory          # <persist_dir>/<name>.b200, None = memory only
        self._persisted_rows = 0            # rows already in rows.bin / keys.bin
        self._dirty_rows: set = set()       # physical rows < _persisted_rows whose bits or key changed
        self._dirty_payload: set = set()    # keys whose payload row must be rewritten (or removed)
        self.metadata = dict(metadata or {})
        self.space = canonical_space(self.metadata.get("hnsw:space"))
        self.dtype = dtype or os.getenv("B200_CHILD_DTYPE", "bf16")
        self.device = int(os.getenv("B200_CHILD_DEVICE", "0")) if device is None else int(device)
        if devices is None and device is None:
            devices = parse_devices(os.getenv("B200_CHILD_DEVICES"))
        self.devices: Optional[List[int]] = [int(d) for d in devices] if devices else None
        if self.devices and len(self.devices) == 1:
            self.device, self.devices = self.devices[0], None
        self._index: Optional[ShardIndex] = None  # created at first upsert: the dimension is not known before
        self._lock = threading.RLock()
        self._key_of_id: Dict[str, int] = {}
        self._payload: Dict[int, Dict[str, Any]] = {}  # key -> {"id", "metadata", "document"}
        self._next_synthetic = -2  # -1 is FR_KEY_NONE; ids that are not int64 decimals get negative keys
        # dense ordinals for callers that pack structure into row keys (multivector_store.py: one ordinal per child in
        # the upper key bits).  They belong to the collection, not to the store objects the reference constructs per
        # request (rag_backend.py:656, pipeline.py:25): every object must see the same child <-> ordinal map.
        self._group_ordinal: Dict[str, int] = {}
        self._group_names: List[str] = []
        self._dirty_groups: List[int] = []
        self._row_files = ("rows.bin", "keys.bin")  # a compacting reload switches to fresh files by committing new names
        self._generation = 0                # bumped by every committed flush
        self._lock_fd: Optional[int] = None  # flock on <directory>/LOCK while this process has the collection open

    # -- helpers -------------------------------------------------------------------------------
    @property
    def dim(self) -> Optional[int]:
        return self._index.dim if self._index is not None else None

    @property
    def index(self) -> Optional[ShardIndex]:
        return self._index

    def _key_for(self, id_str: str) -> int:
        k = self._key_of_id.get(id_str)
        if k is not None:
            return k
        key = None
        if id_str.isdigit() and str(int(id_str)) == id_str and int(id_str) <= _INT64_MAX:
            key = int(id_str)  # Snowflake child ids (snowflake_id.py:27-49) map to themselves
        if key is None:
            key = self._next_synthetic
            self._next_synthetic -= 1
        self._key_of_id[id_str] = key
        return key

    def _ensure_index(self, dim: int, reserve_rows: int = 0) -> ShardIndex:
        if self._index is None:
            if self.devices:  # one process, several GPUs: the row-sharded group behind the same calls
                self._index = ShardGroup(dim=dim, space=self.space, dtype=self.dtype, devices=self.devices,
                                         reserve_rows=reserve_rows,
                                         exchange=os.getenv("B200_CHILD_EXCHANGE", "auto").strip().lower())
            else:
                self._index = ShardIndex(dim=dim, space=self.space, dtype=self.dtype, device=self.device,
                                         reserve_rows=reserve_rows)
        return self._index

    def _lock_directory(self) -> None:
        """One process at a time per collection directory (the reference may run ingest_all.py next to the API
        server): the second one is refused with a clear error instead of truncating rows the first one appended."""
        if self.directory is None or self._lock_fd is not None:
            return
        os.makedirs(self.directory, exist_ok=True)
        fd = os.open(os.path.join(self.directory, "LOCK"), os.O_RDWR | os.O_CREAT, 0o644)
        try:
            fcntl.flock(fd, fcntl.LOCK_EX | fcntl.LOCK_NB)
        except OSError:
            try:
                holder = os.pread(fd, 64, 0).decode("ascii", "replace").strip()
            finally:
                os.close(fd)
            raise RuntimeError(f"collection {self.name!r} at {self.directory} is open in another process"
                               f"{' (pid ' + holder + ')' if holder else ''}; B200 collections are single-writer")
        os.ftruncate(fd, 0)
        os.pwrite(fd, str(os.getpid()).encode(), 0)
        self._lock_fd = fd

    # -- group ordinals (shared by every store object on this collection) ------------------------------------
    def group_ordinal(self, name: str, create: bool = True) -> Optional[int]:
        """Dense ordinal of ``name`` (e.g. a child id); allocated under the collection lock, persisted with it."""
        with self._lock:
            o = self._group_ordinal.get(name)
            if o is None and create:
                o = len(self._group_names)
                self._group_ordinal[name] = o
                self._group_names.append(name)
                self._dirty_groups.append(o)
            return o

    def group_name(self, ordinal: int) -> Optional[str]:
        with self._lock:
            return self._group_names[ordinal] if 0 <= ordinal < len(self._group_names) else None

    def payload_of_key(self, key: int) -> Optional[Dict[str, Any]]:
        """{"id", "metadata", "document"} of a row key, None when the row is gone."""
        with self._lock:
            return self._payload.get(int(key))

    def search_arrays(self, query_embeddings, n_results: int):
        """The scan without the dict building: (dist [B, k] fp32, keys [B, k] int64, -1 / +inf padded)."""
        q = _as_matrix(query_embeddings, self.dim)
        with self._lock:
            idx = self._index
        n_results = int(n_results)
        if n_results < 1:
            raise ValueError("n_results must be >= 1")
        if idx is None or idx.count() == 0:
            return (np.full((q.shape[0], n_results), np.inf, dtype=np.float32),
                    np.full((q.shape[0], n_results), -1, dtype=np.int64))
        if n_results > FR_MAX_K:
            raise ValueError(f"n_results = {n_results} exceeds FR_MAX_K = {FR_MAX_K}")
        return idx.search(q, n_results)  # the index serialises against upserts itself

    # -- chromadb.Collection surface -------------------------------------------------------------
    def count(self) -> int:
        with self._lock:
            return 0 if self._index is None else self._index.count()

    def key_of(self, id_str: str) -> Optional[int]:
        with self._lock:
            k = self._key_of_id.get(str(id_str))
            return k if k is not None and k in self._payload else None

    def metadata_of_key(self, key: int) -> Optional[dict]:
        with self._lock:
            p = self._payload.get(int(key))
            return p["metadata"] if p is not None else None

    def set_metadata(self, id_str: str, metadata: Optional[dict]) -> None:
        """Replace the metadata of an existing id (chromadb ``update`` without embeddings)."""
        with self._lock:
            k = self.key_of(id_str)
            if k is None:
                return
            self._payload[k]["metadata"] = dict(metadata) if metadata is not None else None
            if self.directory is not None:
                self._dirty_payload.add(k)
                if _autopersist():
                    self.persist()

    def upsert(self, ids: Sequence[str], embeddings=None, metadatas: Optional[Sequence[Optional[dict]]] = None,
               documents: Optional[Sequence[Optional[str]]] = None, keys: Optional[Sequence[int]] = None) -> None:
        """``keys`` (ours, not chromadb's): explicit int64 row keys for the ids, for callers that encode
        structure in them (multivector_store.py: (child_ordinal << 16) | token_idx)."""
        if embeddings is None:
            raise ValueError("B200Collection needs explicit embeddings (the reference always passes them)")
        ids = [str(i) for i in ids]
        if not ids:
            return
        with self._lock:
            m = _as_matrix(embeddings, self.dim)
            if m.shape[0] != len(ids):
                raise ValueError("ids and embeddings differ in length")
            idx = self._ensure_index(m.shape[1])
            if keys is not None:
                if len(keys) != len(ids):
                    raise ValueError("ids and keys differ in length")
                for i, k in zip(ids, keys):
                    old = self._key_of_id.get(i)
                    if old is not None and old != int(k) and old in self._payload:
                        raise ValueError(f"id {i!r} is already stored under key {old}")
                    owner = self._payload.get(int(k))
                    if owner is not None and owner["id"] != i:
                        raise ValueError(f"key {int(k)} already belongs to id {owner['id']!r}, not {i!r}")
                for i, k in zip(ids, keys):
                    self._key_of_id[i] = int(k)
            keys = np.array([self._key_for(i) for i in ids], dtype=np.int64)
            idx.upsert(m, keys)
            for j, (i, k) in enumerate(zip(ids, keys.tolist())):
                self._payload[k] = {
                    "id": i,
                    "metadata": dict(metadatas[j]) if metadatas is not None and metadatas[j] is not None else None,
                    "document": documents[j] if documents is not None else None,
                }
            if self.directory is not None:
                rows = idx.lookup_rows(keys)
                self._dirty_rows.update(int(r) for r in rows.tolist() if 0 <= r < self._persisted_rows)
                self._dirty_payload.update(keys.tolist())
                if _autopersist():
                    self.persist()

    def add(self, ids: Sequence[str], embeddings=None, metadatas=None, documents=None) -> None:
        """chromadb ``add`` leaves existing ids untouched; only new ids are inserted."""
        ids = [str(i) for i in ids]
        with self._lock:
            m = _as_matrix(embeddings, self.dim) if embeddings is not None else None
            keep = [j for j, i in enumerate(ids)
                    if i not in self._key_of_id or self._key_of_id[i] not in self._payload]
            first = {}
            for j in keep:  # an id repeated inside one add keeps its first occurrence
                first.setdefault(ids[j], j)
            keep = sorted(first.values())
            if not keep:
                return
            self.upsert(
                [ids[j] for j in keep],
                m[keep] if m is not None else None,
                [metadatas[j] for j in keep] if metadatas is not None else None,
                [documents[j] for j in keep] if documents is not None else None,
            )

    def delete(self, ids: Optional[Sequence[str]] = None, where: Optional[dict] = None) -> None:
        with self._lock:
            if self._index is None:
                return
            keys: List[int] = []
            if ids is not None:
                for i in ids:
                    k = self._key_of_id.get(str(i))
                    if k is not None and k in self._payload:
                        keys.append(k)
            elif where:
                for k, p in self._payload.items():
                    md = p.get("metadata") or {}
                    if all(md.get(f) == v for f, v in where.items()):
                        keys.append(k)
            if keys:
                karr = np.array(keys, dtype=np.int64)
                if self.directory is not None:
                    rows = self._index.lookup_rows(karr)
                    self._dirty_rows.update(int(r) for r in rows.tolist() if 0 <= r < self._persisted_rows)
                    self._dirty_payload.update(keys)
                self._index.delete(karr)
                for k in keys:
                    p = self._payload.pop(k, None)
                    if p is not None:
                        self._key_of_id.pop(p["id"], None)
                if self.directory is not None and _autopersist():
                    self.persist()

    def get(self, ids: Optional[Sequence[str]] = None, include: Optional[Sequence[str]] = None,
            where: Optional[dict] = None, limit: Optional[int] = None) -> Dict[str, Any]:
        include = list(include) if include is not None else ["metadatas", "documents"]
        with self._lock:
            if ids is not None:
                keys = [self._key_of_id[str(i)] for i in ids
                        if str(i) in self._key_of_id and self._key_of_id[str(i)] in self._payload]
            else:
                keys = list(self._payload.keys())
            if where:
                keys = [k for k in keys
                        if all((self._payload[k].get("metadata") or {}).get(f) == v for f, v in where.items())]
            if limit is not None:
                keys = keys[:limit]
            out: Dict[str, Any] = {"ids": [self._payload[k]["id"] for k in keys]}
            out["metadatas"] = [self._payload[k]["metadata"] for k in keys] if "metadatas" in include else None
            out["documents"] = [self._payload[k]["document"] for k in keys] if "documents" in include else None
            out["embeddings"] = None
            return out

    def query(self, query_embeddings=None, n_results: int = 10, include: Optional[Sequence[str]] = None,
              **_unused) -> Dict[str, Any]:
        """Batched exact k-NN.  Returns chromadb's dict-of-lists-of-lists
        (consumed at chroma_child_store.py:65-67 and multivector_store.py:152-154)."""
        if query_embeddings is None:
            raise ValueError("query_embeddings is required (no embedding function is attached)")
        include = list(include) if include is not None else ["metadatas", "documents", "distances"]
        n_results = int(n_results)
        if n_results < 1:
            raise ValueError("n_results must be >= 1")
        # the scan runs WITHOUT the collection lock (concurrent Flask threads, api_server.py:1366-1371, only meet in the
        # index, which orders them on the GPU); the lock is taken to map keys to payloads
        d, keys = self.search_arrays(query_embeddings, n_results)
        b = d.shape[0]
        ids: List[List[str]] = [[] for _ in range(b)]
        dists: List[List[float]] = [[] for _ in range(b)]
        metas: List[List[Optional[dict]]] = [[] for _ in range(b)]
        docs: List[List[Optional[str]]] = [[] for _ in range(b)]
        with self._lock:
            for i in range(b):
                for dist, key in zip(d[i].tolist(), keys[i].tolist()):
                    if key == -1:
                        break
                    p = self._payload.get(key)
                    if p is None:  # deleted between the scan and here
                        continue
                    ids[i].append(p["id"])
                    dists[i].append(dist)
                    metas[i].append(p["metadata"])
                    docs[i].append(p["document"])
            return {
                "ids": ids,
                "distances": dists if "distances" in include else None,
                "metadatas": metas if "metadatas" in include else None,
                "documents": docs if "documents" in include else None,
                "embeddings": None,
                "uris": None,
                "data": None,
                "included": include,
            }

    # -- persistence ------------------------------------------------------------------------------
    def _paths(self):
        d = self.directory
        return (os.path.join(d, "meta.json"), os.path.join(d, self._row_files[0]), os.path.join(d, self._row_files[1]),
                os.path.join(d, "payload.sqlite3"))

    def _open_payload_db(self) -> sqlite3.Connection:
        db = sqlite3.connect(self._paths()[3])
        db.execute("CREATE TABLE IF NOT EXISTS payload (key INTEGER PRIMARY KEY, id TEXT NOT NULL, "
                   "metadata TEXT, document TEXT)")
        db.execute("CREATE TABLE IF NOT EXISTS groups (ordinal INTEGER PRIMARY KEY, name TEXT NOT NULL)")
        db.execute("CREATE TABLE IF NOT EXISTS state (name TEXT PRIMARY KEY, value TEXT NOT NULL)")
        return db

    def persist(self) -> None:
        """Flush the collection to ``self.directory`` (incremental and crash-safe; protocol in the module docstring)."""
        if self.directory is None:
            raise ValueError("collection was created without a persist directory")
        with self._lock:
            self._lock_directory()
            meta_p, rows_p, keys_p, _ = self._paths()
            journal_p = os.path.join(self.directory, "patch.journal")
            idx = self._index
            n_rows = idx.rows() if idx is not None else 0
            generation = self._generation + 1
            patches = []
            if idx is not None:
                rb = idx.row_bytes
                # (1) rows appended since the last flush: written past the committed count, invisible until (3)
                if n_rows > self._persisted_rows:
                    with open(rows_p, "r+b" if os.path.exists(rows_p) else "w+b") as fr, \
                            open(keys_p, "r+b" if os.path.exists(keys_p) else "w+b") as fk:
                        fr.truncate(self._persisted_rows * rb)  # leftovers of a flush that never committed
                        fk.truncate(self._persisted_rows * 8)
                        fr.seek(self._persisted_rows * rb)
                        fk.seek(self._persisted_rows * 8)
                        step = max(1, (64 << 20) // rb)
                        for lo in range(self._persisted_rows, n_rows, step):
                            rows, keys = idx.export_raw(lo, min(step, n_rows - lo))
                            fr.write(rows.tobytes())
                            fk.write(keys.tobytes())
                        fr.flush()
                        fk.flush()
                        os.fsync(fr.fileno())
                        os.fsync(fk.fileno())
                # (2) rows overwritten in place or deleted since the last flush: journal first
                for r in sorted(r for r in self._dirty_rows if r < self._persisted_rows):
                    rows, keys = idx.export_raw(r, 1)
                    patches.append((r, int(keys[0]), rows[0].tobytes()))
                if patches:
                    blob = bytearray(_JOURNAL_MAGIC + struct.pack("<qqq", generation, len(patches), rb))
                    for r, key, bits in patches:
                        blob += struct.pack("<qq", r, key) + bits
                    _write_atomic(journal_p, bytes(blob))
            if not os.path.exists(meta_p):
                _write_atomic(meta_p, json.dumps({"format": FORMAT_VERSION, "name": self.name, "metadata": self.metadata,
                                                  "space": self.space, "dtype": self.dtype}).encode())
            # (3) the commit point: payload, group ordinals and the new row count in ONE transaction
            db = self._open_payload_db()
            try:
                gone = [(k,) for k in self._dirty_payload if k not in self._payload]
                live = [(k, self._payload[k]["id"],
                         json.dumps(self._payload[k]["metadata"]) if self._payload[k]["metadata"] is not None else None,
                         self._payload[k]["document"]) for k in self._dirty_payload if k in self._payload]
                if gone:
                    db.executemany("DELETE FROM payload WHERE key = ?", gone)
                if live:
                    db.executemany("INSERT OR REPLACE INTO payload (key, id, metadata, document) VALUES (?,?,?,?)", live)
                if self._dirty_groups:
                    db.executemany("INSERT OR REPLACE INTO groups (ordinal, name) VALUES (?,?)",
                                   [(o, self._group_names[o]) for o in self._dirty_groups])
                state = {"rows": n_rows, "generation": generation, "next_synthetic": self._next_synthetic,
                         "dim": self.dim if self.dim is not None else 0}
                db.executemany("INSERT OR REPLACE INTO state (name, value) VALUES (?,?)",
                               [(k, str(v)) for k, v in state.items()])
                db.commit()  # sqlite fsyncs its journal and database (synchronous = FULL)
            finally:
                db.close()
            self._generation = generation
            self._persisted_rows = n_rows
            self._dirty_payload.clear()
            self._dirty_groups.clear()
            self._dirty_rows.clear()
            # (4) apply the journalled patches, (5) drop the journal
            if patches:
                self._apply_patches(patches, idx.row_bytes)
                os.remove(journal_p)
                _fsync_file(self.directory)

    @staticmethod
    def committed_state(directory: str) -> Dict[str, Any]:
        """What the last committed flush of the collection at ``directory`` recorded (tools, tests): meta.json's static
        description plus rows / generation / next_synthetic / dim and the names of the row files."""
        with open(os.path.join(directory, "meta.json")) as f:
            out = dict(json.load(f))
        db = sqlite3.connect(os.path.join(directory, "payload.sqlite3"))
        try:
            for name, value in db.execute("SELECT name, value FROM state"):
                out[name] = value if name.endswith("_file") else int(value)
        finally:
            db.close()
        out.setdefault("rows_file", "rows.bin")
        out.setdefault("keys_file", "keys.bin")
        return out

    def _apply_patches(self, patches, rb: int) -> None:
        _, rows_p, keys_p, _ = self._paths()
        with open(rows_p, "r+b") as fr, open(keys_p, "r+b") as fk:
            for r, key, bits in patches:
                fr.seek(r * rb)
                fr.write(bits)
                fk.seek(r * 8)
                fk.write(struct.pack("<q", key))
            fr.flush()
            fk.flush()
            os.fsync(fr.fileno())
            os.fsync(fk.fileno())

    @classmethod
    def load(cls, directory: str, *, device: Optional[int] = None,
             devices: Optional[Sequence[int]] = None) -> "B200Collection":
        """Reopen a persisted collection: recover an interrupted flush, mmap the shard files, one H2D copy, payload
        from sqlite.  Rows deleted before the flush are dropped on the way in (the shard comes back compacted, in the
        same insertion order), after which the files are rewritten to match (to new names, renamed over the old)."""
        with open(os.path.join(directory, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") not in (1, FORMAT_VERSION):
            raise ValueError(f"{directory}: unknown shard format {meta.get('format')!r}")
        col = cls(meta["name"], meta.get("metadata") or {"hnsw:space": meta["space"]}, dtype=meta["dtype"],
                  device=device, directory=directory, devices=devices)
        col._lock_directory()
        try:
            return cls._load_locked(col, directory, meta)
        except BaseException:
            col.close()  # releases the directory lock
            raise

    @classmethod
    def _load_locked(cls, col: "B200Collection", directory: str, meta: dict) -> "B200Collection":
        _, rows_p, keys_p, pay_p = col._paths()
        journal_p = os.path.join(directory, "patch.journal")
        state: Dict[str, str] = {}
        if os.path.exists(pay_p):
            db = col._open_payload_db()
            try:
                state = {k: v for k, v in db.execute("SELECT name, value FROM state")}
                for key, id_str, md, doc in db.execute("SELECT key, id, metadata, document FROM payload"):
                    col._payload[int(key)] = {"id": id_str, "metadata": json.loads(md) if md is not None else None,
                                              "document": doc}
                    col._key_of_id[id_str] = int(key)
                for o, name in db.execute("SELECT ordinal, name FROM groups ORDER BY ordinal"):
                    while len(col._group_names) <= int(o):
                        col._group_names.append("")
                    col._group_names[int(o)] = name
                    col._group_ordinal[name] = int(o)
            finally:
                db.close()
        # format 1 kept the counters in meta.json
        n_rows = int(state.get("rows", meta.get("rows", 0)))
        dim = int(state.get("dim", meta.get("dim") or 0)) or None
        col._next_synthetic = int(state.get("next_synthetic", meta.get("next_synthetic", -2)))
        col._generation = int(state.get("generation", 0))
        col._row_files = (state.get("rows_file", "rows.bin"), state.get("keys_file", "keys.bin"))
        _, rows_p, keys_p, pay_p = col._paths()
        for name in os.listdir(directory):  # row files no committed state refers to (an interrupted compaction)
            if (name.startswith("rows.") or name.startswith("keys.")) and name.endswith(".bin") and name not in col._row_files:
                os.remove(os.path.join(directory, name))
        rb = (dim or 0) * (2 if col.dtype == "bf16" else 4)
        if os.path.exists(journal_p):
            blob = open(journal_p, "rb").read()
            ok = blob[:8] == _JOURNAL_MAGIC and len(blob) >= 32
            if ok:
                gen, count, jrb = struct.unpack("<qqq", blob[8:32])
                ok = gen == col._generation and jrb == rb and len(blob) == 32 + count * (16 + rb)
            if ok:  # the flush had committed: finish it
                patches = []
                for i in range(count):
                    o = 32 + i * (16 + rb)
                    r, key = struct.unpack("<qq", blob[o:o + 16])
                    patches.append((r, key, blob[o + 16:o + 16 + rb]))
                col._apply_patches(patches, rb)
            os.remove(journal_p)  # (otherwise it never committed: the row files still hold the committed bits)
        compacted = False
        if n_rows > 0 and dim:
            for path, unit in ((rows_p, rb), (keys_p, 8)):
                have = os.path.getsize(path) if os.path.exists(path) else -1
                if have < n_rows * unit:
                    raise RuntimeError(f"{path}: {have} bytes on disk, the committed state names {n_rows} rows "
                                       f"({n_rows * unit} bytes) -- the collection directory is damaged")
            idx = col._ensure_index(int(dim), reserve_rows=n_rows)
            mm_r = np.memmap(rows_p, dtype=np.uint8, mode="r", shape=(n_rows, idx.row_bytes))
            mm_k = np.memmap(keys_p, dtype=np.int64, mode="r", shape=(n_rows,))
            keys = np.array(mm_k)
            live = keys != _INT64_MIN
            if live.all():
                idx.import_raw(mm_r, keys)
            else:
                compacted = True
                sel = np.flatnonzero(live)
                step = max(1, (64 << 20) // idx.row_bytes)
                for lo in range(0, sel.size, step):
                    part = sel[lo:lo + step]
                    idx.import_raw(np.ascontiguousarray(mm_r[part]), keys[part])
            del mm_r, mm_k
        if compacted:
            col._rewrite_row_files()
        else:
            col._persisted_rows = n_rows
        return col

    def _rewrite_row_files(self) -> None:
        """After a compacting load: the compacted rows go to FRESH files, and one sqlite commit switches the committed
        state (row count, generation, file names) over to them; the old files are removed afterwards.  A crash before
        the commit leaves the old files and the old state (the next load compacts again), a crash after it the new
        ones (load() sweeps the orphans)."""
        with self._lock:
            idx = self._index
            n_rows = idx.rows()
            old = self._paths()[1:3]
            gen = self._generation + 1
            names = (f"rows.g{gen}.bin", f"keys.g{gen}.bin")
            step = max(1, (64 << 20) // idx.row_bytes)
            with open(os.path.join(self.directory, names[0]), "wb") as fr, \
                    open(os.path.join(self.directory, names[1]), "wb") as fk:
                for lo in range(0, n_rows, step):
                    rows, keys = idx.export_raw(lo, min(step, n_rows - lo))
                    fr.write(rows.tobytes())
                    fk.write(keys.tobytes())
                fr.flush()
                fk.flush()
                os.fsync(fr.fileno())
                os.fsync(fk.fileno())
            _fsync_file(self.directory)
            db = self._open_payload_db()
            try:
                db.executemany("INSERT OR REPLACE INTO state (name, value) VALUES (?,?)",
                               [("rows", str(n_rows)), ("generation", str(gen)), ("rows_file", names[0]),
                                ("keys_file", names[1])])
                db.commit()
            finally:
                db.close()
            self._generation, self._row_files = gen, names
            self._persisted_rows = n_rows
            self._dirty_rows.clear()
            for path in old:
                if os.path.exists(path):
                    os.remove(path)

    # -- document-level helpers the reference probes with hasattr (api_server.py:230-231, 267-270) --------
    def keys_where(self, field: str, value) -> List[int]:
        with self._lock:
            return [k for k, p in self._payload.items() if (p.get("metadata") or {}).get(field) == value]

    def close(self) -> None:
        with self._lock:
            if self._index is not None:
                self._index.close()
                self._index = None
            self._key_of_id.clear()
            self._payload.clear()
            self._group_ordinal.clear()
            self._group_names.clear()
            if self._lock_fd is not None:
                try:
                    fcntl.flock(self._lock_fd, fcntl.LOCK_UN)
                finally:
                    os.close(self._lock_fd)
                    self._lock_fd = None


# ---- process-global registry: (persist_dir, name) -> collection -------------------------------
_REGISTRY: Dict[tuple, B200Collection] = {}
_REGISTRY_LOCK = threading.Lock()


class B200Client:
    """Stand-in for ``chromadb.PersistentClient(path=...)`` (chroma_child_store.py:32)."""

    def __init__(self, path: str = "."):
        self.path = os.path.abspath(path)

    def _dir_of(self, name: str) -> str:
        return os.path.join(self.path, f"{name}.b200")

    def get_or_create_collection(self, name: str, metadata: Optional[Dict[str, Any]] = None, **kw) -> B200Collection:
        key = (self.path, name)
        with _REGISTRY_LOCK:
            col = _REGISTRY.get(key)
            if col is None:
                d = self._dir_of(name)
                if os.path.exists(os.path.join(d, "meta.json")):
                    col = B200Collection.load(d, device=kw.get("device"), devices=kw.get("devices"))  # restart: mmap + H2D
                else:
                    col = B200Collection(name, metadata, directory=d, **kw)
                _REGISTRY[key] = col
            return col

    def get_collection(self, name: str) -> B200Collection:
        with _REGISTRY_LOCK:
            col = _REGISTRY.get((self.path, name))
            if col is None and os.path.exists(os.path.join(self._dir_of(name), "meta.json")):
                col = B200Collection.load(self._dir_of(name))
                _REGISTRY[(self.path, name)] = col
        if col is None:
            raise ValueError(f"Collection {name} does not exist.")
        return col

    def list_collections(self) -> List[B200Collection]:
        with _REGISTRY_LOCK:
            return [c for (p, _), c in _REGISTRY.items() if p == self.path]

    def delete_collection(self, name: str) -> None:
        import shutil

        with _REGISTRY_LOCK:
            col = _REGISTRY.pop((self.path, name), None)
        if col is not None:
            col.close()
        shutil.rmtree(self._dir_of(name), ignore_errors=True)


def PersistentClient(path: str = ".") -> B200Client:  # noqa: N802 - chromadb's spelling
    return B200Client(path)


def reset_registry() -> None:
    """Drop every collection from memory (tests; also simulates a process restart -- persisted
    collections are reloaded from their directory by the next get_or_create_collection)."""
    with _REGISTRY_LOCK:
        cols = list(_REGISTRY.values())
        _REGISTRY.clear()
    for c in cols:
        c.close()
