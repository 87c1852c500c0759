"""Drop-in for parent_child/vector_store_factory.py:1-12.

The reference's factory is hard-wired to Chroma ("Permanently use Chroma backend") although its
README.md:44 advertises ``CHILD_VECTOR_BACKEND``.  This factory keeps the signature and, like the
reference, ignores ``table``; it returns the B200 store.  ``CHILD_VECTOR_BACKEND=chroma`` hands back
the reference's own ChromaChildStore when that module is importable (side-by-side comparison in a
reference checkout), never silently: any other value than ``b200`` / ``chroma`` is an error.
"""
from __future__ import annotations

import os


def get_child_vector_store(collection: str | None = None, table: str | None = None):
    """Return a child vector store instance.

    Parameters:
    - collection: name for backends that support named collections
    - table: accepted for signature compatibility (pgvector backend of the reference), ignored
    """
    backend = os.getenv("CHILD_VECTOR_BACKEND", "b200").strip().lower()
    if backend == "b200":
        from .child_store import B200ChildStore

        return B200ChildStore(collection=collection)
    if backend == "chroma":
        from parent_child.chroma_child_store import ChromaChildStore  # the reference's own class

        return ChromaChildStore(collection=collection)
    raise ValueError(f"CHILD_VECTOR_BACKEND={backend!r}: expected 'b200' or 'chroma'")
