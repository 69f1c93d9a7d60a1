"""`from seqComparer_mt import seqComparer` (the threaded variant the comment at src/findNeighbour3-server.py:74 names)."""
from findneighbour3_b200.seqComparer_mt import seqComparer  # noqa: F401
