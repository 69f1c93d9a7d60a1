"""`from seqComparer import seqComparer` (src/findNeighbour3-server.py:74) resolves here when this directory precedes the
reference's src/ on PYTHONPATH: the server runs unmodified on the B200 engine."""
from findneighbour3_b200.seqComparer import seqComparer  # noqa: F401
