from findneighbour3_b200.seqComparer import seqComparer  # noqa: F401 -- `from seqComparer import seqComparer` (findNeighbour3-server.py:74) lands here
