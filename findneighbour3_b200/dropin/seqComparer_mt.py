from findneighbour3_b200.seqComparer_mt import seqComparer  # noqa: F401 -- `from seqComparer_mt import seqComparer` (the threaded variant named at findNeighbour3-server.py:74) lands here
