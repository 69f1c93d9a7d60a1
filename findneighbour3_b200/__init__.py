"""findneighbour3_b200 — B200-native masked pairwise SNV distance engine behind
findNeighbour3's `seqComparer` API.  See DESIGN.md."""
__version__ = "0.1.0"
