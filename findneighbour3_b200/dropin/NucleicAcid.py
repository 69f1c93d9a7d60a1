"""`from NucleicAcid import NucleicAcid` (src/findNeighbour3-server.py:72)."""
from findneighbour3_b200.NucleicAcid import NucleicAcid  # noqa: F401
