from findneighbour3_b200.NucleicAcid import NucleicAcid  # noqa: F401 -- `from NucleicAcid import NucleicAcid` (findNeighbour3-server.py:72) lands here
