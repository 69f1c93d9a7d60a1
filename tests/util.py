import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def dict_from_json(j):
    """inverse of make_golden.jsonable"""
    if j.get("invalid") == 1:
        return {"invalid": 1}
    out = {"invalid": 0}
    for k in "ACGTN":
        out[k] = set(j[k])
    out["M"] = {int(k): v for k, v in j["M"].items()}
    return out


def lists_from_dict(d):
    """compressed dict -> (pos int32, off int32[7], invalid) without touching the product's packing module"""
    if d.get("invalid") == 1:
        return np.empty(0, np.int32), np.zeros(7, np.int32), True
    parts = [sorted(d[k]) for k in "ACGTN"] + [sorted(d["M"].keys())]
    off = np.zeros(7, np.int32)
    off[1:] = np.cumsum([len(p) for p in parts])
    pos = np.array([x for p in parts for x in p], dtype=np.int32)
    return pos, off, False
