"""Rig JSON helpers for tests and bench (the C++ apps have their own parser in csrc/host)."""
import json


def load_rig(path):
    with open(path) as f:
        return json.load(f)


def save_rig(rig, path):
    with open(path, "w") as f:
        json.dump(rig, f, indent=1)


def filter_destinations(ids, destinations):
    """image_util::filterDestinations (ImageUtil.cpp:110-125): indices of the requested camera ids,
    in the order requested; empty string = all cameras."""
    if not destinations:
        return list(range(len(ids)))
    out = []
    for dest in destinations.split(","):
        for i, cid in enumerate(ids):
            if cid == dest:
                out.append(i)
    return out
