"""Regenerates tests/golden/floors.json: the reference's own float-parity floors (tests/floors.py) for every configuration / stage the GPU
tests gate against, for the synthetic signals as this container generates them (each entry carries the CRC of its signal). Needs the
compiled reference (oracle/_ref, i.e. this container); ~3 minutes of CPU on 8 cores.

    python tests/golden/make_floors.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import floors  # noqa: E402


def main():
    floors._cache = {}
    n = floors.warm()
    with open(floors._CACHE_PATH, "w") as f:
        json.dump(floors._cache, f, indent=1, sort_keys=True)
    print(n, "entries ->", floors._CACHE_PATH)


if __name__ == "__main__":
    main()
