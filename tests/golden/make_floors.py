"""Regenerates tests/golden/floors.json: the reference's own float-parity floors (tests/floors.py) for every configuration / stage the GPU
tests gate against. Needs the compiled reference (oracle/_ref, i.e. this container); ~3 minutes of CPU.

    python tests/golden/make_floors.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import floors  # noqa: E402

CONFIGS = ["metop_ahrpt", "bpsk_half", "jpss_hrd", "dvbs2_front", "hrpt_bpsk", "psk8", "metop_oversampled", "bpsk_decim8"]


def main():
    out = {}
    for name in CONFIGS:
        out[floors._key("chain", name, 21, None, 1e-6, ())] = floors.chain_floor(name, 21)
        out[floors._key("stage", name, 21, "mm", 1e-6, ())] = floors.stage_floor(name, 21, "mm")
        if name != "dvbs2_front":
            out[floors._key("stage", name, 21, "costas", 1e-6, ())] = floors.stage_floor(name, 21, "costas")
        print(name, "done", flush=True)
    gx = (("clock_recovery", "gardner"),)
    for name in ("metop_ahrpt", "bpsk_half"):
        out[floors._key("stage", name, 21, "mm", 1e-6, gx)] = floors.stage_floor(name, 21, "mm", extra=gx)
        out[floors._key("chain", name, 21, None, 1e-6, gx)] = floors.chain_floor(name, 21, extra=gx)
    ex = (("post_costas_dc", True),)
    out[floors._key("chain", "bpsk_half", 20, None, 1e-6, ex)] = floors.chain_floor("bpsk_half", 20, extra=ex)
    with open(floors._CACHE_PATH, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(out), "entries ->", floors._CACHE_PATH)


if __name__ == "__main__":
    main()
