"""Measured ceilings for the scattered 32-byte sector traffic of the particle passes (scb_debug_l2_probe):
the gather's lane-pair record loads and the tile deposit's four-lane fp64 reductions at pseudo-random
addresses, over buffers from L2-resident to the size of the real packed field / tile accumulator.
usage: python tools/l2_probe.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def probe(hd, mode, nbytes, iters=400):
    out = C.c_double(0.0)
    hd.check(hd.lib.scb_debug_l2_probe(hd.h, mode, int(nbytes), iters, C.byref(out)))
    return out.value


def main():
    scb = load_package()
    hd = scb.default_handle(0)
    rows = []
    for mode, name in ((0, "sector_reads"), (1, "sector_reductions")):
        for mb in (32, 64, 100, 256, 537):
            r = probe(hd, mode, mb * 1e6)
            rows.append({"probe": name, "buffer_MB": mb, "Gsectors_per_s": round(r / 1e9, 2), "TBps": round(r * 32 / 1e12, 3)})
            print(rows[-1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "l2_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
