"""Per-region DP work (bk_batch_result.region_dp_cells) next to the static features a host knows before the device
pass, for fitting shard.region_cost.  Usage (GPU box): python tools/cost_model_data.py C2:4000 C5:20000 C3:500 C4:100"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from breakmer_b200 import _lib, batch, synth          # noqa: E402


def main():
    h = _lib.Handle(0)
    rows = []
    for spec in sys.argv[1:]:
        cfg, _, n = spec.partition(":")
        n = int(n)
        for a in range(0, n, 2000):
            regions = [synth.config_region(cfg, i) for i in range(a, min(n, a + 2000))]
            out = batch.run(h, batch.PackedBatch(regions))
            for j, r in enumerate(regions):
                rows.append({"cfg": cfg, "i": a + j, "n_reads": len(r.reads), "read_bases": sum(len(x[1]) for x in r.reads),
                             "sc_bases": sum(len(x[1]) for x in r.sc_records), "ref_len": len(r.ref_fwd),
                             "uniq": int(out.uniq_reg_off[j + 1] - out.uniq_reg_off[j]),
                             "so": int(out.so_off[j + 1] - out.so_off[j]), "cells": int(out.region_dp_cells[j]),
                             "contigs": int(out.ctg_reg_off[j + 1] - out.ctg_reg_off[j])})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "cost_model_data.json"), "w") as f:
        json.dump(rows, f)
    print(len(rows), "rows")


if __name__ == "__main__":
    main()
