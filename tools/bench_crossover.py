#!/usr/bin/env python
"""gat_crossover (cBlockFindCrossover, kent/src/lib/chainConnect.c:61-105) timed on a GPU: pairs per second for short
overlaps (what chainRemovePartialOverlaps mostly meets) and long ones, next to the unmodified reference function on one
host core (oracle/_ref/libkentref.so) on a sample of the same pairs.  Prints one JSON line per case."""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from genomealignmenttools_b200 import ChainScorer, Scoring, synth  # noqa: E402
from genomealignmenttools_b200.records import XPAIR_DTYPE  # noqa: E402


def main():
    import oracle_lib
    import make_golden_helpers as helpers
    import tempfile
    t_names, q_names = ["chrA"], ["chrX"]
    w = synth.make_workload(t_names, [50_000_000], q_names, [45_000_000], 200_000, seed=5, telomere_n=1000, n_fraction=0.001)
    rng = np.random.default_rng(1)
    sc = ChainScorer(0)
    sc.load_genome("t", w.t); sc.load_genome("q", w.q); sc.set_scoring(Scoring(None, "loose"))
    d = tempfile.mkdtemp()
    paths = helpers.write_case(w, t_names, q_names, d)
    ref = oracle_lib.load_ref()
    oracle = oracle_lib.load()
    osc, tg, qg = oracle.scoring(None, "loose"), oracle.genome(paths["t"]), oracle.genome(paths["q"])
    for name, n, lo, hi in (("short overlaps 1..32 bp", 2_000_000, 1, 33), ("overlaps 1..64 bp", 2_000_000, 1, 65),
                            ("long overlaps 65..2000 bp", 200_000, 65, 2001)):
        ov = rng.integers(lo, hi, n)
        pairs = np.zeros(n, dtype=XPAIR_DTYPE)
        pairs["leftTEnd"] = rng.integers(3000, 49_000_000, n); pairs["leftQEnd"] = rng.integers(3000, 44_000_000, n)
        pairs["rightTStart"] = rng.integers(0, 49_000_000, n); pairs["rightQStart"] = rng.integers(0, 44_000_000, n)
        pairs["overlap"] = ov
        sc.crossover(pairs)
        t0 = time.time()
        reps = 5
        for _ in range(reps):
            pos, adj = sc.crossover(pairs)
        dt = (time.time() - t0) / reps
        k = min(n, 50_000)
        cases = [tuple(int(pairs[f][i]) for f in ("leftTEnd", "leftQEnd", "rightTStart", "rightQStart", "overlap")) for i in range(k)]
        t0 = time.time()
        opos, oadj = oracle.crossover(osc, tg, qg, 0, 0, "+", cases)
        cpu_dt = time.time() - t0
        bad = int((pos[:k] != opos).sum() + (adj[:k] != oadj).sum())
        print(json.dumps({"case": name, "pairs": n, "overlap_bases": int(ov.sum()), "gpu_ms_per_call_incl_copies": round(dt * 1e3, 3),
                          "gpu_mpairs_per_s": round(n / dt / 1e6, 2), "gpu_gbases_per_s": round(ov.sum() / dt / 1e9, 3),
                          "cpu_port_1core_mpairs_per_s": round(k / cpu_dt / 1e6, 3), "checked_pairs": k, "mismatches": bad}))
    sc.close()


if __name__ == "__main__":
    main()
