"""Shared by tests: write a synthetic Workload out as .2bit + .chain files (TEST INFRASTRUCTURE)."""
import os
import numpy as np
from genomealignmenttools_b200 import chainio


def chain_headers(w, t_names, q_names):
    counts = np.diff(np.append(w.jobs["blockPtr"].astype(np.int64), w.total))
    b = w.blocks
    heads = []
    for j, job in enumerate(w.jobs):
        fb, nb = int(job["firstBlock"]), int(counts[j])
        last = b[fb + nb - 1]
        ts, qs = int(job["tSeq"]), int(job["qSeq"] & 0x7FFFFFFF)
        heads.append((0, t_names[ts], int(w.t.sizes[ts]), int(b[fb]["tStart"]), int(last["tStart"]) + int(last["size"]),
                      q_names[qs], int(w.q.sizes[qs]), "-" if job["qSeq"] >> 31 else "+",
                      int(b[fb]["qStart"]), int(last["qStart"]) + int(last["size"]), j + 1))
    return heads, counts


def write_genomes(w, d):
    d = str(d)
    os.makedirs(d, exist_ok=True)
    paths = {"t": os.path.join(d, "t.2bit"), "q": os.path.join(d, "q.2bit"), "chain": os.path.join(d, "in.chain")}
    w.t.write_2bit(paths["t"])
    w.q.write_2bit(paths["q"])
    return paths


def write_case(w, t_names, q_names, d):
    paths = write_genomes(w, d)
    heads, counts = chain_headers(w, t_names, q_names)
    chainio.write_chains(paths["chain"], heads, w.blocks, w.jobs["firstBlock"], counts)
    return paths
