#!/usr/bin/env python
"""Tool-level wall clock: the unmodified reference scoreChain (oracle/_ref/scoreChain, one host core, as shipped)
against the drop-in bin/scoreChain on the SAME files in the same run ("end-to-end tool wall").
BASELINE.json configs[0]-style input: chains of hg38 chr1 against mm10 chromosomes, synthetic .2bit genomes at the
real chromosome sizes, default matrix, -linearGap=medium.  Prints one JSON line; the output files must be identical.

usage: tools/cli_bench.py [--blocks 2000000] [--qchroms 8] [--keep DIR]"""
import argparse
import filecmp
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from genomealignmenttools_b200 import synth  # noqa: E402
import make_golden_helpers as helpers  # noqa: E402


def write_chains_fast(path, heads, blocks, first, counts):
    """chainWrite (chain.c:211-227), vectorised: one text blob per chain would be slow at millions of blocks."""
    ts = blocks["tStart"].astype(np.int64); qs = blocks["qStart"].astype(np.int64); sz = blocks["size"].astype(np.int64)
    dt = np.zeros(len(blocks), dtype=np.int64); dq = np.zeros(len(blocks), dtype=np.int64)
    dt[:-1] = ts[1:] - (ts[:-1] + sz[:-1]); dq[:-1] = qs[1:] - (qs[:-1] + sz[:-1])
    with open(path, "w") as f:
        for c, h in enumerate(heads):
            f.write("chain %.0f %s %d + %d %d %s %d %s %d %d %d\n" % ((h[0],) + tuple(h[1:])))
            fb, nb = int(first[c]), int(counts[c])
            if nb > 1:
                body = np.stack([sz[fb:fb + nb - 1], dt[fb:fb + nb - 1], dq[fb:fb + nb - 1]], axis=1)
                f.write("\n".join("%d\t%d\t%d" % tuple(r) for r in body.tolist()))
                f.write("\n")
            f.write("%d\n\n" % sz[fb + nb - 1])


def timed(cmd):
    t0 = time.time()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    dt = time.time() - t0
    if r.returncode != 0:
        sys.stderr.write(r.stderr.decode()[-2000:])
        raise SystemExit("%s failed with %d" % (cmd[0], r.returncode))
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=2_000_000)
    ap.add_argument("--qchroms", type=int, default=8, help="mm10 chromosomes the chains land on")
    ap.add_argument("--keep", default=None)
    ap.add_argument("--tools", default="scoreChain", help="comma list of scoreChain,chainNet,chainCleaner (configs[0..2] of BASELINE.json)")
    ap.add_argument("--gpus-list", default="", help="comma list of N: also time bin/scoreChain -gpus=N (N above the box's GPU count "
                    "runs several contexts per device through GAT_DEVICES) and report each shard's host->device bytes")
    args = ap.parse_args()
    ex = os.path.join(ROOT, "tests", "golden", "example")
    tn, ts = synth.read_chrom_sizes(os.path.join(ex, "hg38.chrom.sizes"))
    qn, qs = synth.read_chrom_sizes(os.path.join(ex, "mm10.chrom.sizes"))
    tn, ts = tn[:1], ts[:1]                                   # chr1
    qn, qs = qn[:args.qchroms], qs[:args.qchroms]
    t0 = time.time()
    w = synth.make_workload(tn, ts, qn, qs, args.blocks, seed=0x5EED0010, telomere_n=10000, n_fraction=0.001)
    d = args.keep or tempfile.mkdtemp(prefix="gat_cli_")
    os.makedirs(d, exist_ok=True)
    paths = helpers.write_genomes(w, d)
    heads, counts = helpers.chain_headers(w, tn, qn)
    write_chains_fast(paths["chain"], heads, w.blocks, w.jobs["firstBlock"], counts)
    sys.stderr.write("inputs written in %.1f s: %d chains, %d blocks, %.1f Mbp aligned\n"
                     % (time.time() - t0, len(w.jobs), w.total, w.aligned_bp / 1e6))
    ref = os.path.join(ROOT, "oracle", "_ref", "scoreChain")
    ours = os.path.join(ROOT, "bin", "scoreChain")
    common = [paths["chain"], paths["t"], paths["q"]]
    out_ref, out_ours = os.path.join(d, "ref.chain"), os.path.join(d, "ours.chain")
    t_ref = min(timed([ref] + common + [out_ref, "-linearGap=medium"]) for _ in range(2))
    # CUDA context creation (0.3 .. 2 s on a box without nvidia-persistenced) is part of every run of ours
    t_ours = min(timed([ours] + common + [out_ours, "-linearGap=medium"]) for _ in range(3))
    env = dict(os.environ, GAT_TOOL_TIMING="1")
    phases = subprocess.run([ours] + common + [out_ours, "-linearGap=medium"], stderr=subprocess.PIPE, env=env).stderr.decode()
    sys.stderr.write(phases)
    phase_s = {" ".join(l.split()[1:-2]): float(l.split()[-2]) for l in phases.splitlines() if l.startswith("[timing]")}
    same = filecmp.cmp(out_ref, out_ours, shallow=False)
    print(json.dumps({"tool": "scoreChain", "config": "hg38 chr1 x mm10 (%d chromosomes), synthetic 2bit, default matrix, linearGap medium" % args.qchroms,
                      "chains": int(len(w.jobs)), "blocks": int(w.total), "aligned_mbp": round(w.aligned_bp / 1e6, 1),
                      "reference_wall_s": round(t_ref, 2), "reference_cores": 1, "ours_wall_s": round(t_ours, 2),
                      "speedup": round(t_ref / t_ours, 2), "outputs_identical": bool(same), "ours_phases_s": phase_s}))
    if not same:
        raise SystemExit("outputs differ")
    if args.gpus_list:
        import torch
        have = torch.cuda.device_count()
        for n in [int(x) for x in args.gpus_list.split(",")]:
            env = dict(os.environ, GAT_TOOL_TIMING="1")
            if n > have:
                env["GAT_DEVICES"] = ",".join(str(i % have) for i in range(n))
            out_n = os.path.join(d, "ours.gpus%d.chain" % n)
            cmd = [ours] + common + [out_n, "-linearGap=medium", "-gpus=%d" % n]
            walls, err = [], ""
            for _ in range(3):
                t0 = time.time()
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
                walls.append(time.time() - t0)
                if r.returncode != 0:
                    sys.stderr.write(r.stderr.decode()[-2000:])
                    raise SystemExit("scoreChain -gpus=%d failed" % n)
                err = r.stderr.decode()
            shards = [l for l in err.splitlines() if l.startswith("gpu shard")]
            h2d = [int(l.split(" bytes host->device")[0].split()[-1]) for l in shards]
            host_s = [[float(x.split()[-2]) for x in l.split("host:")[1].split(",")] for l in shards if "host:" in l]
            phase_s = {" ".join(l.split()[1:-2]): float(l.split()[-2]) for l in err.splitlines() if l.startswith("[timing]")}
            print(json.dumps({"tool": "scoreChain -gpus=%d" % n, "devices": env.get("GAT_DEVICES", "0..%d" % (n - 1)), "blocks": int(w.total),
                              "ours_wall_s": round(min(walls), 2), "outputs_identical": bool(filecmp.cmp(out_ref, out_n, shallow=False)),
                              "h2d_bytes_per_shard": h2d, "h2d_bytes_total": sum(h2d),
                              "host_s_per_shard_build_stage_score": host_s, "phases_s": phase_s}))
    tools = args.tools.split(",")
    if "chainNet" in tools or "chainCleaner" in tools:
        refdir, ourdir = os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "bin")
        for name, names, sizes in (("t.sizes", tn, ts), ("q.sizes", qn, qs)):
            with open(os.path.join(d, name), "w") as f:
                f.write("".join("%s\t%d\n" % p for p in zip(names, sizes)))
        srt = os.path.join(d, "sorted.chain")
        subprocess.check_call([os.path.join(refdir, "chainSort"), out_ref, srt])
        sizes = [os.path.join(d, "t.sizes"), os.path.join(d, "q.sizes")]
    if "chainNet" in tools:      # configs[1]: chainNet -rescore, every partial fill rescored
        common = ["-rescore", "-linearGap=medium", "-tNibDir=" + paths["t"], "-qNibDir=" + paths["q"], srt] + sizes
        nets = {k: [os.path.join(d, "%s.%s.net" % (k, x)) for x in "tq"] for k in ("ref", "our")}
        t_ref = timed([os.path.join(refdir, "chainNet")] + common + nets["ref"])
        t_ours = min(timed([os.path.join(ourdir, "chainNet")] + common + nets["our"]) for _ in range(2))
        sys.stderr.write(subprocess.run([os.path.join(ourdir, "chainNet")] + common + nets["our"], stderr=subprocess.PIPE,
                                        env=dict(os.environ, GAT_TOOL_TIMING="1")).stderr.decode())
        same = all(filecmp.cmp(a, b, shallow=False) for a, b in zip(nets["ref"], nets["our"]))
        fills = sum(1 for l in open(nets["ref"][0]) if l.lstrip().startswith("fill"))
        print(json.dumps({"tool": "chainNet -rescore", "chains": int(len(w.jobs)), "blocks": int(w.total), "t_net_fills": fills,
                          "reference_wall_s": round(t_ref, 2), "reference_cores": 1, "ours_wall_s": round(t_ours, 2),
                          "speedup": round(t_ref / t_ours, 2), "outputs_identical": bool(same)}))
        if not same:
            raise SystemExit("nets differ")
    if "chainCleaner" in tools:  # configs[2]: chainCleaner, linearGap loose, net made like its internal system() call
        env = dict(os.environ, PATH=refdir + ":" + os.environ["PATH"])
        subprocess.check_call("chainNet -minScore=0 sorted.chain t.sizes q.sizes stdout /dev/null | NetFilterNonNested.perl /dev/stdin "
                              "-minScore1 3000 > in.net", shell=True, executable="/bin/bash", cwd=d, env=env, stderr=subprocess.DEVNULL)
        res = {}
        for k, bindir in (("ref", refdir), ("our", ourdir)):
            env = dict(os.environ, PATH=bindir + ":" + os.environ["PATH"])
            cmd = [os.path.join(bindir, "chainCleaner"), "sorted.chain", "t.2bit", "q.2bit", k + ".clean.chain", k + ".clean.bed",
                   "-net=in.net", "-linearGap=loose"]
            t0 = time.time()
            r = subprocess.run(cmd, cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            res[k] = time.time() - t0
            if r.returncode != 0:
                sys.stderr.write(r.stderr.decode()[-2000:])
                raise SystemExit("chainCleaner (%s) failed" % k)
        same = all(filecmp.cmp(os.path.join(d, "ref.clean." + x), os.path.join(d, "our.clean." + x), shallow=False) for x in ("chain", "bed"))
        removed = sum(1 for _ in open(os.path.join(d, "ref.clean.bed")))
        print(json.dumps({"tool": "chainCleaner", "chains": int(len(w.jobs)), "blocks": int(w.total), "removed_suspects": removed,
                          "reference_wall_s": round(res["ref"], 2), "reference_cores": 1, "ours_wall_s": round(res["our"], 2),
                          "speedup": round(res["ref"] / res["our"], 2), "outputs_identical": bool(same)}))
        if not same:
            raise SystemExit("chainCleaner outputs differ")


if __name__ == "__main__":
    main()
