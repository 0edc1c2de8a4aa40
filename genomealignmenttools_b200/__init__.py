"""B200-native chain rescoring: drop-in for the chainCalcScore hot path of
hillerlab/GenomeAlignmentTools (scoreChain, chainNet -rescore, chainCleaner).

The product is the CUDA C-ABI library ``libgat.so`` (include/gat.h); this package is the thin
host-side plumbing around it (ctypes binding, .2bit/.chain readers, work-list builders) that the
tests and bench.py use.  There is no CPU scoring path in here.
"""
from .records import BLOCK_DTYPE, JOB_DTYPE, NRUN_DTYPE, QSEQ_MINUS, BLOCK_JOINED, NO_CLIP_START, NO_CLIP_END
from .twobit import PackedGenome
from .scoring import ScoreScheme, GapCalc, Scoring
from .engine import ChainScorer, GatError

__all__ = [
    "BLOCK_DTYPE", "JOB_DTYPE", "NRUN_DTYPE", "QSEQ_MINUS", "BLOCK_JOINED", "NO_CLIP_START", "NO_CLIP_END",
    "PackedGenome", "ScoreScheme", "GapCalc", "Scoring", "ChainScorer", "GatError",
]
