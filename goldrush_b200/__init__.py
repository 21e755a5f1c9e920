"""goldrush_b200 — B200-native engine for GoldRush-Path's read-selection loop.

The product is the C-ABI shared library ``goldrush_b200/_lib/libgoldrush_b200.so`` (CUDA kernels for
sm_100a + host C++) and the ``build/goldrush-path`` executable.  This package is the thin ctypes
mirror of ``include/goldrush_b200.h`` that tests and ``bench.py`` use; it contains no algorithm and
no CPU fallback: if the library is missing or no CUDA device is usable, calls raise.
"""
from . import api  # noqa: F401
from .api import (  # noqa: F401
    Engine,
    GrbError,
    Params,
    RunOptions,
    RunResult,
    SynthParams,
    calc_optimal_size,
    default_hash_universe,
    lib,
    lib_path,
    make_seed_pattern,
    phred_finalize,
    run_path,
    run_two_stage,
    synth_fastq,
    synth_fastq_raw,
    free_host,
    synth_num_reads,
)
