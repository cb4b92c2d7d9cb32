#!/bin/bash
# MHRS with the separate replay kernel: parity tests, then timed sweeps (per-phase device timers, tail round trace)
timeout -s KILL 1200 python -m pytest tests/test_mhrs_gpu.py tests/test_edges_gpu.py tests/test_chain_gpu.py tests/test_golden_gpu.py tests/test_pi_gpu.py -x -q -m gpu 2>&1 | tail -6
TRACE=1 timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -18 | cut -c1-700
TRACE=1 timeout -s KILL 200 python tools/prof_run.py MHRS 1.25e6 5 2>&1 | tail -16 | cut -c1-700
timeout -s KILL 200 python tools/prof_run.py MHRS 2.5e6 3 5 2>&1 | tail -1 | cut -c1-700
