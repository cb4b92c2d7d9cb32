#!/bin/bash
timeout -s KILL 800 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout -s KILL 300 python tools/e2e_breakdown.py 2>&1 | tail -3
