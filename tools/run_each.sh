#!/bin/bash
# run each GPU test id under its own hard timeout so one hung kernel cannot eat the whole call
for t in "$@"; do
  timeout -s KILL ${TMO:-100} python -m pytest "$t" -x -q -m gpu 2>&1 | tail -3
  echo "== $t rc=${PIPESTATUS[0]}"
done
