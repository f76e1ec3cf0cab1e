#!/bin/bash
# Run the GPU test files one by one under a hard timeout so a hung kernel cannot eat the whole
# gpurun budget; logs land in gpurun_out/.
mkdir -p gpurun_out
rc_all=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout --signal=KILL 600 python -m pytest "$f" -x -q -m gpu --timeout=550 > "gpurun_out/${name}.log" 2>&1
  rc=$?
  echo "== $f -> rc=$rc"
  tail -n 25 "gpurun_out/${name}.log"
  [ $rc -ne 0 ] && rc_all=$rc
done
exit $rc_all
