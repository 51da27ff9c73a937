#!/bin/bash
# smoke() alone (build() first: must find the shipped library up to date, not recompile)
mkdir -p gpurun_out
( time timeout -k 5 100 python -u -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/r2y_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r2y_smoke.log
tail -8 gpurun_out/r2y_smoke.log
