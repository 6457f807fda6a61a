#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== sanitize"; python tests/sanitize_smoke.py 2>&1 | tail -2
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_smoke.py 2>&1 | tail -12; echo "sanitizer rc=${PIPESTATUS[0]}"
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -3
echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
} > gpurun_out/t47.txt 2>&1
tail -30 gpurun_out/t47.txt
