#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for D in 0 1 2 4 3 6 5; do
  echo "dbg=$D: $(VSEARCH_B200_DENSE_DBG=$D timeout 200 python scripts/prof_dense.py 2>&1 | tail -1)"
done
echo "single: $(VSEARCH_B200_DENSE_PAIR=0 timeout 200 python scripts/prof_dense.py 2>&1 | tail -1)"
