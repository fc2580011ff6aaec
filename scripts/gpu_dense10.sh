#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for D in 0 8 1; do
  echo "pair dbg=$D: $(VSEARCH_B200_DENSE_DBG=$D timeout 200 python scripts/prof_dense.py 2>&1 | tail -1)"
  echo "single dbg=$D: $(VSEARCH_B200_DENSE_PAIR=0 VSEARCH_B200_DENSE_DBG=$D timeout 200 python scripts/prof_dense.py 2>&1 | tail -1)"
done
