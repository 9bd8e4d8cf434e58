cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_graphcut.py -m gpu -x -q 2>&1 | tail -3
for T in 1 0; do
  if [ $T = 1 ]; then export PXB_MF_TOP_DOWN=1; else unset PXB_MF_TOP_DOWN; fi
  echo "== top_down_only=$T: $(timeout 300 python tools/profile_fit.py 10000 0.05 2>&1 | grep 'N=' ) $(timeout 300 python tools/profile_fit.py 5000 0.3 2>&1 | grep 'N=' ) $(timeout 300 python tools/profile_fit.py 2000 0.05 2>&1 | grep 'N=' )"
done
PXB_MF_STATS=1 timeout 300 python tools/profile_fit.py 10000 0.05 2>&1 | grep "labelling:" | tail -4
timeout 300 python tools/profile_fit.py 10000 0.05 2>&1 | grep -E "lo_labeling|pearl" | tail -2
