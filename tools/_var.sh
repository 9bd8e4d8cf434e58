cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_graphcut.py -m gpu -x -q 2>&1 | tail -3
echo "$(timeout 300 python tools/profile_fit.py 10000 0.05 2>&1 | grep 'N=' ) $(timeout 300 python tools/profile_fit.py 5000 0.3 2>&1 | grep 'N=' ) $(timeout 300 python tools/profile_fit.py 2000 0.05 2>&1 | grep 'N=' )"
PXB_MF_STATS=2 timeout 300 python tools/profile_fit.py 10000 0.05 2>&1 | grep "pxb expansion\] alpha" | sort -t= -k6 -n | tail -3
timeout 300 python tools/profile_fit.py 10000 0.05 2>&1 | grep -E "lo_labeling|pearl" | tail -2
