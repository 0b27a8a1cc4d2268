# parity of the product build, then profiles/ab_lib.sh over A/B builds
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_icp.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -3
timeout 400 bash profiles/ab_lib.sh libsm_b200_2p4.so libsm_b200.so
