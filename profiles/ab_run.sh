cd $GRAFT_REPO_ROOT
for v in pk256 pp pk; do
  SM_B200_LIB=$PWD/staticmapping_b200/libsm_b200_$v.so timeout 300 python -m pytest tests/test_gpu_icp.py -x -q -m gpu 2>&1 | tail -2
done
timeout 600 bash profiles/ab_lib.sh libsm_b200_base.so libsm_b200_pk.so libsm_b200_pk256.so libsm_b200_pp.so libsm_b200_base.so
