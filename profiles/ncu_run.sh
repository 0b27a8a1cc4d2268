# ncu captures of the k-NN kernels (run on the GPU box): --set full of one launch of each launch shape, then the launch list
cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:icp_knn_kernel -s 35 -c 1 \
  -o gpurun_out/prof_knn_single_r2 -f python profiles/one_align.py 2 0 > gpurun_out/ncu_knn_single.log 2>&1
tail -2 gpurun_out/ncu_knn_single.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_final.csv \
  python profiles/one_align.py 3 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
