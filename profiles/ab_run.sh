cd $GRAFT_REPO_ROOT
bash profiles/ncu_run.sh
timeout 400 bash profiles/ab_pipelines.sh "16 2" "24 2" "32 2" "24 3"
SM_B200_KNN_QPC=768 timeout 100 bash profiles/ab_pipelines.sh "16 2"
