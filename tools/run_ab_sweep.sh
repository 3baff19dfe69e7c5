cd $GRAFT_REPO_ROOT
bash tools/ab.sh 2>&1 | cut -c1-120
bash tools/sweep.sh "$@"
