for N in 148 111 74 37; do EA_SOLVE_MAX_CTAS=$N bash tools/sweep_cluster.sh 592 1 | sed "s/^/ctas $N /"; done
