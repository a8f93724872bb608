for pf in 0 200 800 1500 3000; do export PYSTL_B200_PREFETCH=$pf; echo "== prefetch $pf"; WORKLOADS="cfg2 cfg5" bash gpu_sweep.sh; done
