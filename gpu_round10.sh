W=cfg1xb TAG=r1m KERNEL=pystl_fused_value_and_grad FLAGS="" bash gpu_ncu.sh
W=cfg2 TAG=r1m KERNEL=pystl_fused_value_and_grad FLAGS="" bash gpu_ncu.sh
