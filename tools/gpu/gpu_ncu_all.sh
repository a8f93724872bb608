# launch lists + one full capture per headline kernel (final state of round 1); the captured launch is a
# kernel-only one (no fused mean), the launch bench.py's roofline is computed from
W=cfg2 TAG=r1z SKIP=8 KERNEL=pystl_fused_value_and_grad FLAGS="" bash tools/gpu/gpu_ncu.sh
W=cfg2 TAG=r1z_fwd SKIP=5 KERNEL=pystl_fused_forward FLAGS="--forward-only" bash tools/gpu/gpu_ncu.sh
W=cfg3 TAG=r1z SKIP=8 KERNEL=pystl_fused_value_and_grad FLAGS="" bash tools/gpu/gpu_ncu.sh
W=cfg1xb TAG=r1z SKIP=8 KERNEL=pystl_fused_value_and_grad FLAGS="" bash tools/gpu/gpu_ncu.sh
W=cfg1x TAG=r1z SKIP=5 KERNEL=pystl_fused_forward FLAGS="" bash tools/gpu/gpu_ncu.sh
W=cfg5 TAG=r1z SKIP=5 KERNEL=pystl_fused_forward FLAGS="" bash tools/gpu/gpu_ncu.sh
W=cfg4 TAG=r1z SKIP=5 KERNEL=pystl_fused_forward FLAGS="" bash tools/gpu/gpu_ncu.sh
