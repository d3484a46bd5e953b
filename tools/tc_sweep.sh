for m in tf32 tf32x3; do
for cfg in 1,4 2,4 2,3 3,3 3,2 4,2 4,3 5,2; do
  NBPC_MATH=$m NBPC_GLT_FWD=$cfg python tools/tc_tune.py 2>&1 | grep "edge_out"
done
for cfg in 1,4 1,6 2,2 2,3 3,2; do
  NBPC_MATH=$m NBPC_GLT_BWD=$cfg python tools/tc_tune.py 2>&1 | grep "edge_bwd"
done
done
NBPC_MATH=fp32 python tools/tc_tune.py 2>&1 | grep edge
