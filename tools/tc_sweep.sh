export NBPC_MATH=tf32x3
for cfg in 3,2,1 2,4,1 2,3,2 2,5,1; do
  NBPC_GLT_FWD=$cfg python tools/tc_tune.py 2>&1 | grep "edge_out"
done
for cfg in 1,4,1 1,3,2 1,3,1 1,2,2; do
  NBPC_GLT_BWD=$cfg python tools/tc_tune.py 2>&1 | grep "edge_bwd"
done
