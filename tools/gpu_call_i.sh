#!/bin/bash
# A/B of the k_call_alleles / k_fold_edges loop variants, then the GPU tests on the default build
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
export AB_ARGS="--workload weak --no-e2e"
bash tools/ab_k1.sh "24:1536:192:-DLPS_SPLIT_CHAINS=0 -DLPS_FOLD_PINGPONG=0" "24:1536:192:-DLPS_SPLIT_CHAINS=1 -DLPS_FOLD_PINGPONG=0" "24:1536:192:-DLPS_SPLIT_CHAINS=0 -DLPS_FOLD_PINGPONG=1" "24:1536:192" 2>&1 | tee $O/ab_i.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_i.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest_i.log
