#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ARGS="--workload weak --no-e2e" bash tools/ab_k1.sh "$@" 2>&1 | tee gpurun_out/ab_k1_latest.log
