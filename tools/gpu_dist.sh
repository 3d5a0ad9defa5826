#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01i}
timeout 600 python -m pytest tests/test_distance.py -m gpu -x -q > gpurun_out/pytest_dist_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_dist_$TAG.log
