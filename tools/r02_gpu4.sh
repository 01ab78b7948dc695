#!/bin/bash
# GPU tests only (arguments: tag, then pytest selectors), without -x: every failure is listed
mkdir -p gpurun_out
tag=$1; shift
( timeout 1500 python -m pytest "$@" -m gpu -q 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest.txt
tail -15 gpurun_out/${tag}_pytest.txt
