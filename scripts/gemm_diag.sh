#!/bin/bash
# Diagnostic (results are WRONG by construction): rebuilds libisg.so with parts of the GEMM splitter's per-k-block work
# removed (-DISG_TC_DIAG bits, csrc/linear_tc.cu) and times the E-sized products.  Restores the normal build at the end.
cd "$(dirname "$0")/.."
for bits in 1 2 3 6 7; do
  ISG_NVCC_EXTRA="-DISG_TC_DIAG=$bits" python intrinsic-subgraph-generation-for-vqa_b200/build.py --force > /dev/null 2>&1
  echo "## ISG_TC_DIAG=$bits"
  timeout 120 python scripts/gemm_probe.py 2>&1 | head -2 | cut -c1-170
done
python intrinsic-subgraph-generation-for-vqa_b200/build.py --force > /dev/null 2>&1
echo "## normal build"
timeout 120 python scripts/gemm_probe.py 2>&1 | head -2 | cut -c1-170
