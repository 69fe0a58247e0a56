#!/bin/bash
# bf16 pair edge kernels: CTAs per SM (register cap) A/B.  Rebuilds libisg.so on the box; restores the default at the end.
cd "$(dirname "$0")/.."
for cfg in "5 4" "4 4" "4 3" "6 4" "5 3"; do
  set -- $cfg
  ISG_NVCC_EXTRA="-DISG_PAIR_FWD_CTAS=$1 -DISG_PAIR_DST_CTAS=$2" python intrinsic-subgraph-generation-for-vqa_b200/build.py --force > /dev/null 2>&1
  echo "## fwd CTAs/SM=$1 dst CTAs/SM=$2"
  timeout 120 python scripts/bench_edge.py --bf16 2>&1 | cut -c1-250
done
python intrinsic-subgraph-generation-for-vqa_b200/build.py --force > /dev/null 2>&1
