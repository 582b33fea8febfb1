#!/bin/bash
# Builds graphnet_classifier_b200/variants/libgnc_<name>.so for each "name:-DFLAG=1 -DOTHER=2" argument (kernel A/B runs:
# scripts/chain_ab.py, scripts/slic_ab.py load them through GNC_LIB).  Prints the spill lines of kernels matching $KERNEL.
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$root/graphnet_classifier_b200/variants"
rm -f "$root"/graphnet_classifier_b200/variants/*.so
cd "$root/graphnet_classifier_b200/csrc"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --threads 0 -Xcompiler -fPIC -shared -Xcompiler -pthread \
      -Xcompiler -ffp-contract=off $flags -Xptxas -v -o "../variants/libgnc_$name.so" graph_build.cu aggregate.cu dense.cu \
      tc_linear.cu tc_chain.cu tc_wgrad.cu tc_bwd.cu train_ops.cu narrow.cu slic.cu slic_connect.cu resize.cu jpeg.cu 2>&1 \
      | grep -A2 "${KERNEL:-NOKERNEL}" | grep -i "spill\|error" | tr '\n' ' '; echo "built $name" ) &
done
wait
