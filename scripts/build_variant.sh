#!/bin/bash
# Build a tuning variant of libb200tree.so next to the default one:
#   scripts/build_variant.sh NAME -DPT_ITEMS_CFG=8 -DPT_WARPS_CFG=16 ...   ->  np_decision_tree_b200/variants/libb200tree_NAME.so
# run it with B200DT_LIB=np_decision_tree_b200/variants/libb200tree_NAME.so (see _lib.load).
set -e
cd "$(dirname "$0")/../np_decision_tree_b200"
name=$1; shift
mkdir -p variants/obj_$name
pids=""
for s in sort l2 weighted partition relabel children predict l1 classes rows fit; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr \
       -diag-suppress 550 "$@" -c csrc/$s.cu -o variants/obj_$name/$s.o &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libb200tree_$name.so variants/obj_$name/*.o -lcudart
rm -rf variants/obj_$name
ls -la variants/libb200tree_$name.so
