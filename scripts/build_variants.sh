#!/bin/bash
# Builds tuning variants of libmdacp_b200.so into build/variants/ (git-ignored; travels with gpurun).
# usage: scripts/build_variants.sh "G U B" ["G U B" ...]   (lanes per local, unroll, min blocks per SM)
set -e
cd "$(dirname "$0")/../mdacp_b200/csrc"
mkdir -p ../../build/variants
for v in "$@"; do
  set -- $v
  G=$1; U=$2; B=$3
  name=g${G}_u${U}_b${B}
  make -s -j8 OBJDIR=../../build/var_$name LIB=../../build/variants/lib_$name.so \
    EXTRA="-DLANES_PER_LOCAL=$G -DFORCE_UNROLL=$U -DFORCE_MINBLK=$B" 2>&1 | grep -i "error" || true
  echo built $name
done
