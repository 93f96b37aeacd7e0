#!/bin/bash
# Builds tuning variants of libmdacp_b200.so into build/variants/ (git-ignored; travels with gpurun).
# usage: scripts/build_variants.sh "name -DFLAG=.. -DFLAG=.." ...
set -e
cd "$(dirname "$0")/../mdacp_b200/csrc"
mkdir -p ../../build/variants
for v in "$@"; do
  set -- $v
  name=$1; shift
  make -s -j8 OBJDIR=../../build/var_$name LIB=../../build/variants/lib_$name.so EXTRA="$*" 2>&1 | grep -i "error" || true
  echo built $name
done
