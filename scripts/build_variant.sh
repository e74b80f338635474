#!/bin/bash
# usage: build_variant.sh <git-ref|WORK> <out.so>  -- builds the library from the csrc of a git ref (or the working tree) for A/B runs (BT_LIB_PATH)
set -e
ref=$1; out=$2
tmp=$(mktemp -d)
mkdir -p $tmp/beetroots_b200
if [ "$ref" = WORK ]; then cp -r beetroots_b200/csrc $tmp/beetroots_b200/csrc; cp -r include $tmp/include; else git archive $ref beetroots_b200/csrc include | tar -x -C $tmp; fi
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -lcuda -o $out $tmp/beetroots_b200/csrc/bt_api.cu
rm -rf $tmp
