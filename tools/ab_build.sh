#!/bin/bash
# Build the library of a git ref (or the working tree with "WORK") into 3dr_b200/lib/libdr3lk_<tag>.so for A/B timing.
# usage: tools/ab_build.sh <ref|WORK> <tag> [extra nvcc flags]
set -e
ref=$1; tag=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
if [ "$ref" = "WORK" ]; then cp -r $root/3dr_b200 $root/include $tmp/; else (cd $root && git archive $ref 3dr_b200 include | tar -x -C $tmp); fi
make -C $tmp/3dr_b200/csrc -j4 EXTRA="$3" > /dev/null 2>&1 || make -C $tmp/3dr_b200/csrc EXTRA="$3"
cp $tmp/3dr_b200/lib/libdr3lk.so $root/3dr_b200/lib/libdr3lk_$tag.so
rm -rf $tmp
echo built 3dr_b200/lib/libdr3lk_$tag.so
