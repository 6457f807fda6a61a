#!/bin/bash
# Builds A/B variants of the library into denovo-variant-caller-java_b200/lib/variants/<name>.so (not tracked).
# usage: build_variants.sh name1="-DFLAG=.." name2="..."
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
CS=$ROOT/denovo-variant-caller-java_b200/csrc
OUT=$ROOT/denovo-variant-caller-java_b200/lib/variants
mkdir -p $OUT
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  rm -f $CS/trio.o $CS/pileup.o $CS/mendel.o
  make -s -C $CS EXTRA="$flags" OUT=$OUT/$name.so > /dev/null
  echo "built $name ($flags)"
done
rm -f $CS/trio.o $CS/pileup.o $CS/mendel.o
make -s -C $CS > /dev/null
