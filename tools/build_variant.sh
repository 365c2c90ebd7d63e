#!/bin/sh
# Builds an instrumented copy of the library next to the product one:  tools/build_variant.sh trace -DSPB_TRACE
# -> variants/lib_<name>.so (git-ignored; travels to the GPU box), then restores the production objects.
set -e
name=$1; shift
cd "$(dirname "$0")/../self-play-ai_b200/csrc"
mkdir -p ../../variants
make -B -s OUT=../../variants/lib_$name.so VARIANT="$*"
make -B -s
