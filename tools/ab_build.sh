#!/bin/sh
# A/B builds of the kernel library (development aid): tools/ab_build.sh NAME "-DKNOB=..." builds
# variants/NAME/libmonty_b200.so (variants/ is git-ignored but travels with gpurun); a gpurun
# command then copies one variant at a time over monty_b200/libmonty_b200.so and benches it.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/variants/$name"
make -s -j8 -C "$root/monty_b200/csrc" OBJDIR="$root/variants/$name/build" TARGET="$root/variants/$name/libmonty_b200.so" EXTRA="$*" "$root/variants/$name/libmonty_b200.so"
