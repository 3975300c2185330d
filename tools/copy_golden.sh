#!/bin/sh
# Copies the reference's golden vectors (data) into tests/golden/.
# Run in the build container, where /root/reference exists.
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/golden/xoshiro-ref
cp /root/reference/tests/testthat/xoshiro-ref/* tests/golden/xoshiro-ref/
