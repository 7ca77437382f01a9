#!/bin/sh
# Builds tests/stubs/build/adapter_demo against the reference's abstract headers.  Needs the reference checkout
# (first argument, default /root/reference): it is read for its HEADERS only, at build time, in the build container.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
mkdir -p "$HERE/build"
g++ -std=c++11 -O1 -Wall -I"$HERE/include" -I"$REF/aicp_core/include" -I"$ROOT/include" \
    "$HERE/adapter_demo.cpp" -o "$HERE/build/adapter_demo" \
    -L"$ROOT/aicp_mapping_b200/lib" -laicp_b200 -Wl,-rpath,'$ORIGIN/../../../aicp_mapping_b200/lib'
echo "$HERE/build/adapter_demo"
