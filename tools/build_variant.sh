#!/bin/bash
# Builds a variant of the library for kernel experiments:  tools/build_variant.sh <suffix> <extra nvcc flags...>
# -> llampc_b200/libllampc_b200_<suffix>.so (use with LLAMPC_LIB=...); the default library is rebuilt afterwards by `make`.
set -e
cd "$(dirname "$0")/../llampc_b200/csrc"
suffix=$1; shift
rm -rf build_$suffix
make --no-print-directory -j$(nproc) OBJDIR=build_$suffix OUT=../libllampc_b200_$suffix.so EXTRA="$*" ../libllampc_b200_$suffix.so > /dev/null
grep -h -A2 "lookback_equal_kernelILb1ELb1" build_$suffix/lookback_k1e.log | grep -E "registers|spill" || true
rm -rf build_$suffix
