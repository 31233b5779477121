#!/bin/bash
# Bounds check of every kernel body without a GPU: the emulation harness (tests/emul, kernel bodies as host loops) built with
# AddressSanitizer, then the CPU-tier engine, encryption and codec parity tests on it. "Device" buffers are heap blocks and the shared-memory window of a CTA
# is a heap block of exactly the size the launch asks for, so an out-of-range index in a kernel is an ASan report here.
# usage: tools/asan_emul.sh [pytest -k expression]
set -eu
cd "$(dirname "$0")/.."
GXX=""
for c in ${CXX:-g++} g++ /usr/bin/g++; do  # a compiler that ships the sanitizer runtime
  case "$($c -print-file-name=libasan.so 2>/dev/null)" in /*) GXX=$c; break;; esac
done
[ -n "$GXX" ] || { echo "no g++ with libasan found"; exit 2; }
make -s -C tests/emul SAN=_asan CXX=$GXX
ASAN=$($GXX -print-file-name=libasan.so)  # libstdc++ is preloaded too: the interceptor of __cxa_throw needs it and python is not a C++ program
LD_PRELOAD="$ASAN $($GXX -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:log_path=${ASAN_LOG:-/tmp/hhe_asan} HHE_EMUL_LIB=$PWD/tests/emul/libhhe_emul_asan.so \
  python -m pytest tests/test_engine_parity.py tests/test_encrypt.py tests/test_seal_codec.py -m "not gpu" -x -q ${1:+-k "$1"}
