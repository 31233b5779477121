#!/bin/bash
# Race check without a GPU: the emulation harness runs the threads of each barrier-separated phase one after the other, so a
# phase whose result depends on the ORDER of its threads is a missing barrier (or warp barrier) in the CUDA kernel. This builds the
# harness with the threads of every phase in descending (SAN=_rev) and in scrambled (SAN=_shuf) order and runs the CPU-tier parity
# tests on both: they must still match the oracle bit for bit.
# usage: tools/order_emul.sh [pytest -k expression]
set -eu
cd "$(dirname "$0")/.."
for v in _rev _shuf; do
  make -s -C tests/emul SAN=$v
  echo "== thread order variant $v"
  HHE_EMUL_LIB=$PWD/tests/emul/libhhe_emul$v.so python -m pytest tests/test_engine_parity.py tests/test_encrypt.py -m "not gpu" -x -q ${1:+-k "$1"}
done
