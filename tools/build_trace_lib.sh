#!/bin/bash
# Builds consistencytta_b200/libctta_trace.so: libctta with resblock_pair.cu compiled -DRBP_TRACE=1 (tools/trace_pair.py)
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -v deprecated || true
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DRBP_TRACE=1 -c consistencytta_b200/csrc/resblock_pair.cu -o /tmp/rbp_trace.o 2>&1 | grep -v deprecated || true
nvcc -shared -o consistencytta_b200/libctta_trace.so $(ls consistencytta_b200/build/*.o | grep -v resblock_pair) /tmp/rbp_trace.o 2>&1 | grep -v deprecated || true
ls -la consistencytta_b200/libctta_trace.so
