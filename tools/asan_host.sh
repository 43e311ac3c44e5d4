#!/bin/bash
# Builds the C++ host library with AddressSanitizer and runs the CPU host tests against it (the ctypes-loaded library needs libasan
# preloaded; leak detection is off because the Python interpreter itself is not leak-clean).  Usage: tools/asan_host.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
H=$ROOT/longphase-s_b200/host
B=$(mktemp -d)
for f in host_common phase_host haplotag_host somatic_host; do
  g++ -std=c++17 -O1 -g -fPIC -fopenmp -fsanitize=address -fno-omit-frame-pointer -I$H -I$ROOT/include -I${REF:-/root/reference}/htslib -c $H/$f.cpp -o $B/$f.o
done
g++ -shared -fopenmp -fsanitize=address -o $B/liblps_host.so $B/*.o $H/_build/libhts.a -L$ROOT/longphase-s_b200 -llps_b200 -Wl,-rpath,$ROOT/longphase-s_b200 -lz -lpthread -lm
cp $ROOT/longphase-s_b200/liblps_host.so $B/liblps_host.so.orig
trap 'cp $B/liblps_host.so.orig $ROOT/longphase-s_b200/liblps_host.so' EXIT
cp $B/liblps_host.so $ROOT/longphase-s_b200/liblps_host.so
cd $ROOT
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  python -m pytest tests/test_host_cli.py tests/test_host_somatic_cli.py tests/test_host_golden.py -x -q -m "not gpu" -k "not a_gpu"
