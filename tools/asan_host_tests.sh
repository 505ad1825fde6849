#!/bin/sh
# Host-side code of the library (JSON parameter parsing, DB snapshot I/O, clique finder, sampler, rigid fit, shard
# arithmetic) under AddressSanitizer + UndefinedBehaviorSanitizer, through the host-only entry points (no GPU needed).
# SURVEY.md §5 lists two latent UB sources in the reference on this path (quirks Q11, Q13); this is the check that the
# restatement has none.  Usage: sh tools/asan_host_tests.sh   (from the repo root)
set -e
export TOD_B200_VARIANT=asan
export TOD_B200_DEFINES="-Xcompiler=-fsanitize=address -Xcompiler=-fsanitize=undefined -Xcompiler=-fno-omit-frame-pointer"
export TOD_B200_LINK_FLAGS="-Xcompiler=-fsanitize=address -Xcompiler=-fsanitize=undefined"
python -m tod_b200._build
unset TOD_B200_VARIANT TOD_B200_DEFINES TOD_B200_LINK_FLAGS
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
TOD_B200_LIB="$PWD/tod_b200/libtod_b200_asan.so" \
python -m pytest tests/test_host_geometry.py tests/test_snapshot.py tests/test_abi.py -q -m "not gpu" -p no:cacheprovider
rm -rf tod_b200/libtod_b200_asan.so tod_b200/build_asan
