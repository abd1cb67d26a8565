#!/bin/bash
# ASan + UBSan over the host-side C code (oracle + synthetic IQ source).  usage: tools/san/run.sh
set -e
cd "$(dirname "$0")/../.."
mkdir -p /tmp/xrd_san
gcc -O1 -g -std=gnu11 -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -ffp-contract=off \
    -mavx2 -mfma -fopenmp -Wall -Wextra -o /tmp/xrd_san/host_sanitize \
    tools/san/host_sanitize.c oracle/xrit_oracle.c xritdemod_b200/csrc/siggen.c -lm
ASAN_OPTIONS=detect_leaks=1:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 /tmp/xrd_san/host_sanitize
