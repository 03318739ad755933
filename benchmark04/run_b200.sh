#!/bin/bash
# Sweep of the polynomial orders the reference run.sh (shipped unchanged next to this file) covers; one log per
# order next to this script, in the names postprocess.py globs for.
# The reference pins CUDA_VISIBLE_DEVICES=1; here the device is whatever
# B200FE_DEVICE says (default: leave the environment alone).
set -u
here="$(cd "$(dirname "$0")" && pwd)"
orders="${B200FE_ORDERS:-2 4 6 8 10 12 14 16 32}"
[ -n "${B200FE_DEVICE:-}" ] && export CUDA_VISIBLE_DEVICES="$B200FE_DEVICE"
for nq in $orders; do
  echo "nq=$nq"
  "$here/build/benchmark04" "$nq" "$nq" &> "$here/nq${nq}x${nq}.log"
done
