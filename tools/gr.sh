#!/bin/bash
# build in-tree, then run a command on the GPU box:  tools/gr.sh [--gpus N] <timeout_s> '<command>'
set -e
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
make -C "$(dirname "$0")/../slamkit_b200/csrc" -j8 2>&1 | grep -E "error|Error" && exit 1
T=$1; shift
exec /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
