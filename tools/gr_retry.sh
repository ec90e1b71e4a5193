#!/bin/bash
# tools/gr_retry.sh <logfile> [--gpus N] <timeout_s> '<command>' : like gr.sh, but retries while the pod answers busy/transient
LOG=$1; shift
for i in $(seq 1 30); do
  "$(dirname "$0")/gr.sh" "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient\|status=busy\|rc=None" "$LOG"; then exit 0; fi
  sleep 90
done
