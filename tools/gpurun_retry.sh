#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout-seconds> [--gpus N] -- <command...>
# retries while gpurun answers "no box / slot free" (exit 3), every 90 s, up to 40 times; log in gpurun_out/<tag>.log
TAG=$1; TMO=$2; shift 2
LOG=gpurun_out/$TAG.log; mkdir -p gpurun_out; : > $LOG
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" >> $LOG 2>&1; rc=$?
  echo "[retry] attempt $i rc=$rc" >> $LOG
  if [ $rc -ne 3 ] && ! grep -q "status=transient" <(tail -5 $LOG); then break; fi
  sleep 90
done
echo "[retry] finished rc=$rc" >> $LOG
