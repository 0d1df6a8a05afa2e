#!/bin/bash
# usage: tools/gpu/grun.sh <log name> <timeout s> [--gpus N] -- <command...>   (retries while the pod answers busy)
name=$1; to=$2; shift 2
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift 2; fi
[ "$1" = "--" ] && shift
for i in $(seq 1 ${GRUN_TRIES:-30}); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > gpurun_out/$name.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/$name.log; then break; fi
  sleep 45
done
echo "grun $name rc=$rc" >> gpurun_out/$name.log
exit $rc
