#!/bin/bash
# gpurun with retry while the pod answers "transient" (busy): scripts/grun.sh [gpurun options] -- 'command'
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); echo "$out" | tail -60
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  echo "[grun] busy, retry $i"; sleep 120
done
