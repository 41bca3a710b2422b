#!/bin/bash
# Developer tool (GPU box): ncu --set full capture of ONE kernel of a bench command, exported to small CSVs.
#   tools/ncu_capture.sh <tag> <kernel-regex> <skip> -- <command...>
# Writes gpurun_out/<tag>_raw.csv and gpurun_out/<tag>_source.csv.gz and removes the .ncu-rep (gpurun brings back
# at most 64 MiB).  The same command must have just exited 0 without ncu (B200_PROFILING.md).
set -u
tag=$1; regex=$2; skip=$3; shift 4
"$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run of $tag failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c 1 -f -o /tmp/${tag} "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_source.csv.gz
rm -f /tmp/${tag}.ncu-rep
tail -1 gpurun_out/${tag}_ncu.log
