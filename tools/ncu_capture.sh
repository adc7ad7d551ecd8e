#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <skip> <cmd...>
# Captures ONE launch with ncu --set full, exports raw/details/source pages as CSV into gpurun_out/ and
# deletes the (large) .ncu-rep so the 64 MiB copy-back limit holds.
name=$1; regex=$2; skip=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/${name}_details.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
rm -f /tmp/$name.ncu-rep
