#!/bin/bash
# One plain run, then `ncu --set full` on the named kernels of the same command; the report is reduced to its raw-page
# CSV (and, with SRC=1, the SASS source page) under gpurun_out/ and deleted (reports are tens of MB, gpurun merges 64 MiB).
#   tools/ncu_capture.sh <name> <kernel-regex> <skip> <count> -- <command ...>
set -u
name=$1; regex=$2; skip=$3; count=$4; shift 5
"$@" > gpurun_out/${name}_plain.log 2>&1 || { echo "plain run failed: $name"; tail -5 gpurun_out/${name}_plain.log; exit 1; }
ncu --set full --clock-control none ${SRC:+--import-source on} -k regex:"$regex" -s "$skip" -c "$count" -o gpurun_out/${name} "$@" \
    > gpurun_out/${name}_ncu.log 2>&1
ncu -i gpurun_out/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
if [ -n "${SRC:-}" ]; then ncu -i gpurun_out/${name}.ncu-rep --page source --csv --print-source sass > gpurun_out/${name}_sass.csv 2>/dev/null; fi
rm -f gpurun_out/${name}.ncu-rep
tail -1 gpurun_out/${name}_ncu.log
