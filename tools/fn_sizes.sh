#!/bin/bash
# usage: fn_sizes.sh <cubin> : device function sizes inside k_rate, largest first
readelf -sW "$1" 2>/dev/null | awk '$4=="FUNC" && $8 ~ /6k_rateE/ {print $3, $8}' | sed -E 's/\$_ZN4hmp36k_rateE[^$]*\$//' | sed -E 's/_ZN[0-9]+_INTERNAL_[0-9a-f_]+pipeline_cu_[0-9a-f]+4hmp3[0-9]+//; s/_ZN[0-9]+_INTERNAL_[^ ]*4hmp3[0-9]+//' | sort -rn | head -${2:-30}
readelf -sW "$1" 2>/dev/null | awk '$4=="FUNC" && $8 ~ /6k_rateE/ {s+=$3} END {print "total bytes (functions):", s}'
