#!/bin/bash
# variant_bench.sh for several SMH_B200_ROWS_PER_BAND values (GPU box). usage: variant_rows.sh "32 64 128" [bench args]
cd "$(dirname "$0")/.."
rows="$1"; shift
for r in $rows; do echo "== rows_per_band $r"; SMH_B200_ROWS_PER_BAND=$r bash tools/variant_bench.sh --steps 3 "$@"; done
