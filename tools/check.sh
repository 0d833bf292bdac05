#!/bin/bash
# check.sh -- the reference's acceptance procedure (scripts/check_it.sh:25-59) with the
# `salloc ... srun ./kmer_hash_19` launch line (:35) replaced by this repo's binary, and K
# taken from the file instead of being hard-coded to 19.
#
#   tools/check.sh <input_file.txt>         expects <dir>/<root>_solution.txt next to the input
#   KH_GPU_SORT=1 tools/check.sh <input>     the binary also writes the sorted contig set itself (KH_SOLUTION: contigs
#                                            sorted on the GPU, per-rank lists merged on the host) -- no `sort` here
#   KH_RANKS=N tools/check.sh <input>        N ranks (GPUs), N files test_0.dat .. test_<N-1>.dat, as `srun -n N` gives
#
# Steps kept verbatim: remove old test_*.dat (:32), run in `test` mode, `cat test_*.dat | sort`
# into <root>_test.txt (:47-48), `diff -q` against the solution, print PASSED/FAILED (:55-59).
if [ $# -ne 1 ]; then
    echo "Usage: $0 <input_file>"
    exit 1
fi
INPUT_FILE=$1
HERE=$(cd "$(dirname "$0")/.." && pwd)
ROOT_NAME=$(basename "$INPUT_FILE" .txt)
INPUT_DIR=$(dirname "$INPUT_FILE")
EXPECTED_FILE="${INPUT_DIR}/${ROOT_NAME}_solution.txt"
OUTPUT_FILE="${ROOT_NAME}_test.txt"
K=$(head -c 200 "$INPUT_FILE" | awk 'NR==1{print length($1)}')
BIN="${HERE}/kmer_hash_${K}"
if [ ! -x "$BIN" ]; then
    echo "ERROR: no binary for K=${K} (${BIN}); build with: python -m cs267_hw3_b200.build"
    exit 1
fi
rm -f test_*.dat
if [ -n "$KH_GPU_SORT" ]; then export KH_SOLUTION="$OUTPUT_FILE"; fi
CMD="$BIN $INPUT_FILE test"
echo "Running command: $CMD"
if ! eval "$CMD"; then
    echo "ERROR: Execution failed."
    exit 1
fi
if [ -n "$KH_GPU_SORT" ] && [ -f "$OUTPUT_FILE" ]; then
    :                                        # already sorted by the binary
elif ls test_*.dat 1> /dev/null 2>&1; then
    cat test_*.dat | LC_ALL=C sort > "$OUTPUT_FILE"
else
    echo "ERROR: Missing output files from ranks"
    exit 1
fi
if diff -q "$OUTPUT_FILE" "$EXPECTED_FILE"; then
    echo "PASSED: $INPUT_FILE"
else
    echo "FAILED: $INPUT_FILE"
    exit 1
fi
