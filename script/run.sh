#!/bin/bash
# The counterpart of the reference's script/run.sh (cmake + make + `./src/virgo_plus_run ../data/SHA256_64.pws`):
# builds libvirgo_b200.so (nvcc, sm_100a), compiles the reference's UNMODIFIED main.cpp / verifier.cpp / polynomial commitment
# from a checkout of TAMUCrypto/virgo-plus against virgo-plus_b200/host/prover.h, links them with the B200 prover
#  and proves + verifies a .pws circuit on the GPU.
#   VIRGO_PLUS=/path/to/virgo-plus script/run.sh [circuit.pws]        (default checkout: /root/reference)
#   several GPUs of one box: VP_WORLD=N script/run.sh [circuit.pws]   (N copies of the program, one per GPU, one sharded prover)
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${VIRGO_PLUS:-/root/reference}"
PWS="${1:-$REF/data/SHA256_64.pws}"
[ -f "$REF/src/verifier.cpp" ] || { echo "no virgo-plus checkout at $REF (set VIRGO_PLUS)" >&2; exit 2; }
make -s -C "$HERE/virgo-plus_b200"
make -s -C "$HERE/oracle" REF="$REF" "$HERE/oracle/_ref/virgo_plus_run_b200" 2>&1 | grep -v "executable stack\|deprecated and will be removed" || true
EXE="$HERE/oracle/_ref/virgo_plus_run_b200"
if [ "${VP_WORLD:-1}" -gt 1 ]; then
    export VP_NCCL_ID_FILE="${VP_NCCL_ID_FILE:-$(mktemp -u /tmp/vp_nccl_id.XXXXXX)}"
    for r in $(seq 0 $((VP_WORLD - 1))); do VP_RANK=$r "$EXE" "$PWS" & done
    wait
else
    "$EXE" "$PWS"
fi
