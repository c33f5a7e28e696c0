#!/bin/sh
# Developer aid: compile the CUDA sources for the CPU fiber emulator (see cuda_emul.h).
# Output: tools/cpu_emul/libsoftgnss_emul.so (git-ignored, never loaded by the package).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
SRC=$ROOT/softgnss_python_b200/csrc
FILES=""
for f in sgx_api.cu sgx_track.cu sgx_synth.cu sgx_acq.cu sgx_bitsync.cu sgx_nav.cu; do
  [ -f "$SRC/$f" ] && FILES="$FILES $SRC/$f"
done
g++ -x c++ -std=c++17 -O2 -g -DSGX_EMUL -ffp-contract=off -fPIC -shared -pthread \
    -Wno-unknown-pragmas -Wno-attributes -I"$HERE" -I"$SRC" $FILES -o "$HERE/libsoftgnss_emul.so"
echo "$HERE/libsoftgnss_emul.so"
