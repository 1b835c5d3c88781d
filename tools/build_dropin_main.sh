#!/usr/bin/env bash
# Drop-in proof: compiles the REFERENCE's own benchmark driver, CSR5_cuda/main.cu,
# UNMODIFIED against this repository's include/anonymouslib_cuda.h and links it with libcsr5_b200.so.
# main.cu and its Matrix-Market reader mmio.h are copied to a scratch directory for the build
# (main.cu's quote-includes resolve relative to its own directory) and deleted afterwards; only the
# binary is kept, in oracle/_ref/ (git-ignored, travels to the GPU box).
#   usage: tools/build_dropin_main.sh [double|float]   ->  oracle/_ref/spmv_dropin_<type>
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REFERENCE="${REFERENCE:-/root/reference}"
VT="${1:-double}"
if [ ! -f "$REFERENCE/CSR5_cuda/main.cu" ]; then
    echo "$REFERENCE/CSR5_cuda/main.cu absent: keeping prebuilt binary (if any)"; exit 0
fi
SCRATCH="$(mktemp -d /tmp/csr5_dropin.XXXXXX)"
trap 'rm -rf "$SCRATCH"' EXIT
cp "$REFERENCE/CSR5_cuda/main.cu" "$REFERENCE/CSR5_cuda/mmio.h" "$SCRATCH/"
cp "$ROOT/include/anonymouslib_cuda.h" "$ROOT/include/csr5_b200.h" "$SCRATCH/"
mkdir -p "$ROOT/oracle/_ref"
LIBDIR="$ROOT/benchmark_spmv_using_csr5_b200"
nvcc -O3 -w -m64 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fpermissive \
     -D VALUE_TYPE="$VT" -D NUM_RUN="${NUM_RUN:-1000}" "$SCRATCH/main.cu" -o "$ROOT/oracle/_ref/spmv_dropin_$VT" \
     -L"$LIBDIR" -lcsr5_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../benchmark_spmv_using_csr5_b200'
echo "built oracle/_ref/spmv_dropin_$VT: reference main.cu (unmodified) + include/anonymouslib_cuda.h + libcsr5_b200.so"
