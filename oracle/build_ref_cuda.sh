#!/usr/bin/env bash
# oracle/build_ref_cuda.sh -- builds oracle/_ref/libref_cuda.so: the reference's OWN CSR5_cuda backend
# (anonymouslib_cuda.h + detail/cuda/*.h) behind a small C driver (oracle/ref_cuda_driver.cu).
#
# TEST INFRASTRUCTURE ONLY.  No reference source is copied into this repository: the reference tree is
# copied to a scratch directory under /tmp, given the five-point CUDA-12 / sm_100a compatibility patch
# of SURVEY.md App. C there (the code targets Kepler/Maxwell and does not compile as shipped), compiled,
# and the scratch directory is deleted.  Only the binary lands in oracle/_ref/ (git-ignored; it travels
# to the GPU box with the snapshot).  None of the five points changes the arithmetic:
#   1. stub helper_cuda.h / helper_functions.h (un-vendored CUDA-samples headers: only checkCudaErrors)
#   2. drop the legacy `double __shfl_*` overloads and the CAS-loop atomicAdd(double) that collide with
#      toolkit builtins (the builtins do the same thing)
#   3. map __shfl* -> __shfl*_sync(0xffffffff, ...) (non-sync shuffles are rejected for >= sm_70)
#   4. shfl.up.b32 -> shfl.sync.up.b32 in the inline-PTX FP64 scan
#   5. -Xcompiler -fpermissive (a negative constant shifted in an enum initialiser)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REFERENCE="${REFERENCE:-/root/reference}"
if [ ! -d "$REFERENCE/CSR5_cuda" ]; then
    echo "$REFERENCE/CSR5_cuda absent: keeping prebuilt oracle/_ref/libref_cuda.so (if any)"
    exit 0
fi
SCRATCH="$(mktemp -d /tmp/csr5_refcuda.XXXXXX)"
trap 'rm -rf "$SCRATCH"' EXIT
cp -r "$REFERENCE/CSR5_cuda/." "$SCRATCH/"
chmod -R u+w "$SCRATCH"
mkdir -p "$SCRATCH/stub" "$HERE/_ref"

cat > "$SCRATCH/stub/helper_functions.h" <<'EOS'
#pragma once
EOS
cat > "$SCRATCH/stub/helper_cuda.h" <<'EOS'
#pragma once
#include <cstdio>
#include <cstdlib>
#define checkCudaErrors(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
EOS
cat > "$SCRATCH/stub/compat.h" <<'EOS'
#pragma once
#include <cuda_runtime.h>
#define __shfl_down(...) __shfl_down_sync(0xffffffffu, __VA_ARGS__)
#define __shfl_up(...)   __shfl_up_sync(0xffffffffu, __VA_ARGS__)
#define __shfl_xor(...)  __shfl_xor_sync(0xffffffffu, __VA_ARGS__)
#define __shfl(...)      __shfl_sync(0xffffffffu, __VA_ARGS__)
EOS

python3 - "$SCRATCH/detail/cuda/utils_cuda.h" <<'EOP'
import re, sys
p = sys.argv[1]
s = open(p).read()
# point 2a: the `#if __CUDA_ARCH__ <= 300` block holding the double shuffle overloads
a = s.index("#if __CUDA_ARCH__ <= 300")
b = s.index("#endif", a) + len("#endif")
assert "__shfl_xor(double" in s[a:b].replace("\n", " ").replace("double __shfl_xor(double", "__shfl_xor(double")
s = s[:a] + s[b:]
# point 2b: CAS-loop atomicAdd(double*, double)
a = s.index("static double atomicAdd(double *addr, double val)")
a = s.rfind("__forceinline__", 0, a)
b = s.index("return old;", a)
b = s.index("}", b) + 1
s = s[:a] + s[b:]
# point 4: inline PTX shuffles
n0 = s.count("shfl.up.b32")
s = re.sub(r"shfl\.up\.b32 (lo|hi)\|p, (lo|hi), %2, %3;", r"shfl.sync.up.b32 \1|p, \2, %2, %3, 0xffffffff;", s)
assert n0 == 2 and s.count("shfl.sync.up.b32") == 2
open(p, "w").write(s)
EOP

nvcc -O3 -w -m64 -std=c++14 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fpermissive -Xcompiler -fPIC \
     -shared -I"$SCRATCH" -I"$SCRATCH/stub" -include "$SCRATCH/stub/compat.h" \
     "$HERE/ref_cuda_driver.cu" -o "$HERE/_ref/libref_cuda.so"
echo "built oracle/_ref/libref_cuda.so from $REFERENCE/CSR5_cuda (compat-patched scratch copy, deleted)"
