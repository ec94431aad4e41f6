#!/bin/bash
# SASS of one Fq Montgomery product (fp.cuh mul_cios) for sm_100a: mnemonic histogram + full listing -> profiles/r02_sass_mul_cios.txt
set -eu
cd "$(dirname "$0")/.."
OUT=profiles/r02_sass_mul_cios.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -cubin -o /tmp/mul_sass.cubin tools/probes/mul_sass.cu
cuobjdump -sass -fun k_one_mul_fq /tmp/mul_sass.cubin > /tmp/mul_sass.txt
{
  echo "# cuobjdump -sass of k_one_mul_fq (tools/probes/mul_sass.cu: out[i] = a[i] * b[i], Fp<FqParams>::mul_cios), sm_100a, nvcc $(nvcc --version | grep -o 'V[0-9.]*')"
  echo "# mnemonic histogram (whole kernel: 2 x 32-byte loads, one product, one store):"
  grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/mul_sass.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn
  echo
  echo "# instruction stream:"
  grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/mul_sass.txt | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///'
} > $OUT
head -30 $OUT
