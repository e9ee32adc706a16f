#!/bin/bash
# runs every patched cubin at 1 and 2 warps per sub-partition (one process per cubin: a bad encoding kills the context)
cd "$(dirname "$0")"
for thr in 128 256; do
  for f in base.cubin variants/k_ffma2__*.cubin; do ./run k_ffma2 $thr 2000 96 $f 2>&1; done
  for f in base.cubin variants/k_ffma__*.cubin; do ./run k_ffma $thr 2000 96 $f 2>&1; done
done
