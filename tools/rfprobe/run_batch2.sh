#!/bin/bash
cd "$(dirname "$0")"
for f in variants/k_ffma__r_*.cubin; do ./run k_ffma 256 2000 96 $f 2>&1; done
for f in variants/k_ffma2__p2_bgroup*.cubin; do ./run k_ffma2 256 2000 96 $f 2>&1; done
